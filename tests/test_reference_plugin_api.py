"""The reference's own plugin code against the drop-in: ``IPAdapter.set_ip_adapter / load_ip_adapter /
set_scale / enable / disable`` (diffusion/ip_adapter/ip_adapter.py:120-169,211-214), AST-extracted and run unmodified against a TINY
``B200UNet`` in a fresh interpreter where ``instructany2pix.diffusion.ip_adapter.attention_processor`` is importable -- so the
``class B200IPAttnProcessor(_RefIP)`` subclass branch is the one exercised.  Build container only (needs /root/reference)."""
import os
import subprocess
import sys

import pytest

from oracle import ref_shims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not ref_shims.available(), reason="/root/reference is not present (GPU box)")
def test_reference_ipadapter_methods_drive_the_drop_in():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_ref_plugin_probe.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "REF_PLUGIN_OK" in r.stdout, r.stdout[-2000:]
    print(r.stdout.strip().splitlines()[-1])
