"""The C-ABI library loads and exports exactly the symbols include/ia2p.h declares, with matching arity/types in the ctypes
binding (no compute calls: runs without a GPU).  Also: the product package must not import the oracle."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from instructany2pix_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "ia2p.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    out = {}
    for m in re.finditer(r"(ia2p_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", body):
        out[m.group(1)] = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
    return out


def test_header_binding_and_library_agree():
    decl = _declared()
    assert set(decl) == set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "instructany2pix_b200", "csrc"), "-j8"], check=True)
    lib = _lib.load()
    for name, args in decl.items():
        assert hasattr(lib, name), f"{name} not exported"
        sig = _lib.SIGNATURES[name][0]
        assert len(sig) == len(args), name
        for a, t in zip(args, sig):
            want = C.c_void_p if "*" in a else (C.c_int64 if "int64_t" in a else (C.c_float if a.startswith("float") else C.c_int))
            assert want is t, (name, a, t)
    assert lib.ia2p_version() >= 100
    # only C-ABI symbols are exported under the ia2p_ prefix
    syms = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in syms.splitlines() if " T " in ln and ln.split()[-1].startswith("ia2p_")}
    assert exported == set(decl)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from instructany2pix_b200 import ops
    with pytest.raises(_lib.IA2PError):
        ops.layernorm(torch.zeros(4, 64), torch.ones(64), torch.zeros(64), 1e-5)
    lib = _lib.load()
    assert lib.ia2p_device_check(-1) != 0          # no device -> IA2P_E_DEVICE, never a fallback
    assert b"no" in lib.ia2p_last_error().lower() or lib.ia2p_last_error()


def test_product_does_not_import_oracle():
    code = ("import sys; import instructany2pix_b200, instructany2pix_b200.unet, instructany2pix_b200.prior, "
            "instructany2pix_b200.sampler, instructany2pix_b200.scheduler, instructany2pix_b200.parallel; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'product imports oracle'")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "instructany2pix_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
