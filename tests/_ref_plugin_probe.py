"""Probe run in a FRESH interpreter by tests/test_reference_plugin_api.py (build container only: needs /root/reference).

Drives a TINY ``B200UNet`` with the REFERENCE's own plugin code -- ``IPAdapter.set_ip_adapter / load_ip_adapter / set_scale /
enable / disable`` AST-extracted from diffusion/ip_adapter/ip_adapter.py:120-169,211-214 and executed unmodified -- with the
reference's ``attention_processor`` module importable under its real dotted name, so that ``B200IPAttnProcessor`` takes its
subclass branch (``isinstance(p, IPAttnProcessor)`` at ip_adapter.py:213 must hold).  Arithmetic runs on the CPU test double
(tests/emu_ops.py); what is checked is the boundary: processor installation, checkpoint key landing, scale plumbing, disable()."""
import importlib.util
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import ref_shims

assert ref_shims.available()
# ---- the reference package skeleton: only the torch-only attention_processor module is real
for name in ("instructany2pix", "instructany2pix.diffusion", "instructany2pix.diffusion.ip_adapter"):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
p = os.path.join(ref_shims.REF_ROOT, "instructany2pix/diffusion/ip_adapter/attention_processor.py")
spec = importlib.util.spec_from_file_location("instructany2pix.diffusion.ip_adapter.attention_processor", p)
ref_ap = importlib.util.module_from_spec(spec)
sys.modules[spec.name] = ref_ap
spec.loader.exec_module(ref_ap)

import instructany2pix_b200.attention_processor as AP  # noqa: E402  (imported AFTER the shim: subclass branch)
import instructany2pix_b200.ops as real_ops  # noqa: E402
from tests import emu_ops  # noqa: E402

for n in dir(emu_ops):
    if not n.startswith("_") and callable(getattr(emu_ops, n)) and hasattr(real_ops, n):
        setattr(real_ops, n, getattr(emu_ops, n))
from instructany2pix_b200.image_proj import B200ImageProj  # noqa: E402
from instructany2pix_b200.unet import TINY_CONFIG, B200UNet  # noqa: E402
from oracle.synth import synth_state_dict  # noqa: E402
from oracle.unet import TINY, OracleUNet  # noqa: E402
from tests.test_host_unet_emu import make_inputs, rel  # noqa: E402

torch.set_grad_enabled(False)
assert AP._RefIP is ref_ap.IPAttnProcessor2_0 and issubclass(AP.B200IPAttnProcessor, ref_ap.IPAttnProcessor2_0), "subclass branch not taken"
assert issubclass(AP.B200AttnProcessor, ref_ap.AttnProcessor2_0)

# ---- the reference's IPAdapter class, unmodified source, with the names its module imports
IPAdapter = ref_shims.extract_source("instructany2pix/diffusion/ip_adapter/ip_adapter.py", "IPAdapter", extra_globals=dict(
    os=os, AttnProcessor=ref_ap.AttnProcessor2_0, IPAttnProcessor=ref_ap.IPAttnProcessor2_0, CNAttnProcessor=ref_ap.CNAttnProcessor2_0,
    MultiControlNetModel=type("MultiControlNetModel", (), {}), safe_open=None, Image=None))

o = OracleUNet(TINY).eval()
o.load_state_dict(synth_state_dict(o, 0))
unet = B200UNet.from_module(o, device="cpu")
ad = IPAdapter.__new__(IPAdapter)                       # __init__ needs the hub (CLIP image encoder): set what the methods read
ad.device, ad.num_tokens = "cpu", 4
ad.pipe = types.SimpleNamespace(unet=unet)
ad.image_proj_model = B200ImageProj(TINY.cross_attention_dim, 1024, 4, device="cpu")

# ---- set_ip_adapter (ip_adapter.py:120-148): reference processors handed to B200UNet.set_attn_processor
ad.set_ip_adapter()
procs = unet.attn_processors
names = list(procs)
assert names == list(o.attn_processors), "processor registration order (down -> up -> mid) differs from the reference UNet"
n_ip = 0
for name, pr in procs.items():
    if name.endswith("attn2.processor"):
        assert isinstance(pr, AP.B200IPAttnProcessor) and isinstance(pr, ref_ap.IPAttnProcessor2_0), name
        assert pr.scale == 1.0 and pr.num_tokens == 4
        n_ip += 1
    else:
        assert isinstance(pr, ref_ap.AttnProcessor2_0) and not hasattr(pr, "to_k_ip"), name
assert n_ip == len(names) // 2

# ---- load_ip_adapter (ip_adapter.py:155-169): a synthetic checkpoint in the released layout {"image_proj", "ip_adapter"}
ip_sd, expect = {}, {}
for idx, name in enumerate(names):
    if name.endswith("attn2.processor"):
        C = procs[name].to_k_ip.weight.shape[0]
        for w in ("to_k_ip", "to_v_ip"):
            key = f"{idx}.{w}.weight"
            t = synth_state_dict({name + key: (C, TINY.cross_attention_dim)}, 5)[name + key]
            ip_sd[key] = t
            expect[(name, w)] = t
assert all(int(k.split(".")[0]) % 2 == 1 for k in ip_sd), "IP layers must sit at the odd (attn2) positions: {2i+1}.to_k_ip.weight"
proj_sd = synth_state_dict({k: v.shape for k, v in ad.image_proj_model.state_dict().items()}, 9)
ck = os.path.join(tempfile.mkdtemp(), "ip_adapter_synth.bin")
torch.save({"image_proj": proj_sd, "ip_adapter": ip_sd}, ck)
ad.ip_ckpt = ck
ad.load_ip_adapter()
for (name, w), t in expect.items():
    got = getattr(unet.attn_processors[name], w).weight
    assert torch.equal(got.float(), t.to(got.dtype).float()), (name, w)
assert torch.equal(ad.image_proj_model.proj.weight, proj_sd["proj.weight"])

# ---- the same weights in the oracle, driven through the reference's processor classes
op = {}
for name, pr in o.attn_processors.items():
    if name.endswith("attn2.processor"):
        C = procs[name].to_k_ip.weight.shape[0]
        r = ref_ap.IPAttnProcessor2_0(C, TINY.cross_attention_dim, scale=1.0, num_tokens=4)
        r.to_k_ip.weight.copy_(expect[(name, "to_k_ip")])
        r.to_v_ip.weight.copy_(expect[(name, "to_v_ip")])
        op[name] = r
    else:
        op[name] = pr
o.set_attn_processor(op)
lat, ctx, added = make_inputs(TINY)
x = torch.cat([lat, lat])
fwd_b = lambda: unet(x, 981, ctx, added_cond_kwargs=added)[0]
fwd_o = lambda: o(x, torch.tensor(981), ctx, added_cond_kwargs=added)[0]
e1 = rel(fwd_b(), fwd_o())
assert e1 < 1.5e-2, e1

# ---- set_scale (ip_adapter.py:211-214): isinstance-gated attribute write, seen by the next forward
y1 = fwd_b()
ad.set_scale(0.3)
assert all(p.scale == 0.3 for n, p in unet.attn_processors.items() if n.endswith("attn2.processor"))
for pr in o.attn_processors.values():
    if isinstance(pr, ref_ap.IPAttnProcessor2_0):
        pr.scale = 0.3
y2 = fwd_b()
assert rel(y2, y1) > 1e-3, "set_scale did not change the forward"
e2 = rel(y2, fwd_o())
assert e2 < 1.5e-2, e2

# ---- disable() (ip_adapter.py:153-154): one plain processor for every layer -> text-only attention over ALL tokens
ad.disable()
assert all(not hasattr(p, "to_k_ip") for p in unet.attn_processors.values())
o.set_attn_processor(ref_ap.AttnProcessor2_0())
e3 = rel(fwd_b(), fwd_o())
assert e3 < 1.5e-2, e3

# ---- enable() (ip_adapter.py:149-151) = set_ip_adapter + load_ip_adapter again
ad.enable()
for (name, w), t in expect.items():
    got = getattr(unet.attn_processors[name], w).weight
    assert torch.equal(got.float(), t.to(got.dtype).float()), (name, w)
print(f"REF_PLUGIN_OK ip_layers={n_ip} forward_rel={e1:.2e} scale_rel={e2:.2e} disable_rel={e3:.2e}")
