#!/usr/bin/env python
"""BUILD-CONTAINER ONLY (needs /root/reference): times the reference's OWN prior code -- InstructAny2PixPrior.generate_diffusion
(prior/model.py:527-658), imported unmodified under oracle/ref_shims.load_prior -- next to the oracle port of it on the same host
cores, 25-step CFG sampling of one sample with synthetic (name-seeded) weights.  The GPU box has no /root/reference, so bench.py's
cpu_baseline / reference arm time the port; this script is the evidence that the two run at the same speed (and give the same
numbers: tests/test_oracle_golden.py).  Writes profiles/prior_reference_cpu_r02.json."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle import gen_golden as G
from oracle import ref_shims
from oracle.prior import OraclePrior
from oracle.synth import synth_state_dict

assert ref_shims.available(), "needs /root/reference"
torch.set_grad_enabled(False)
threads = os.cpu_count() or 1
torch.set_num_threads(threads)
src, clip_hidden = G.prior_inputs("l24_nodiff")
kw = dict(num_inference_steps=25, guidance_scale=10, score=6.5)
o = OraclePrior(n_layer=24).eval()
o.load_state_dict(synth_state_dict(o, seed=3))
ref, mod, fake = ref_shims.load_prior(n_layer=24)                      # prior/model.py imported unmodified
missing, unexpected = ref.load_state_dict(synth_state_dict({k: v.shape for k, v in ref.state_dict().items()}, seed=3), strict=False)
assert not missing and not unexpected
fake.hidden = clip_hidden
res = {}


def timed(fn, n=3):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.time()
        fn()
        ts.append(time.time() - t0)
    return sum(ts) / len(ts)


res["oracle_port_s"] = timed(lambda: o.generate_diffusion(3, 0, src, clip_hidden, **kw))
if ref is not None:
    res["reference_s"] = timed(lambda: ref.generate_diffusion(mod.MODALITY.VIDEO, mod.MODALITY.IMAGE, src, device="cpu", dtype=torch.float32,
                                                              image_bind_overwrite=None, do_classifier_free_guidance=True, **kw))
res.update(cores=threads, what="one complete 25-step CFG prior sampling of one sample, fp32, synthetic weights")
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "profiles", "prior_reference_cpu_r02.json"), "w"), indent=1)
