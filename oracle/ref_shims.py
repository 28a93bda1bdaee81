"""Oracle (test infrastructure): run the REFERENCE's own hot-path code in the build container.

Only usable where ``/root/reference`` exists (the build container; never the GPU box).
Nothing is copied: modules are imported / source fragments are exec'd from where they lie.

* ``load_attention_processors()`` -- imports diffusion/ip_adapter/attention_processor.py by path
  (torch-only file).
* ``load_prior()`` -- imports prior/model.py unmodified under (i) an empty parent package that bypasses
  ``instructany2pix/__init__.py`` (which needs imagebind/diffusers/gdino), (ii) a ``diffusers`` shim exporting
  the restated ``DDPMScheduler`` + ``get_timestep_embedding`` (third-party; oracle/schedulers.py), (iii) offline
  patches of the three hub calls (GPT2Config / tokenizer / CLIP text tower -- prior/model.py:33-34,187).
  The CLIP tower is replaced by a stub that returns a caller-provided hidden state for "".
* ``extract_source(path, name)`` -- AST-extracts one top-level def/class (``_backward_ddim``,
  ``ImageProjModel``) or method (``polar_intrtpolate``) and exec's it, because their files import
  diffusers at module level.
"""
from __future__ import annotations

import ast
import importlib
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("IA2P_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "instructany2pix"))


def load_attention_processors():
    p = os.path.join(REF_ROOT, "instructany2pix/diffusion/ip_adapter/attention_processor.py")
    spec = importlib.util.spec_from_file_location("_ref_attention_processor", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def extract_source(relpath: str, name: str, extra_globals=None):
    src = open(os.path.join(REF_ROOT, relpath)).read()
    tree = ast.parse(src)
    node = None
    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name == name:
            node = n
            break
    if node is None:
        raise KeyError(name)
    code = ast.get_source_segment(src, node)
    import textwrap
    g = {"torch": torch, "nn": nn}
    g.update(extra_globals or {})
    exec(textwrap.dedent(code), g)
    return g[name]


class _FakeClip(nn.Module):
    """Stands in for CLIPTextModelHiddenState (prior/model.py:20-105): same return contract."""
    hidden = None  # (1, T, 1024), set by the caller

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, batch):
        h = _FakeClip.hidden
        return [h.expand(len(batch), -1, -1).clone(), torch.ones(len(batch), h.shape[1])]


def load_prior(n_layer=24):
    """-> (InstructAny2PixPrior instance, module).  GPT-2 depth is patchable (golden fixtures use 2 and 24)."""
    from . import schedulers as S

    if "instructany2pix" not in sys.modules or not hasattr(sys.modules["instructany2pix"], "__ia2p_shim__"):
        pkg = types.ModuleType("instructany2pix")
        pkg.__path__ = [os.path.join(REF_ROOT, "instructany2pix")]
        pkg.__ia2p_shim__ = True
        sys.modules["instructany2pix"] = pkg
        dif = types.ModuleType("diffusers")

        class DDPMScheduler(S.DDPMSchedulerOracle):
            @classmethod
            def from_pretrained(cls, *a, **k):
                return cls()

        dif.DDPMScheduler = DDPMScheduler
        dm = types.ModuleType("diffusers.models")
        de = types.ModuleType("diffusers.models.embeddings")
        de.get_timestep_embedding = S.get_timestep_embedding
        dm.embeddings = de
        dif.models = dm
        sys.modules["diffusers"] = dif
        sys.modules["diffusers.models"] = dm
        sys.modules["diffusers.models.embeddings"] = de

    import transformers
    from transformers import GPT2Config

    orig = GPT2Config.from_pretrained
    GPT2Config.from_pretrained = classmethod(
        lambda cls, *a, **k: GPT2Config(n_embd=1024, n_layer=n_layer, n_head=16, n_positions=1024, vocab_size=50257))
    try:
        mod = importlib.import_module("instructany2pix.prior.model")
        mod.CLIPTextModelHiddenState = _FakeClip
        pkg_init = {}
        src = open(os.path.join(REF_ROOT, "instructany2pix/prior/__init__.py")).read()
        src = src.replace("from .model import InstructAny2PixPrior", "")
        exec(src, pkg_init)
        prior = mod.InstructAny2PixPrior(**pkg_init["prior_config"])
    finally:
        GPT2Config.from_pretrained = orig
    prior.device = torch.device("cpu")
    return prior.eval(), mod, _FakeClip


def load_ldm_modules():
    """-> (blocks, attention, util): the reference's in-tree LDM building blocks (llm/model/vae/modules/*.py, torch + einops
    only), imported unmodified under a private package name so that ``instructany2pix/llm/__init__.py`` (which needs the LLM
    stack) never runs.  The diffusers UNet / AutoencoderKL blocks the hot path uses are ports of these (SURVEY A.7 lists the
    differences), which makes them the only reference-run pin available for the restated ResnetBlock2D,
    BasicTransformerBlock, GEGLU FF, Transformer2D wrapper, VAE Encoder/Decoder, sinusoid and beta schedule."""
    name = "_ia2p_ref_ldm"
    if name not in sys.modules:
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REF_ROOT, "instructany2pix/llm/model/vae/modules")]
        sys.modules[name] = pkg
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        blocks = importlib.import_module(name + ".blocks")
        attention = importlib.import_module(name + ".attention")
        util = importlib.import_module(name + ".util")
    return blocks, attention, util
