"""CPU oracle for the InstructAny2Pix denoising hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker (or the
timed CPU baseline), never as the thing shipped.  The product package
``instructany2pix_b200`` must not import this package.

What it is: a plain PyTorch fp32 restatement of the arithmetic the reference
executes on the hot path.  The reference (``/root/reference``) is pure Python on
top of two pinned third-party packages that are NOT vendored there and NOT
installed in this image:

* ``diffusers==0.26.3`` (requirements.txt:3) -- ``UNet2DConditionModel``,
  ``DDIMScheduler``, ``DDPMScheduler``, ``get_timestep_embedding`` and the SDXL
  sampler loop.  Restated in :mod:`oracle.unet`, :mod:`oracle.schedulers`,
  :mod:`oracle.sampler` from the published algorithm; anchored on the
  reference's own call sites (pipeline.py:101-116,307; ddim/pnp_pipeline.py:
  133,251-275; diffusion/ip_adapter/custom_pipelines.py:250,324-363).
* ``transformers==4.34.1`` (requirements.txt:4) -- ``GPT2Model`` (prior trunk).
  Restated in :mod:`oracle.prior`; validated here against the installed
  transformers 5.5 ``GPT2Model`` and against the reference's own
  ``prior/model.py`` executed under shims (``oracle/ref_shims.py``).

The reference's own hot-path code (attention processors, ImageProjModel,
``_backward_ddim``, ``polar_intrtpolate``, the prior wrapper) IS importable in
the build container; ``oracle/gen_golden.py`` runs it there and commits small
input/output fixtures under ``tests/golden/``.  The restatements in this
package are pinned against those fixtures (tests/test_oracle_golden.py).

Parity status: the reference ships no tests/golden vectors for this path
(SURVEY.md section 4).  Pinned against outputs of reference code RUN in the
build container: the in-tree parts (processors, projector, inversion step,
polar blend, prior wrapper) and -- through the in-tree LDM modules the
diffusers blocks are ports of (llm/model/vae/modules; oracle/ldm_map.py) --
ResnetBlock2D, GEGLU FF, BasicTransformerBlock, the Transformer2D wrapper, the
VAE encoder/decoder trunks, the sinusoid, the alpha-bar table, the DDIM
timestep tables, and DDIMScheduler.step (as the inverse of the reference's
``_backward_ddim``).  Still "parity unpinned" (restated from the published
diffusers 0.26.3 semantics, pinned only structurally): the assembly of those
blocks into the SDXL UNet, the Euler scheduler, the 1x1 (post_)quant convs.
"""
