"""Weight re-layouts (instructany2pix_b200/packing.py) against the plain formulas they must preserve (CPU, fp32 where exactness matters)."""
import pytest
import torch
import torch.nn.functional as F

from instructany2pix_b200.packing import conv1d_to_linear, interleave_geglu, pack_conv3x3, pack_conv3x3_up2x, pack_conv_out
from tests import emu_ops

torch.set_grad_enabled(False)


def _q(t):            # bf16-representable values: packing itself then adds no rounding
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("H,W", [(4, 4), (6, 10), (5, 3)])
def test_up2x_parity_weights_equal_upsample_then_conv(H, W):
    """nearest-2x upsample + 3x3 conv == four parity 2x2 convs over the low-res map with summed taps (exact in fp32)"""
    g = torch.Generator().manual_seed(0)
    cin, cout = 8, 6
    x, w = torch.randn(2, H, W, cin, generator=g), torch.randn(cout, cin, 3, 3, generator=g)
    w4 = pack_conv3x3_up2x(w, dtype=torch.float32)
    assert w4.shape == (4, cout, 4 * cin)
    up = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, w, padding=1).permute(0, 2, 3, 1)
    out = torch.zeros_like(ref)
    xp = F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1))
    for py in range(2):
        for px in range(2):
            wk = w4[py * 2 + px].reshape(cout, 2, 2, cin).permute(0, 3, 1, 2)
            win = xp[:, :, py:py + H + 1, px:px + W + 1]
            out[:, py::2, px::2, :] = F.conv2d(win, wk).permute(0, 2, 3, 1)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)


def test_conv3x3_k_order_and_fused_shortcut():
    """K order (ky, kx, cin) then the 1x1 shortcut channels: one GEMM over [im2col | raw input] == conv + shortcut conv"""
    g = torch.Generator().manual_seed(1)
    cin, csc, cout = 4, 5, 3
    x, s = torch.randn(1, 5, 6, cin, generator=g), torch.randn(1, 5, 6, csc, generator=g)
    w, ws = torch.randn(cout, cin, 3, 3, generator=g), torch.randn(cout, csc, 1, 1, generator=g)
    p = pack_conv3x3(w, ws, dtype=torch.float32)
    assert p.shape == (cout, 9 * cin + csc)
    cols = F.unfold(x.permute(0, 3, 1, 2), 3, padding=1)                         # [1, cin*9, HW], order (cin, ky, kx)
    cols = cols.reshape(1, cin, 9, -1).permute(0, 2, 1, 3).reshape(1, 9 * cin, -1)   # -> (ky, kx, cin)
    a = torch.cat([cols[0].t(), s.reshape(-1, csc)], dim=1)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, padding=1) + F.conv2d(s.permute(0, 3, 1, 2), ws)
    torch.testing.assert_close((a @ p.t()).reshape(5, 6, cout), ref[0].permute(1, 2, 0), rtol=1e-5, atol=1e-5)


def test_geglu_interleave_is_a_pairing_permutation():
    n, c, group = 256, 16, 32
    w, b = torch.arange(n * c, dtype=torch.float32).reshape(n, c), torch.arange(n, dtype=torch.float32)
    wi, bi = interleave_geglu(w, b, group)
    half = n // 2
    for blk in range(half // group):
        v = slice(2 * blk * group, (2 * blk + 1) * group)
        gt = slice((2 * blk + 1) * group, (2 * blk + 2) * group)
        assert torch.equal(wi[v], w[blk * group:(blk + 1) * group])              # value rows of block blk
        assert torch.equal(wi[gt], w[half + blk * group: half + (blk + 1) * group])   # their gate rows right behind
        assert torch.equal(bi[v], b[blk * group:(blk + 1) * group]) and torch.equal(bi[gt], b[half + blk * group: half + (blk + 1) * group])
    assert sorted(bi.tolist()) == b.tolist()


def test_conv_out_padding_and_conv1d():
    g = torch.Generator().manual_seed(2)
    w, b = _q(torch.randn(4, 8, 3, 3, generator=g)), torch.randn(4, generator=g)
    wp, bp = pack_conv_out(w, b)
    assert wp.shape == (32, 72) and bp.shape == (32,) and wp.dtype == torch.bfloat16
    assert torch.equal(wp[:4].float(), pack_conv3x3(w, dtype=torch.float32)) and not wp[4:].any() and not bp[4:].any()
    assert torch.equal(bp[:4], b)
    x = _q(torch.randn(1, 4, 4, 8, generator=g)).to(torch.bfloat16)
    y = emu_ops.conv_out_tc(x, wp, bp, 4)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=1)
    torch.testing.assert_close(y.reshape(ref.shape) if y.shape != ref.shape else y, ref, rtol=1e-4, atol=1e-4)
    c1 = torch.randn(6, 10, generator=g)
    assert torch.equal(conv1d_to_linear(c1), c1.t()) and conv1d_to_linear(c1).is_contiguous()
