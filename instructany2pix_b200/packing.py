"""Weight re-layout from diffusers / HF state-dict tensors to the kernel layouts (done once at load time).

Pure tensor reshapes (device-agnostic torch indexing, no arithmetic on the hot path).
"""
from __future__ import annotations

import torch


def pack_conv3x3(w, w_shortcut=None, dtype=torch.bfloat16):
    """Conv2d weight [Cout,Cin,3,3] (+ optional 1x1 shortcut [Cout,Csc,1,1]) -> [Cout, 9*Cin (+Csc)],
    K order (ky, kx, cin) then shortcut channels: the tap order of ia2p_conv3x3_nhwc_bf16."""
    cout = w.shape[0]
    p = w.permute(0, 2, 3, 1).reshape(cout, -1)
    if w_shortcut is not None:
        p = torch.cat([p, w_shortcut.reshape(cout, -1)], dim=1)
    return p.to(dtype).contiguous()


def pack_conv3x3_up2x(w, dtype=torch.bfloat16):
    """Conv2d weight [Cout,Cin,3,3] of a conv that follows a nearest-2x upsample -> [4, Cout, 4*Cin]: for output parity
    (py, px) the 3x3 taps collapse onto a 2x2 neighbourhood of the low-resolution map (rows {y-1, y} for py = 0, {y, y+1} for
    py = 1; same for columns), with the weights of coinciding taps summed in fp32.  Tap order (row tap, col tap), channels inner:
    the layout ia2p_conv_up2x_nhwc_bf16 expects."""
    cout, cin = w.shape[:2]
    wf = w.detach().float()
    sets = {0: ([0], [1, 2]), 1: ([0, 1], [2])}            # parity -> source kernel indices of (first tap, second tap)
    out = []
    for py in range(2):
        for px in range(2):
            taps = []
            for i in range(2):
                for j in range(2):
                    acc = torch.zeros(cout, cin, device=w.device)
                    for ky in sets[py][i]:
                        for kx in sets[px][j]:
                            acc = acc + wf[:, :, ky, kx]
                    taps.append(acc)
            out.append(torch.cat(taps, dim=1))
    return torch.stack(out, 0).to(dtype).contiguous()


def pack_conv_out(w, b, pad_to=32, dtype=torch.bfloat16):
    """few-output-channel Conv2d [cout <= 8, Cin, 3, 3] -> ([pad_to, 9*Cin] bf16 with zero rows, [pad_to] fp32 bias): conv_out as a
    tensor-core conv whose extra output channels are discarded (ops.conv_out_tc)."""
    cout = w.shape[0]
    wp = torch.zeros(pad_to, w.shape[1] * 9, device=w.device, dtype=dtype)
    wp[:cout] = pack_conv3x3(w, dtype=dtype)
    bp = torch.zeros(pad_to, device=w.device, dtype=torch.float32)
    if b is not None:
        bp[:cout] = b.detach().float()
    return wp.contiguous(), bp


def interleave_geglu(w, b=None, group=32):
    """GEGLU proj weight [8C, C] (rows [0,4C) value | [4C,8C) gate) -> rows interleaved in `group`-row
    blocks [value_0 | gate_0 | value_1 | gate_1 ...] so value/gate of one output land in the same accumulator tile."""
    n2 = w.shape[0]
    half = n2 // 2
    assert half % group == 0
    idx = torch.arange(n2, device=w.device).reshape(2, half // group, group).permute(1, 0, 2).reshape(-1)
    wi = w[idx].contiguous()
    bi = None if b is None else b[idx].contiguous()
    return wi, bi


def conv1d_to_linear(w):
    """HF Conv1D weight [in, out] -> nn.Linear layout [out, in] (K-major rows)."""
    return w.t().contiguous()
