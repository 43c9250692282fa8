"""Calibration helper for the loop-level parity bounds ("derive the bound by running the oracle itself in
16-bit-operand / fp32-accumulate emulation").

`operand_rounding(dtype)` makes every contraction of a torch module tree -- `F.linear`, `F.conv2d`, `torch.bmm`,
`torch.baddbmm` (all the oracle UNet uses, oracle/sd_unet.py) -- round BOTH operands to `dtype` and then compute in
fp32, which is the arithmetic contract of a tcgen05 `kind::f16` tile (16-bit operands, fp32 accumulate).  Everything else
(normalisation statistics, softmax, residual adds, scheduler algebra) stays fp32, as in the CUDA path.  The emulation
is not bit-identical to the CUDA kernels (different summation order, fused epilogues, fp16 storage of q/k/v) -- it is a
second, independent realisation of the same rounding-noise process, so the deviation it produces from the fp32 run
is the yardstick the CUDA path's deviation is compared with.
"""
import contextlib

import torch
import torch.nn.functional as F


def _r(t, dtype):
    return t.to(dtype).to(torch.float32) if (t is not None and torch.is_floating_point(t)) else t


@contextlib.contextmanager
def operand_rounding(dtype=torch.float16):
    lin, conv, bmm, baddbmm = F.linear, F.conv2d, torch.bmm, torch.baddbmm

    def linear(x, w, b=None):
        return lin(_r(x, dtype), _r(w, dtype), b)

    def conv2d(x, w, b=None, *a, **k):
        return conv(_r(x, dtype), _r(w, dtype), b, *a, **k)

    def bmm_(a, b, **k):
        return bmm(_r(a, dtype), _r(b, dtype), **k)

    def baddbmm_(c, a, b, **k):
        return baddbmm(c, _r(a, dtype), _r(b, dtype), **k)

    F.linear, F.conv2d, torch.bmm, torch.baddbmm = linear, conv2d, bmm_, baddbmm_
    try:
        yield
    finally:
        F.linear, F.conv2d, torch.bmm, torch.baddbmm = lin, conv, bmm, baddbmm
