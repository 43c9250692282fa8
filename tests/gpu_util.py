"""Helpers for the -m gpu parity tests: call the C ABI with torch-owned device memory."""
import ctypes as C

import torch

from hedit_b200 import _lib


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def lib():
    return _lib.load()


def bf(t):
    """Round to the library's 16-bit operand dtype (fp16 by default, bf16 with -DHEDIT_OPERAND_BF16)."""
    return t.to(_lib.operand_torch_dtype()).contiguous()


def opdtype():
    return _lib.operand_torch_dtype()


def sync_check(rc, what):
    _lib.check(rc, what)
    torch.cuda.synchronize()


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), (a - b).abs().max().item()
