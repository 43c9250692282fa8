"""Helpers for the -m gpu parity tests: call the C ABI with torch-owned device memory."""
import ctypes as C

import torch

from hedit_b200 import _lib


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def lib():
    return _lib.load()


def bf(t):
    return t.to(torch.bfloat16).contiguous()


def sync_check(rc, what):
    _lib.check(rc, what)
    torch.cuda.synchronize()


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), (a - b).abs().max().item()
