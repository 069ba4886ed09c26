"""Device guard for the C-ABI calls.

libmavlm.so launches on the CUDA runtime's CURRENT device (cudaGetDevice: SM count, function attributes, kernels)
and the host passes `torch.cuda.current_stream()`.  Tensors may live on another device (`model.to('cuda:1')` or a
device_map without `torch.cuda.set_device`): without a guard the kernels would run on the wrong GPU and stream --
an illegal address without peer access, silently unordered work with it.  `on_tensor_device` runs the wrapped call
under `torch.cuda.device(<the operands' device>)`, so both the stream lookup and every launch inside use that
device, and refuses operands that sit on different devices.
"""
from __future__ import annotations

import functools

import torch


def tensor_device(args, kwargs):
    """The one CUDA device of all tensor operands (lists / tuples of tensors are looked into), or None."""
    dev = None

    def visit(a):
        nonlocal dev
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise RuntimeError(f"mavlm: operands are on different CUDA devices ({dev} and {a.device})")
        elif isinstance(a, (list, tuple)):
            for e in a:
                if isinstance(e, torch.Tensor):
                    visit(e)

    for a in args:
        visit(a)
    for a in kwargs.values():
        visit(a)
    return dev


def on_tensor_device(fn):
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = tensor_device(args, kwargs)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped
