"""Host side of csrc/conv_tc.cu: route a convolution / Linear of the R-CNN graph to the tcgen05 implicit-GEMM kernel when its
shape is one that kernel serves (bf16, channels-last, channel counts multiples of 64, 1x1 stride 1|2 or 3x3 stride 1 pad 1);
return None otherwise so that model/ops.py falls through to the cuDNN / cuBLAS call."""
from __future__ import annotations

from typing import Optional

import torch

from .. import _dev, _lib

def _bias(b: Optional[torch.Tensor]):
    """(pointer, is_bf16) of a bias vector the kernel can read as it is (float32 or bf16), or None when it cannot."""
    if b is None:
        return _dev.ptr(None), 0
    if b.dtype not in (torch.float32, torch.bfloat16) or not b.is_contiguous() or b.data_ptr() % 16:
        return None
    return _dev.ptr(b), int(b.dtype == torch.bfloat16)


def try_conv2d(x, w, b, z, relu, stride, pad) -> Optional[torch.Tensor]:
    k = int(w.shape[2])
    cout, cin = int(w.shape[0]), int(w.shape[1])
    if not (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.dim() == 4 and int(w.shape[3]) == k):
        return None
    if not ((k == 1 and pad == 0 and stride in (1, 2)) or (k == 3 and pad == 1 and stride == 1)) or cin % 64 or cout % 64:
        return None
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    if not w.is_contiguous(memory_format=torch.channels_last):
        w = w.contiguous(memory_format=torch.channels_last)
    n, _, h, wd = (int(v) for v in x.shape)
    ho, wo = (h - 1) // stride + 1, (wd - 1) // stride + 1
    out = torch.empty((n, cout, ho, wo), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    if z is not None:
        if z.dtype != torch.bfloat16 or tuple(z.shape) != tuple(out.shape):
            return None
        if not z.is_contiguous(memory_format=torch.channels_last):
            z = z.contiguous(memory_format=torch.channels_last)
    bias = _bias(b)
    if bias is None:
        return None
    _lib.call('msq_conv_tc', _dev.ptr(x), n, h, wd, cin, _dev.ptr(w), cout, k, int(stride), bias[0], bias[1], _dev.ptr(z), int(bool(relu)),
              _dev.ptr(out), _dev.stream())
    return out


def try_linear(x, w, b, relu) -> Optional[torch.Tensor]:
    if not (x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.dim() == 2 and x.is_contiguous() and w.is_contiguous()):
        return None
    rows, kdim = int(x.shape[0]), int(x.shape[1])
    nout = int(w.shape[0])
    if kdim % 64 or nout % 64 or rows == 0:
        return None
    out = torch.empty((rows, nout), dtype=torch.bfloat16, device=x.device)
    bias = _bias(b)
    if bias is None:
        return None
    _lib.call('msq_conv_tc', _dev.ptr(x), 1, 1, rows, kdim, _dev.ptr(w), nout, 1, 1, bias[0], bias[1], None, int(bool(relu)), _dev.ptr(out),
              _dev.stream())
    return out
