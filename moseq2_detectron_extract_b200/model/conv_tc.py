"""Host side of csrc/conv_tc.cu: route a convolution / Linear of the R-CNN graph to the tcgen05 implicit-GEMM kernel when its
shape is one that kernel serves (bf16, channels-last, channel counts multiples of 64, 1x1 stride 1|2 or 3x3 stride 1 pad 1);
return None otherwise so that model/ops.py falls through to the cuDNN / cuBLAS call."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .. import _dev, _lib

_BIAS_F32: Dict[int, torch.Tensor] = {}


def _bias_f32(b: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """The kernel adds the bias in float32; the graph keeps it in the compute dtype for cuDNN.  Converted once per tensor."""
    if b is None:
        return None
    key = b.data_ptr()
    hit = _BIAS_F32.get(key)
    if hit is None or hit.numel() != b.numel():
        hit = b.detach().float().contiguous()
        _BIAS_F32[key] = hit
    return hit


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
    _lib.call('msq_conv_tc', _dev.ptr(x), n, h, wd, cin, _dev.ptr(w), cout, k, int(stride), _dev.ptr(_bias_f32(b)), _dev.ptr(z), int(bool(relu)),
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
    _lib.call('msq_conv_tc', _dev.ptr(x), 1, 1, rows, kdim, _dev.ptr(w), nout, 1, 1, _dev.ptr(_bias_f32(b)), None, int(bool(relu)), _dev.ptr(out),
              _dev.stream())
    return out
