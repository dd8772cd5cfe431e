"""Device-memory plumbing: PyTorch owns device tensors and streams, nothing else.

Every public function of this package accepts numpy arrays (host; copied to the current CUDA device,
results copied back as numpy) or CUDA torch tensors (used in place; results stay on the device).
"""
from __future__ import annotations

import ctypes
from typing import Any, Optional

import numpy as np
import torch

from . import _lib


class CudaRequiredError(RuntimeError):
    """Raised instead of silently computing on the CPU."""


def require_cuda() -> None:
    _lib.load()                                           # raises if the extension is not built
    if not torch.cuda.is_available():
        raise CudaRequiredError('moseq2_detectron_extract_b200 needs a CUDA device (B200, sm_100a); '
                                'there is no CPU fallback for the extract hot path.')


def is_device_tensor(x: Any) -> bool:
    return isinstance(x, torch.Tensor) and x.is_cuda


def as_device(x: Any, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Contiguous CUDA tensor of `dtype` (no copy when `x` already is one)."""
    require_cuda()
    if isinstance(x, torch.Tensor):
        t = x if x.is_cuda else x.cuda(non_blocking=True)
    else:
        arr = np.ascontiguousarray(x)
        if arr.dtype == np.uint16:                        # torch has limited uint16 support: move the bits
            t = torch.from_numpy(arr.view(np.int16)).cuda(non_blocking=True)
            return t if dtype is None or dtype == torch.int16 else t.to(torch.int32).bitwise_and_(0xFFFF).to(dtype)
        t = torch.from_numpy(arr).cuda(non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def empty(shape, dtype) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, device='cuda')


def ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def give_back(t: torch.Tensor, like: Any):
    """numpy in -> numpy out; torch tensor in (CUDA, or host/pinned) -> the result stays on the device."""
    if isinstance(like, torch.Tensor):
        return t
    return t.cpu().numpy()


def positive_bits_like(chunk: torch.Tensor) -> torch.Tensor:
    """Room for the positive-pixel bit rows of a (n,h,w) u8 chunk: (n, h, ceil(w/32)) int32 on its device."""
    n, h, w = (int(v) for v in chunk.shape)
    return torch.empty((n, h, (w + 31) // 32), dtype=torch.int32, device=chunk.device)


def clean_frames_ws(src: torch.Tensor, out: torch.Tensor, positive_bits: Optional[torch.Tensor] = None) -> None:
    """`msq_clean_frames_ws` with a torch-owned scratch buffer (the row pre-pass runs as its own launch)."""
    n, h, w = (int(v) for v in src.shape)
    nbytes = int(_lib.load().msq_clean_scratch_bytes(n, h, w))
    scratch = torch.empty((max(nbytes, 8),), dtype=torch.uint8, device=src.device)
    _lib.call('msq_clean_frames_ws', ptr(src), ptr(positive_bits), ptr(out), n, h, w, ptr(scratch), nbytes, stream())


def roi_bands(roi_box: np.ndarray, n_bands: int = 16, align: int = 8):
    """Horizontal bands that cover a ROI inside its bounding box, for `msq_copy_roi_bands`: (band_y (n+1), band_x0 (n), band_x1 (n))
    int32 host arrays; columns rounded outwards to `align` pixels (16-byte DMA rows for int16 frames)."""
    roi_box = np.asarray(roi_box) > 0
    h, w = roi_box.shape
    n_bands = max(1, min(int(n_bands), h))
    edges = np.linspace(0, h, n_bands + 1).astype(np.int32)
    x0s, x1s = np.zeros(n_bands, np.int32), np.zeros(n_bands, np.int32)
    for b in range(n_bands):
        cols = np.flatnonzero(roi_box[edges[b]:edges[b + 1]].any(axis=0))
        if cols.size:
            x0s[b] = cols.min() // align * align
            x1s[b] = min(w, (cols.max() + align) // align * align)
    return np.ascontiguousarray(edges), x0s, x1s


def copy_roi_bands(frames_host: torch.Tensor, y0: int, x0: int, bands, out: torch.Tensor) -> None:
    """`msq_copy_roi_bands`: pinned (n,H,W) int16 host frames -> the ROI pixels of the dense (n,h,w) int16 device array `out`."""
    n, H, W = (int(v) for v in frames_host.shape)
    _, h, w = (int(v) for v in out.shape)
    by, bx0, bx1 = bands
    _lib.call('msq_copy_roi_bands', ptr(frames_host), n, H, W, int(y0), int(x0), h, w, by.ctypes.data_as(ctypes.c_void_p),
              bx0.ctypes.data_as(ctypes.c_void_p), bx1.ctypes.data_as(ctypes.c_void_p), len(bx0), ptr(out), stream())
