"""Chunk engine: owns the device workspace for one GPU and drives the whole-chunk C-ABI entry point.

This is the host-side object the step classes, the sharded runner and bench.py share.  One engine per
process / GPU; buffers are sized once for the largest chunk seen and reused (no per-chunk allocation).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch

from . import _dev, _lib


class ChunkEngine:
    _shared: Optional["ChunkEngine"] = None

    def __init__(self):
        _dev.require_cuda()
        self._cap: Tuple[int, int, int, int, int] = (0, 0, 0, 0, 0)
        self._buf: Dict[str, torch.Tensor] = {}
        # the library-side state of the whole-chunk call (a side stream + two events on the current device), owned here
        self._handle = ctypes.c_void_p()
        _lib.call('msq_engine_create', ctypes.byref(self._handle))
        self._destroy = _lib.load().msq_engine_destroy

    def __del__(self):
        handle = getattr(self, '_handle', None)
        if handle:
            try:
                self._destroy(handle)
            except Exception:      # interpreter shutdown; pylint: disable=broad-except
                pass
            self._handle = None

    @classmethod
    def shared(cls) -> "ChunkEngine":
        if cls._shared is None:
            cls._shared = cls()
        return cls._shared

    # ---- workspace ---------------------------------------------------------------------------
    def _ensure(self, n: int, h: int, w: int, cw: int, ch: int, n_chunks: int) -> None:
        cap_n, cap_h, cap_w, cap_cw, cap_ch = self._cap
        if n <= cap_n and (h, w, cw, ch) == (cap_h, cap_w, cap_cw, cap_ch) and self._buf['filter_passes'].numel() >= n_chunks:
            return
        n_alloc = max(n, cap_n if (h, w, cw, ch) == (cap_h, cap_w, cap_cw, cap_ch) else 0)
        e = _dev.empty
        self._buf = {
            'cleaned': e((n_alloc, h, w), torch.uint8),
            'centroid': e((n_alloc, 2), torch.float64),
            'orientation': e((n_alloc,), torch.float64),
            'angle_deg': e((n_alloc,), torch.float64),
            'axis_length': e((n_alloc, 2), torch.float64),
            'flips': e((n_alloc,), torch.uint8),
            'flip_conf': e((n_alloc,), torch.float64),
            'scalars': e((_lib.NUM_SCALARS * n_alloc,), torch.float64),
            'kpt_cols': e((_lib.NUM_KPT_COLS * n_alloc,), torch.float64),
            'depth_crops': e((n_alloc, ch, cw), torch.uint8),
            'mask_crops': e((n_alloc, ch, cw), torch.uint8),
            'filter_passes': e((max(n_chunks, 64),), torch.int32),
            'scratch': e((int(_lib.load().msq_extract_scratch_bytes(n_alloc, h, w)) + 256,), torch.uint8),
        }
        self._cap = (n_alloc, h, w, cw, ch)

    # ---- full chunk ----------------------------------------------------------------------------
    def extract(self, chunk: torch.Tensor, masks: torch.Tensor, keypoints: torch.Tensor, *, chunk_size: int,
                min_height: float, max_height: float, true_depth: float, crop_size=(80, 80),
                positive_bits: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Run clean -> features -> angles/flips/filter -> scalars + keypoints -> crops on device tensors
        chunk (n,h,w) u8, masks (n,h,w) u8, keypoints (n,8,3) f32.  Returns VIEWS into the engine's workspace
        (valid until the next call); everything stays on the current stream, nothing synchronises.
        positive_bits: the bit rows `prep_raw_frames(..., positive_bits_out=)` produced for this chunk (optional)."""
        assert chunk.is_cuda and masks.is_cuda and keypoints.is_cuda
        if positive_bits is not None:
            assert positive_bits.is_cuda and positive_bits.is_contiguous() and \
                tuple(positive_bits.shape) == (chunk.shape[0], chunk.shape[1], (chunk.shape[2] + 31) // 32) and positive_bits.element_size() == 4
        n, h, w = (int(v) for v in chunk.shape)
        cw, ch = int(crop_size[0]), int(crop_size[1])
        n_chunks = (n + chunk_size - 1) // chunk_size
        self._ensure(n, h, w, cw, ch, n_chunks)
        b = self._buf
        outs = _lib.ChunkOutputs(*(b[k].data_ptr() for k in ('cleaned', 'centroid', 'angle_deg', 'axis_length', 'flips',
                                                             'scalars', 'kpt_cols', 'depth_crops', 'mask_crops',
                                                             'filter_passes')))
        _lib.call('msq_extract_chunk_engine', self._handle, _dev.ptr(chunk), _dev.ptr(positive_bits), _dev.ptr(masks), _dev.ptr(keypoints), n, h, w,
                  int(chunk_size),
                  float(min_height), float(max_height), float(true_depth), cw, ch, ctypes.byref(outs),
                  _dev.ptr(b['scratch']), b['scratch'].numel(), _dev.stream())
        return {
            'cleaned': b['cleaned'][:n], 'centroid': b['centroid'][:n], 'angle_deg': b['angle_deg'][:n],
            'axis_length': b['axis_length'][:n], 'flips': b['flips'][:n],
            'scalars': b['scalars'][:_lib.NUM_SCALARS * n].view(_lib.NUM_SCALARS, n),
            'kpt_cols': b['kpt_cols'][:_lib.NUM_KPT_COLS * n].view(_lib.NUM_KPT_COLS, n),
            'depth_crops': b['depth_crops'][:n], 'mask_crops': b['mask_crops'][:n],
            'filter_passes': b['filter_passes'][:n_chunks],
        }

    # ---- a6 + a7 only (the tracking branch post-processes the raw moment features itself) ----------
    def clean_and_features(self, chunk: torch.Tensor, masks: torch.Tensor) -> Dict[str, torch.Tensor]:
        n, h, w = (int(v) for v in chunk.shape)
        e = _dev.empty
        cleaned = torch.empty_like(chunk)
        centroid, orientation, axis = e((n, 2), torch.float64), e((n,), torch.float64), e((n, 2), torch.float64)
        flist = e((n + 1,), torch.int32)            # msq_frame_features scratch: frames left to the general kernel
        st = _dev.stream()
        _dev.clean_frames_ws(chunk, cleaned)
        _lib.call('msq_frame_features', _dev.ptr(cleaned), _dev.ptr(masks), n, h, w, 3.0, _dev.ptr(centroid),
                  _dev.ptr(orientation), _dev.ptr(axis), ctypes.c_void_p(0), _dev.ptr(flist), flist.numel() * 4, st)
        return {'cleaned': cleaned, 'centroid': centroid, 'orientation_rad': orientation, 'axis_length': axis}

    # ---- instances_to_features only --------------------------------------------------------------
    def features_only(self, chunk: torch.Tensor, masks: torch.Tensor, keypoints: torch.Tensor,
                      chunk_size: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """clean_frames + get_frame_features + flips + iterative filter (ref: proc/proc.py:700-848)."""
        n, h, w = (int(v) for v in chunk.shape)
        chunk_size = int(chunk_size or n)
        e = _dev.empty
        cleaned = torch.empty_like(chunk)
        centroid, orientation, axis = e((n, 2), torch.float64), e((n,), torch.float64), e((n, 2), torch.float64)
        angle, flips, conf = e((n,), torch.float64), e((n,), torch.uint8), e((n,), torch.float64)
        passes = e(((n + chunk_size - 1) // chunk_size,), torch.int32)
        flist = e((n + 1,), torch.int32)            # msq_frame_features scratch: frames left to the general kernel
        st = _dev.stream()
        _dev.clean_frames_ws(chunk, cleaned)
        _lib.call('msq_frame_features', _dev.ptr(cleaned), _dev.ptr(masks), n, h, w, 3.0, _dev.ptr(centroid),
                  _dev.ptr(orientation), _dev.ptr(axis), ctypes.c_void_p(0), _dev.ptr(flist), flist.numel() * 4, st)
        _lib.call('msq_angles_and_flips', _dev.ptr(orientation), _dev.ptr(axis), _dev.ptr(centroid), _dev.ptr(keypoints),
                  n, chunk_size, _dev.ptr(angle), _dev.ptr(flips), _dev.ptr(conf), _dev.ptr(passes), st)
        return {'cleaned': cleaned, 'centroid': centroid, 'orientation_rad': orientation, 'angle_deg': angle,
                'axis_length': axis, 'flips': flips, 'flip_conf': conf, 'filter_passes': passes}
