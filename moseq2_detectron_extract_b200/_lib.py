"""ctypes binding of libmoseq_b200.so (C ABI declared in include/moseq_b200.h).

There is deliberately NO fallback: if the library cannot be loaded, or a call fails, an exception is
raised.  Nothing in this package computes the hot path on the CPU.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libmoseq_b200.so')

MSQ_BG_NONE, MSQ_BG_F32, MSQ_BG_F64, MSQ_BG_U16 = 0, 1, 2, 3
MSQ_PREP_HAS_VMIN, MSQ_PREP_HAS_VMAX = 1, 2
NUM_SCALARS, NUM_KPT_COLS, NUM_KEYPOINTS = 17, 96, 8


class MoseqB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


class ChunkOutputs(ctypes.Structure):
    """struct msq_chunk_outputs"""
    _fields_ = [(name, c_void_p) for name in (
        'cleaned', 'centroid', 'angle_deg', 'axis_length', 'flips', 'scalars', 'kpt_cols', 'depth_crops',
        'mask_crops', 'filter_passes')]


# name -> (restype, argtypes); every symbol include/moseq_b200.h declares
SIGNATURES = {
    'msq_version': (c_int, []),
    'msq_last_error': (c_char_p, []),
    'msq_device_info': (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int), c_char_p, c_int]),
    'msq_kernel_timing_enable': (c_int, [c_int]),
    'msq_kernel_timing_collect': (c_int, [POINTER(c_double), POINTER(ctypes.c_longlong), c_int]),
    'msq_kernel_count': (c_int, []),
    'msq_kernel_name': (c_char_p, [c_int]),
    'msq_kernel_launches': (ctypes.c_longlong, [c_int]),
    'msq_prep_frames': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                c_double, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'msq_positive_bits_bytes': (c_size_t, [c_int, c_int, c_int]),
    'msq_prep_frames_bits': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                     c_double, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'msq_unpack_mask_bits': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'msq_copy_roi_rows': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'msq_copy_roi_bands': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                   c_void_p]),
    'msq_inpaint_scratch_bytes': (c_size_t, [c_int, c_int, c_int]),
    'msq_inpaint_frames': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    'msq_scale_frames': (c_int, [c_void_p, c_void_p, c_size_t, c_double, c_double, c_int, c_void_p]),
    'msq_scale_frames_chw3_f32': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double, c_int, c_void_p]),
    'msq_clean_frames': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'msq_clean_scratch_bytes': (c_size_t, [c_int, c_int, c_int]),
    'msq_clean_frames_ws': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    'msq_frame_features_scratch_bytes': (c_size_t, [c_int, c_int, c_int]),
    'msq_frame_features': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    'msq_paste_masks': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    'msq_nms_sorted': (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p]),
    'msq_stem_conv_pool': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_double, c_double, c_int, c_float, c_float, c_void_p, c_void_p,
                                   c_void_p, c_int, c_void_p]),
    'msq_stem_conv_pool_tc': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_double, c_double, c_int, c_float, c_float, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    'msq_nms_scratch_bytes': (c_size_t, [c_int, c_int]),
    'msq_nms_levels_scratch_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'msq_nms_levels_long': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, ctypes.c_float, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                    c_void_p]),
    'msq_nms_sorted_long': (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'msq_rpn_select': (c_int, [POINTER(c_void_p), POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_float), c_int, c_int, c_int, c_int,
                               c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'msq_keypoints_from_heatmaps': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'msq_angles_and_flips': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    'msq_flips_from_keypoints': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    'msq_iterative_filter_angles': (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_int, c_void_p, c_void_p,
                                            c_void_p, c_void_p]),
    'msq_scalars_scratch_bytes': (c_size_t, [c_int]),
    'msq_scalars_and_keypoints': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int, c_int, c_int, c_int, c_double, c_double, c_double, c_void_p,
                                          c_void_p, c_void_p, c_size_t, c_void_p]),
    'msq_scalar_name': (c_char_p, [c_int]),
    'msq_keypoint_col_name': (c_char_p, [c_int]),
    'msq_crop_scratch_bytes': (c_size_t, [c_int]),
    'msq_crop_rotate': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_void_p]),
    'msq_kalman_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'msq_kalman_smooth': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                  c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'msq_kalman_em': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                              c_int, c_void_p, c_size_t, c_void_p]),
    'msq_keypoint_alignment_scores': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    'msq_tracking_prepare': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    'msq_flips_from_keypoints_f64': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    'msq_scalars_and_keypoints_f64': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_int, c_int, c_int, c_int, c_double, c_double, c_double, c_void_p,
                                              c_void_p, c_void_p, c_size_t, c_void_p]),
    'msq_track_angles': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                 c_void_p, c_int, c_void_p]),
    'msq_bground_scratch_bytes': (c_size_t, [c_int, c_int, c_int]),
    'msq_get_bground_im': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'msq_detector_input': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float), c_double, c_double, c_int, c_void_p]),
    'msq_roi_align_levels': (c_int, [POINTER(c_void_p), POINTER(c_int), POINTER(c_int), POINTER(c_float), c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'msq_conv_tc': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'msq_group_norm_scratch_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'msq_group_norm_nhwc': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_float,
                                    c_void_p, c_void_p, c_size_t, c_void_p]),
    'msq_roi_align_v2': (c_int, [POINTER(c_void_p), POINTER(c_int), POINTER(c_int), POINTER(c_float), c_int, c_int, c_int, c_void_p, c_int,
                                 c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    'msq_fastrcnn_top1': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, POINTER(c_float), c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    'msq_upsample2x_bilinear': (c_int, [c_void_p, c_int, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int,
                                        c_int, c_int, c_void_p, c_void_p]),
    'msq_keypoints_from_heatmaps_d2': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'msq_sobel_gradient_mask': (c_int, [c_void_p, c_int, c_int, POINTER(c_double), c_int, POINTER(c_double), c_int, c_double, c_void_p, c_void_p]),
    'msq_plane_ransac_score': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    'msq_plane_distance': (c_int, [c_void_p, c_int, c_int, ctypes.POINTER(c_double), c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    'msq_label_scratch_bytes': (c_size_t, [c_int, c_int]),
    'msq_label_regions': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'msq_region_props': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'msq_region_rois': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'msq_extract_scratch_bytes': (c_size_t, [c_int, c_int, c_int]),
    'msq_engine_create': (c_int, [POINTER(c_void_p)]),
    'msq_engine_destroy': (c_int, [c_void_p]),
    'msq_extract_chunk_engine': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_double,
                                         c_double, c_int, c_int, POINTER(ChunkOutputs), c_void_p, c_size_t, c_void_p]),
    'msq_extract_chunk': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_double,
                                  c_double, c_int, c_int, POINTER(ChunkOutputs), c_void_p, c_size_t, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built -- run
    `python -m moseq2_detectron_extract_b200.build` or `__graft_entry__.build()`."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MoseqB200Error(f'{LIB_PATH} is missing: the CUDA extension has not been built '
                             '(python -m moseq2_detectron_extract_b200.build). There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().msq_last_error()
    return msg.decode('utf-8', 'replace') if msg else ''


def check(status: int, what: str) -> None:
    if status != 0:
        raise MoseqB200Error(f'{what} failed with status {status}: {last_error()}')


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise on failure."""
    check(getattr(load(), name)(*args), name)


def scalar_names():
    lib = load()
    return [lib.msq_scalar_name(i).decode() for i in range(NUM_SCALARS)]


def keypoint_col_names():
    lib = load()
    return [lib.msq_keypoint_col_name(i).decode() for i in range(NUM_KPT_COLS)]


def kernel_names():
    lib = load()
    return [lib.msq_kernel_name(i).decode() for i in range(lib.msq_kernel_count())]


def kernel_launches() -> dict:
    """Kernel launches issued by the library since it was loaded, per kernel."""
    lib = load()
    return {name: int(lib.msq_kernel_launches(i)) for i, name in enumerate(kernel_names())}


def kernel_timing(enable: bool) -> None:
    check(load().msq_kernel_timing_enable(int(enable)), 'msq_kernel_timing_enable')


def kernel_timing_collect() -> dict:
    """{kernel: (total_ms, launches_timed)} accumulated since the previous collect."""
    lib = load()
    k = lib.msq_kernel_count()
    ms = (c_double * k)()
    cnt = (ctypes.c_longlong * k)()
    check(lib.msq_kernel_timing_collect(ms, cnt, k), 'msq_kernel_timing_collect')
    return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(kernel_names())}
