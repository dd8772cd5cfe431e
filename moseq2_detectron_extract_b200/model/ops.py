"""`torch.ops.msq.*`: the operators the repo's own R-CNN graph (model/rcnn.py) is made of.

They are registered with the PyTorch dispatcher (torch.library) so that a TorchScript export of the graph can name them;
every implementation is CUDA-only and goes through the C ABI of libmoseq_b200.so (ctypes) or a cuDNN / cuBLAS library call
for the dense contractions.  There is no CPU implementation: on a CPU tensor the dispatcher raises.

Importing this module registers the operators; `Predictor.from_torchscript` imports it before `torch.jit.load`.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict

import torch
import torch.nn.functional as F

from .. import _dev, _lib

_LIB = torch.library.Library('msq', 'DEF')
_LIB.define('detector_input(Tensor chunk_u8, float vmin, float vmax, bool int_limits, float[] mean, float[] std, int ph, int pw, '
            'bool bf16) -> Tensor')
_LIB.define('stem_conv_pool(Tensor chunk_u8, float vmin, float vmax, bool int_limits, float mean, float std, int ph, int pw, '
            'Tensor w49x64, Tensor bias64, bool bf16) -> Tensor')
_LIB.define('stem_conv_pool_tc(Tensor chunk_u8, float vmin, float vmax, bool int_limits, float mean, float std, int ph, int pw, '
            'Tensor b_tile, Tensor bias64) -> Tensor')
_LIB.define('conv2d(Tensor x, Tensor w, Tensor? b, Tensor? z, bool relu, int stride, int pad) -> Tensor')
_LIB.define('linear(Tensor x, Tensor w, Tensor? b, bool relu) -> Tensor')
_LIB.define('group_norm_nhwc(Tensor x, Tensor gamma, Tensor beta, int groups, float eps, Tensor? top, float scale) -> Tensor')
_LIB.define('rpn_proposals(Tensor[] preds, int[] strides, float[] sizes, float[] ratios, int img_h, int img_w, int pre_topk, '
            'int post_topk, float nms_thresh) -> (Tensor, Tensor, Tensor)')
_LIB.define('roi_align_v2(Tensor[] feats, float[] scales, Tensor boxes, int rois_per_image, int pooled, int sampling_ratio, '
            'int min_level, int canonical_level, float canonical_size) -> Tensor')
_LIB.define('fastrcnn_top1(Tensor pred, Tensor proposals, Tensor counts, int img_h, int img_w, float score_thresh, float[] weights) '
            '-> (Tensor, Tensor, Tensor)')
_LIB.define('keypoints_from_heatmaps_d2(Tensor heatmaps, Tensor boxes) -> Tensor')
_LIB.define('upsample2x_bilinear(Tensor x) -> Tensor')

# implementation switches (bench / tests): which engine runs the dense contractions
RPN_ENGINE = {'mode': 'fused'}            # 'fused' (msq_rpn_select) | 'torch' (operator by operator)
CONV_ENGINE = {'mode': 'auto'}           # 'auto' (per layer shape, the faster of the two) | 'cudnn' | 'tcgen05' (csrc/conv_tc.cu)


def _is_cl(x: torch.Tensor) -> bool:
    return x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)


def _cl(x: torch.Tensor) -> torch.Tensor:
    return x if _is_cl(x) else x.contiguous(memory_format=torch.channels_last)


# ---- detector input -----------------------------------------------------------------------------------------------------
def _detector_input(chunk_u8, vmin, vmax, int_limits, mean, std, ph, pw, bf16):
    n, h, w = (int(v) for v in chunk_u8.shape)
    x = torch.empty((n, 3, ph, pw), dtype=torch.bfloat16 if bf16 else torch.float32, device=chunk_u8.device,
                    memory_format=torch.channels_last)
    _lib.call('msq_detector_input', _dev.ptr(chunk_u8.contiguous()), _dev.ptr(x), int(bf16), n, h, w, h, w, ph, pw,
              (ctypes.c_float * 3)(*[float(v) for v in mean]), (ctypes.c_float * 3)(*[float(v) for v in std]),
              float(vmin), float(vmax), int(int_limits), _dev.stream())
    return x


def _stem_conv_pool(chunk_u8, vmin, vmax, int_limits, mean, std, ph, pw, w49x64, bias64, bf16):
    n, h, w = (int(v) for v in chunk_u8.shape)
    conv_h, conv_w = (ph - 1) // 2 + 1, (pw - 1) // 2 + 1
    pool_h, pool_w = (conv_h - 1) // 2 + 1, (conv_w - 1) // 2 + 1
    out = torch.empty((n, 64, pool_h, pool_w), dtype=torch.bfloat16 if bf16 else torch.float32, device=chunk_u8.device,
                      memory_format=torch.channels_last)
    _lib.call('msq_stem_conv_pool', _dev.ptr(chunk_u8.contiguous()), n, h, w, int(ph), int(pw), float(vmin), float(vmax), int(int_limits),
              float(mean), float(std), _dev.ptr(w49x64), _dev.ptr(bias64), _dev.ptr(out), int(bf16), _dev.stream())
    return out


def _stem_conv_pool_tc(chunk_u8, vmin, vmax, int_limits, mean, std, ph, pw, b_tile, bias64):
    n, h, w = (int(v) for v in chunk_u8.shape)
    conv_h, conv_w = (ph - 1) // 2 + 1, (pw - 1) // 2 + 1
    pool_h, pool_w = (conv_h - 1) // 2 + 1, (conv_w - 1) // 2 + 1
    out = torch.empty((n, 64, pool_h, pool_w), dtype=torch.bfloat16, device=chunk_u8.device, memory_format=torch.channels_last)
    _lib.call('msq_stem_conv_pool_tc', _dev.ptr(chunk_u8.contiguous()), n, h, w, int(ph), int(pw), float(vmin), float(vmax), int(int_limits),
              float(mean), float(std), _dev.ptr(b_tile), _dev.ptr(bias64), _dev.ptr(out), _dev.stream())
    return out


# ---- dense contractions -------------------------------------------------------------------------------------------------
# Two engines run the convolutions / Linear layers of the graph: cuDNN / cuBLAS (library calls) and csrc/conv_tc.cu (the repo's
# tcgen05 + TMA implicit GEMM).  CONV_ENGINE['mode']: 'cudnn', 'tcgen05' (wherever the shape is served), or 'auto': the first
# time a layer shape is seen both engines are timed on it (CUDA events, 3 runs each) and the faster one keeps the shape.
_ENGINE_CHOICE: Dict[tuple, str] = {}


def engine_choices() -> Dict[tuple, str]:
    """Layer shape -> engine picked by 'auto' so far (bench.py prints the split)."""
    return dict(_ENGINE_CHOICE)


def _time_call(fn, runs: int = 3) -> float:
    fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(runs):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / runs


def _conv2d_cudnn(x, w, b, z, relu, stride, pad):
    x = _cl(x)
    st, pd, dl = [stride, stride], [pad, pad], [1, 1]
    if relu and b is not None and x.dtype == w.dtype:
        if z is None:
            return torch.cudnn_convolution_relu(x, w, b, st, pd, dl, 1)
        return torch.cudnn_convolution_add_relu(x, w, _cl(z), 1.0, b, st, pd, dl, 1)
    y = F.conv2d(x, w, b, st, pd)
    if z is not None:
        y = y + z
    return F.relu(y) if relu else y


def _conv2d(x, w, b, z, relu, stride, pad):
    mode = CONV_ENGINE['mode']
    if mode != 'cudnn' and x.dtype == torch.bfloat16:
        from . import conv_tc
        if mode == 'tcgen05':
            y = conv_tc.try_conv2d(x, w, b, z, relu, stride, pad)
            if y is not None:
                return y
        else:                                                    # 'auto'
            key = ('conv', tuple(x.shape), tuple(w.shape), int(stride), int(pad), z is not None, bool(relu), b is not None)
            choice = _ENGINE_CHOICE.get(key)
            if choice is None:
                choice = 'cudnn'
                if conv_tc.try_conv2d(x, w, b, z, relu, stride, pad) is not None:
                    t_tc = _time_call(lambda: conv_tc.try_conv2d(x, w, b, z, relu, stride, pad))
                    t_cd = _time_call(lambda: _conv2d_cudnn(x, w, b, z, relu, stride, pad))
                    choice = 'tcgen05' if t_tc < t_cd else 'cudnn'
                _ENGINE_CHOICE[key] = choice
            if choice == 'tcgen05':
                return conv_tc.try_conv2d(x, w, b, z, relu, stride, pad)
    return _conv2d_cudnn(x, w, b, z, relu, stride, pad)


def _linear_cublas(x, w, b, relu):
    y = F.linear(x, w, b)
    return F.relu_(y) if relu else y


def _linear(x, w, b, relu):
    mode = CONV_ENGINE['mode']
    if mode != 'cudnn' and x.dtype == torch.bfloat16:
        from . import conv_tc
        if mode == 'tcgen05':
            y = conv_tc.try_linear(x, w, b, relu)
            if y is not None:
                return y
        else:
            key = ('linear', tuple(x.shape), tuple(w.shape), bool(relu), b is not None)
            choice = _ENGINE_CHOICE.get(key)
            if choice is None:
                choice = 'cudnn'
                if conv_tc.try_linear(x, w, b, relu) is not None:
                    t_tc = _time_call(lambda: conv_tc.try_linear(x, w, b, relu))
                    t_cd = _time_call(lambda: _linear_cublas(x, w, b, relu))
                    choice = 'tcgen05' if t_tc < t_cd else 'cudnn'
                _ENGINE_CHOICE[key] = choice
            if choice == 'tcgen05':
                return conv_tc.try_linear(x, w, b, relu)
    return _linear_cublas(x, w, b, relu)


# ---- GroupNorm (+ top-down merge) on channels-last maps -------------------------------------------------------------------
def _group_norm_nhwc(x, gamma, beta, groups, eps, top, scale):
    x = _cl(x)
    n, c, h, w = (int(v) for v in x.shape)
    if top is not None:
        top = _cl(top)
        if (int(top.shape[2]), int(top.shape[3])) != ((h + 1) // 2, (w + 1) // 2) or h % 2 or w % 2:
            raise ValueError(f'group_norm_nhwc: top-down map {tuple(top.shape)} is not half of {tuple(x.shape)}')
    out = torch.empty_like(x, memory_format=torch.channels_last)
    nbytes = int(_lib.load().msq_group_norm_scratch_bytes(n, h, w, c))
    scratch = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    _lib.call('msq_group_norm_nhwc', _dev.ptr(x), int(x.dtype == torch.bfloat16), n, h, w, c, int(groups), float(eps),
              _dev.ptr(gamma), _dev.ptr(beta), _dev.ptr(top), float(scale), _dev.ptr(out), _dev.ptr(scratch), nbytes, _dev.stream())
    return out


# ---- RPN: detectron2 find_top_rpn_proposals for a batch of equally sized images ---------------------------------------------
_ANCHOR_CACHE: Dict[tuple, torch.Tensor] = {}


def cell_anchors(size: float, ratios) -> torch.Tensor:
    """detectron2 DefaultAnchorGenerator.generate_cell_anchors for one size."""
    out = []
    for ar in ratios:
        area = size ** 2.0
        w = math.sqrt(area / ar)
        h = ar * w
        out.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(out, dtype=torch.float32)


def grid_anchors(gh: int, gw: int, stride: int, size: float, ratios, device) -> torch.Tensor:
    """(gh * gw * A, 4) anchors of one pyramid level in detectron2's (y, x, anchor) order, offset 0."""
    key = (gh, gw, stride, float(size), tuple(float(r) for r in ratios), str(device))
    hit = _ANCHOR_CACHE.get(key)
    if hit is None:
        sx = torch.arange(0, gw * stride, step=stride, dtype=torch.float32)
        sy = torch.arange(0, gh * stride, step=stride, dtype=torch.float32)
        yy, xx = torch.meshgrid(sy, sx, indexing='ij')
        shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
        hit = (shifts.view(-1, 1, 4) + cell_anchors(size, ratios).view(1, -1, 4)).reshape(-1, 4).to(device)
        _ANCHOR_CACHE[key] = hit
    return hit


_SCALE_CLAMP = math.log(1000.0 / 16)


def apply_deltas(deltas: torch.Tensor, boxes: torch.Tensor, weights=(1.0, 1.0, 1.0, 1.0)) -> torch.Tensor:
    """detectron2 Box2BoxTransform.apply_deltas on (..., 4) tensors."""
    deltas = deltas.float()
    boxes = boxes.to(deltas.dtype)
    widths = boxes[..., 2] - boxes[..., 0]
    heights = boxes[..., 3] - boxes[..., 1]
    ctr_x = boxes[..., 0] + 0.5 * widths
    ctr_y = boxes[..., 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx, dy = deltas[..., 0] / wx, deltas[..., 1] / wy
    dw = torch.clamp(deltas[..., 2] / ww, max=_SCALE_CLAMP)
    dh = torch.clamp(deltas[..., 3] / wh, max=_SCALE_CLAMP)
    pcx, pcy = dx * widths + ctr_x, dy * heights + ctr_y
    pw, ph = torch.exp(dw) * widths, torch.exp(dh) * heights
    return torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=-1)


def _rpn_proposals(preds, strides, sizes, ratios, img_h, img_w, pre_topk, post_topk, nms_thresh):
    """preds[l]: (n, 16, H_l, W_l) channels-last, channels [A objectness logits, A x 4 anchor deltas, padding] (A = 3).
    Returns proposals (n, post_topk, 4) float32 (zero boxes beyond the count), their logits (n, post_topk), counts (n) int32."""
    n = int(preds[0].shape[0])
    dev = preds[0].device
    A = len(ratios)
    fused = _rpn_select_fused(preds, strides, sizes, ratios, img_h, img_w, pre_topk) if RPN_ENGINE['mode'] == 'fused' else None
    if fused is not None:
        boxes, shifted, scores, valid_u8 = fused            # valid_u8 = 1 + pyramid level of every candidate (0: none)
        return _nms_and_gather(boxes, shifted, scores, valid_u8, post_topk, nms_thresh, levels=len(preds), per_level=int(pre_topk))
    boxes_l, scores_l, level_l = [], [], []
    for lvl, p in enumerate(preds):
        gh, gw = int(p.shape[2]), int(p.shape[3])
        p = _cl(p).permute(0, 2, 3, 1)                                   # (n, H, W, 16) view of the channels-last memory
        logits = p[..., :A].reshape(n, -1).float()                       # (h, w, a) order, as detectron2 flattens it
        deltas = p[..., A:5 * A].reshape(n, -1, 4).float()
        anchors = grid_anchors(gh, gw, int(strides[lvl]), float(sizes[lvl]), ratios, dev)
        k = min(int(pre_topk), gh * gw * A)
        top, idx = logits.topk(k, dim=1)
        sel = torch.gather(deltas, 1, idx[..., None].expand(-1, -1, 4))
        boxes_l.append(apply_deltas(sel, anchors[idx]))
        scores_l.append(top)
        level_l.append(torch.full((k,), lvl, dtype=torch.float32, device=dev))
    boxes = torch.cat(boxes_l, 1)
    scores = torch.cat(scores_l, 1)
    levels = torch.cat(level_l)[None].expand(n, -1)
    finite = torch.isfinite(boxes).all(-1) & torch.isfinite(scores)
    boxes = torch.stack([boxes[..., 0].clamp(0, img_w), boxes[..., 1].clamp(0, img_h),
                         boxes[..., 2].clamp(0, img_w), boxes[..., 3].clamp(0, img_h)], dim=-1)
    valid = finite & ((boxes[..., 2] - boxes[..., 0]) > 0) & ((boxes[..., 3] - boxes[..., 1]) > 0)
    order = torch.sort(torch.where(valid, scores, scores.new_full((), float('-inf'))), dim=1, descending=True, stable=True).indices
    boxes = torch.gather(boxes, 1, order[..., None].expand(-1, -1, 4))
    scores = torch.gather(scores, 1, order)
    levels = torch.gather(levels, 1, order)
    valid = torch.gather(valid, 1, order)
    # torchvision batched_nms' coordinate trick (what detectron2's batched_nms runs for < 40 000 boxes): shift every level by
    # (largest coordinate of the image's boxes + 1) so that levels never suppress each other
    max_coord = torch.where(valid[..., None], boxes, boxes.new_full((), float('-inf'))).amax(dim=(1, 2))
    max_coord = torch.where(torch.isfinite(max_coord), max_coord, torch.zeros_like(max_coord))
    shifted = (boxes + (levels * (max_coord[:, None] + 1))[..., None]).contiguous()
    return _nms_and_gather(boxes, shifted, scores, valid.to(torch.uint8).contiguous(), post_topk, nms_thresh)


def _rpn_select_fused(preds, strides, sizes, ratios, img_h, img_w, pre_topk):
    """Everything of find_top_rpn_proposals before the NMS in one launch (`msq_rpn_select`, csrc/nms.cu); None when the shapes are
    not the ones that kernel serves (then the operator-by-operator path below runs)."""
    k = len(preds)
    if len(ratios) != 3 or k > 8 or any(int(p.shape[1]) != 16 for p in preds) or preds[0].dtype not in (torch.bfloat16, torch.float32):
        return None
    if any(p.dtype != preds[0].dtype for p in preds) or any(int(p.shape[2]) * int(p.shape[3]) * 3 >= (1 << 14) for p in preds):
        return None
    total = sum(min(int(pre_topk), int(p.shape[2]) * int(p.shape[3]) * 3) for p in preds)
    if total > 4096:
        return None
    preds = [_cl(p) for p in preds]
    n, dev = int(preds[0].shape[0]), preds[0].device
    cells = torch.stack([cell_anchors(float(sizes[i]), ratios) for i in range(k)]).reshape(-1).tolist()
    boxes = torch.empty((n, total, 4), dtype=torch.float32, device=dev)
    shifted = torch.empty((n, total, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((n, total), dtype=torch.float32, device=dev)
    valid = torch.empty((n, total), dtype=torch.uint8, device=dev)
    _lib.call('msq_rpn_select', (ctypes.c_void_p * k)(*[p.data_ptr() for p in preds]), (ctypes.c_int * k)(*[int(p.shape[2]) for p in preds]),
              (ctypes.c_int * k)(*[int(p.shape[3]) for p in preds]), (ctypes.c_int * k)(*[int(v) for v in strides]),
              (ctypes.c_float * len(cells))(*cells), k, int(preds[0].dtype == torch.bfloat16), n, int(pre_topk), int(img_h), int(img_w),
              _dev.ptr(boxes), _dev.ptr(shifted), _dev.ptr(scores), _dev.ptr(valid), _dev.stream())
    return boxes, shifted, scores, valid


def _nms_and_gather(boxes, shifted, scores, valid_u8, post_topk, nms_thresh, levels=None, per_level=None):
    n, K = int(boxes.shape[0]), int(boxes.shape[1])
    dev = boxes.device
    keep = torch.empty((n, int(post_topk)), dtype=torch.int32, device=dev)
    count = torch.empty((n,), dtype=torch.int32, device=dev)
    if int(post_topk) > 128 and levels is not None and levels <= 8 and min(K, int(per_level)) <= 2048:
        # long keep lists, levels known: overlap bits per level only, one walk per (image, level)
        nbytes = int(_lib.load().msq_nms_levels_scratch_bytes(n, K, int(levels), int(per_level)))
        scratch = torch.empty((nbytes + 256,), dtype=torch.uint8, device=dev)
        scratch = scratch[(-scratch.data_ptr()) % 256:]
        _lib.call('msq_nms_levels_long', _dev.ptr(shifted), _dev.ptr(valid_u8), n, K, int(levels), int(per_level), float(nms_thresh),
                  int(post_topk), _dev.ptr(keep), _dev.ptr(count), _dev.ptr(scratch), nbytes, _dev.stream())
    elif int(post_topk) > 128 and K <= 6144:       # long keep lists: overlap matrix as bit rows + one ordered walk per image
        nbytes = int(_lib.load().msq_nms_scratch_bytes(n, K))
        scratch = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        _lib.call('msq_nms_sorted_long', _dev.ptr(shifted), _dev.ptr(valid_u8), n, K, float(nms_thresh), int(post_topk), _dev.ptr(keep),
                  _dev.ptr(count), _dev.ptr(scratch), nbytes, _dev.stream())
    else:
        _lib.call('msq_nms_sorted', _dev.ptr(shifted), _dev.ptr(valid_u8), n, K, float(nms_thresh), int(post_topk), _dev.ptr(keep),
                  _dev.ptr(count), _dev.stream())
    ok = keep >= 0
    idx = keep.clamp(min=0).long()
    out_boxes = torch.gather(boxes, 1, idx[..., None].expand(-1, -1, 4)) * ok[..., None]
    out_scores = torch.where(ok, torch.gather(scores, 1, idx), scores.new_full((), float('-inf')))
    return out_boxes.contiguous(), out_scores.contiguous(), count


# ---- ROIAlignV2 over the pyramid -------------------------------------------------------------------------------------------
def _roi_align_v2(feats, scales, boxes, rois_per_image, pooled, sampling_ratio, min_level, canonical_level, canonical_size):
    """feats[l] (n, C, H_l, W_l) channels-last; boxes (R, 4) float32, box r on image r // rois_per_image.  Returns (R, C, P, P)
    in channels-last memory (= (R, P, P, C) contiguous) in the dtype of the maps."""
    feats = [_cl(f) for f in feats]
    k = len(feats)
    c = int(feats[0].shape[1])
    boxes = boxes.reshape(-1, 4).float().contiguous()
    r = int(boxes.shape[0])
    out = torch.empty((r, c, int(pooled), int(pooled)), dtype=feats[0].dtype, device=feats[0].device, memory_format=torch.channels_last)
    if r == 0:
        return out
    _lib.call('msq_roi_align_v2', (ctypes.c_void_p * k)(*[f.data_ptr() for f in feats]), (ctypes.c_int * k)(*[int(f.shape[2]) for f in feats]),
              (ctypes.c_int * k)(*[int(f.shape[3]) for f in feats]), (ctypes.c_float * k)(*[float(s) for s in scales]), k, c,
              int(feats[0].dtype == torch.bfloat16), _dev.ptr(boxes), r, int(rois_per_image), int(pooled), int(sampling_ratio),
              int(min_level), int(canonical_level), float(canonical_size), _dev.ptr(out), _dev.stream())
    return out


# ---- Fast R-CNN outputs, one detection per image ----------------------------------------------------------------------------
def _fastrcnn_top1(pred, proposals, counts, img_h, img_w, score_thresh, weights):
    n, k = int(proposals.shape[0]), int(proposals.shape[1])
    pred = pred.float().contiguous()
    proposals = proposals.float().contiguous()
    dev = pred.device
    box = torch.empty((n, 4), dtype=torch.float32, device=dev)
    score = torch.empty((n,), dtype=torch.float32, device=dev)
    has = torch.empty((n,), dtype=torch.uint8, device=dev)
    _lib.call('msq_fastrcnn_top1', _dev.ptr(pred), int(pred.shape[1]), _dev.ptr(proposals), _dev.ptr(counts.to(torch.int32).contiguous()), n, k,
              int(img_h), int(img_w), float(score_thresh), (ctypes.c_float * 4)(*[float(v) for v in weights]), _dev.ptr(box),
              _dev.ptr(score), _dev.ptr(has), None, _dev.stream())
    return box, score, has


def _keypoints_from_heatmaps_d2(heatmaps, boxes):
    r, k, hm, wm = (int(v) for v in heatmaps.shape)
    xyp = torch.empty((r, k, 3), dtype=torch.float32, device=heatmaps.device)
    if r:
        _lib.call('msq_keypoints_from_heatmaps_d2', _dev.ptr(heatmaps.float().contiguous()), _dev.ptr(boxes.float().contiguous()), r, k, hm, wm,
                  _dev.ptr(xyp), None, _dev.stream())
    return xyp


def _upsample2x_bilinear(x):
    n, k, h, w = (int(v) for v in x.shape)
    if x.dtype not in (torch.bfloat16, torch.float32):
        x = x.float()
    out = torch.empty((n, k, 2 * h, 2 * w), dtype=torch.float32, device=x.device)
    if n:
        sn, sc, sh, sw = (int(v) for v in x.stride())
        _lib.call('msq_upsample2x_bilinear', _dev.ptr(x), int(x.dtype == torch.bfloat16), sn, sc, sh, sw, n, k, h, w, _dev.ptr(out), _dev.stream())
    return out


for _name, _fn in (('upsample2x_bilinear', _upsample2x_bilinear), ('detector_input', _detector_input), ('stem_conv_pool', _stem_conv_pool), ('stem_conv_pool_tc', _stem_conv_pool_tc), ('conv2d', _conv2d), ('linear', _linear), ('group_norm_nhwc', _group_norm_nhwc),
                   ('rpn_proposals', _rpn_proposals), ('roi_align_v2', _roi_align_v2), ('fastrcnn_top1', _fastrcnn_top1),
                   ('keypoints_from_heatmaps_d2', _keypoints_from_heatmaps_d2)):
    _LIB.impl(_name, _fn, 'CUDA')
