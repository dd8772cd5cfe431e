"""Batched replacements for the per-image loops of torchvision's detection post-processing (a4 glue).

`RegionProposalNetwork.filter_proposals` and `RoIHeads.postprocess_detections` loop over the images of a batch in Python;
every iteration is a dozen tiny kernels plus an NMS that synchronises with the host (tools/rcnn_profile.py: 0.6 and 0.45 ms
per frame, more than the backbone).  When all images of the batch have the same size -- always the case on the extract
path -- the same arithmetic runs on (n, K, ...) tensors with ONE segmented NMS launch (`msq_nms_sorted`, csrc/nms.cu).
`enable_batched_heads(model)` patches a torchvision GeneralizedRCNN in place; outputs are identical to the per-image code
(tests/test_gpu_pipeline.py), and any batch the fast path cannot serve (mixed image sizes, fewer survivors than
`post_nms_top_n`, several detections per image) falls back to torchvision's own method.
"""
from __future__ import annotations

import types
from typing import List, Tuple

import torch
import torch.nn.functional as F

from .. import _dev, _lib


def _same_shapes(image_shapes) -> bool:
    return len(image_shapes) > 0 and all(tuple(s) == tuple(image_shapes[0]) for s in image_shapes)


def _filter_proposals(self, proposals, objectness, image_shapes, num_anchors_per_level):
    """RegionProposalNetwork.filter_proposals for a batch of equally sized images."""
    if not (proposals.is_cuda and _same_shapes(image_shapes)):
        return self._msq_filter_proposals(proposals, objectness, image_shapes, num_anchors_per_level)
    n = proposals.shape[0]
    device = proposals.device
    objectness = objectness.detach().reshape(n, -1)
    levels = torch.cat([torch.full((c,), idx, dtype=torch.int64, device=device) for idx, c in enumerate(num_anchors_per_level)], 0)
    levels = levels.reshape(1, -1).expand_as(objectness)
    top_n_idx = self._get_top_n_idx(objectness, num_anchors_per_level)
    batch_idx = torch.arange(n, device=device)[:, None]
    objectness = objectness[batch_idx, top_n_idx]
    levels = levels[batch_idx, top_n_idx]
    boxes = proposals[batch_idx, top_n_idx].float()
    prob = torch.sigmoid(objectness).float()
    # ---- per-image loop of torchvision, on (n, K) tensors ----
    height, width = image_shapes[0]
    boxes = torch.stack([boxes[..., 0].clamp(0, width), boxes[..., 1].clamp(0, height),
                         boxes[..., 2].clamp(0, width), boxes[..., 3].clamp(0, height)], dim=-1)
    ws, hs = boxes[..., 2] - boxes[..., 0], boxes[..., 3] - boxes[..., 1]
    valid = (ws >= self.min_size) & (hs >= self.min_size) & (prob >= self.score_thresh)
    order = torch.sort(torch.where(valid, prob, prob.new_full((), -1.0)), dim=1, descending=True, stable=True).indices
    boxes = torch.gather(boxes, 1, order[..., None].expand(-1, -1, 4))
    prob = torch.gather(prob, 1, order)
    levels = torch.gather(levels, 1, order)
    valid = torch.gather(valid, 1, order)
    # batched_nms' coordinate trick: shift every level by (largest coordinate of the image's boxes + 1)
    max_coord = torch.where(valid[..., None], boxes, boxes.new_full((), float('-inf'))).amax(dim=(1, 2))
    shifted = (boxes + (levels.to(boxes) * (max_coord[:, None] + 1))[..., None]).contiguous()
    K, top_n = int(boxes.shape[1]), int(self.post_nms_top_n())
    keep = torch.empty((n, top_n), dtype=torch.int32, device=device)
    count = torch.empty((n,), dtype=torch.int32, device=device)
    valid_u8 = valid.to(torch.uint8).contiguous()
    _lib.call('msq_nms_sorted', _dev.ptr(shifted), _dev.ptr(valid_u8), n, K, float(self.nms_thresh), top_n, _dev.ptr(keep),
              _dev.ptr(count), _dev.stream())
    if int(count.min()) < top_n:            # an image with fewer survivors than requested: ragged per-image lists
        return _ragged(boxes, prob, keep, count)
    keep = keep.long()
    final_boxes = torch.gather(boxes, 1, keep[..., None].expand(-1, -1, 4))
    final_scores = torch.gather(prob, 1, keep)
    return list(final_boxes.unbind(0)), list(final_scores.unbind(0))


def _ragged(boxes, prob, keep, count):
    """Per-image lists when some image kept fewer than `post_nms_top_n` boxes."""
    counts = count.tolist()
    out_b, out_s = [], []
    for i, c in enumerate(counts):
        idx = keep[i, :c].long()
        out_b.append(boxes[i, idx])
        out_s.append(prob[i, idx])
    return out_b, out_s


def _postprocess_detections(self, class_logits, box_regression, proposals, image_shapes):
    """RoIHeads.postprocess_detections when one detection per image is kept: after score / size filtering the best box
    always survives its class's NMS, so the result is a (stable) arg-max -- no NMS at all."""
    per_image = [int(p.shape[0]) for p in proposals]
    num_classes = int(class_logits.shape[-1])
    if not (class_logits.is_cuda and self.detections_per_img == 1 and num_classes == 2 and _same_shapes(image_shapes)
            and len(set(per_image)) == 1 and per_image[0] > 0):
        return self._msq_postprocess_detections(class_logits, box_regression, proposals, image_shapes)
    n, k = len(proposals), per_image[0]
    pred_boxes = self.box_coder.decode(box_regression, proposals)              # (n*k, classes, 4)
    pred_scores = F.softmax(class_logits, -1)
    height, width = image_shapes[0]
    boxes = pred_boxes.reshape(n, k, num_classes, 4)[:, :, 1:].reshape(n, -1, 4)
    scores = pred_scores.reshape(n, k, num_classes)[:, :, 1:].reshape(n, -1)
    boxes = torch.stack([boxes[..., 0].clamp(0, width), boxes[..., 1].clamp(0, height),
                         boxes[..., 2].clamp(0, width), boxes[..., 3].clamp(0, height)], dim=-1)
    ws, hs = boxes[..., 2] - boxes[..., 0], boxes[..., 3] - boxes[..., 1]
    valid = (scores > self.score_thresh) & (ws >= 1e-2) & (hs >= 1e-2)
    best = torch.sort(torch.where(valid, scores, scores.new_full((), -1.0)), dim=1, descending=True, stable=True).indices[:, :1]
    top_boxes = torch.gather(boxes, 1, best[..., None].expand(-1, -1, 4))      # (n, 1, 4)
    top_scores = torch.gather(scores, 1, best)
    has = torch.gather(valid, 1, best)[:, 0].tolist()                          # one small read-back per batch
    labels = torch.ones((1,), dtype=torch.int64, device=class_logits.device)
    all_boxes, all_scores, all_labels = [], [], []
    for i in range(n):
        c = 1 if has[i] else 0
        all_boxes.append(top_boxes[i, :c])
        all_scores.append(top_scores[i, :c])
        all_labels.append(labels[:c])
    return all_boxes, all_scores, all_labels


def keypoints_from_heatmaps(maps: torch.Tensor, rois: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """torchvision's heatmaps_to_keypoints for all RoIs at once (`msq_keypoints_from_heatmaps`, csrc/nms.cu):
    maps (R, K, Hm, Wm), rois (R, 4) -> xy_preds (R, K, 3) float32, scores (R, K) float32."""
    r, k, hm, wm = (int(v) for v in maps.shape)
    xyv = torch.empty((r, k, 3), dtype=torch.float32, device=maps.device)
    scores = torch.empty((r, k), dtype=torch.float32, device=maps.device)
    if r == 0:
        return xyv, scores
    m32 = maps.float().contiguous()
    b32 = rois.float().contiguous()
    _lib.call('msq_keypoints_from_heatmaps', _dev.ptr(m32), _dev.ptr(b32), r, k, hm, wm, int(maps.dtype == torch.bfloat16),
              _dev.ptr(xyv), _dev.ptr(scores), _dev.stream())
    return xyv, scores


def _keypointrcnn_inference(x, boxes):
    """torchvision.models.detection.roi_heads.keypointrcnn_inference without the per-RoI Python loop."""
    if not x.is_cuda:
        return _keypointrcnn_inference.original(x, boxes)
    per_image = [int(b.shape[0]) for b in boxes]
    xyv, scores = keypoints_from_heatmaps(x, torch.cat(boxes, dim=0))
    return list(xyv.split(per_image, 0)), list(scores.split(per_image, 0))


def _transform_forward(self, images, targets=None):
    """GeneralizedRCNNTransform.forward for a batch of equally sized CUDA images in eval mode: normalise, resize and pad
    once on the stacked batch (per-sample arithmetic identical to the per-image loop)."""
    from torchvision.models.detection.image_list import ImageList
    same = (targets is None and not self.training and len(images) > 0 and self.fixed_size is None
            and all(i.is_cuda and i.dim() == 3 and i.shape == images[0].shape and i.dtype == images[0].dtype for i in images))
    if not same:
        return self._msq_forward(images, targets)
    x = torch.stack(list(images))
    if not x.is_floating_point():
        raise TypeError(f'Expected input images to be of floating type (in range [0, 1]), but found type {x.dtype} instead')
    mean = torch.as_tensor(self.image_mean, dtype=x.dtype, device=x.device)
    std = torch.as_tensor(self.image_std, dtype=x.dtype, device=x.device)
    x = (x - mean[None, :, None, None]) / std[None, :, None, None]
    h, w = int(x.shape[-2]), int(x.shape[-1])
    scale_factor = min(self.min_size[-1] / min(h, w), self.max_size / max(h, w))
    x = F.interpolate(x, size=None, scale_factor=scale_factor, mode='bilinear', recompute_scale_factor=True, align_corners=False)
    oh, ow = int(x.shape[-2]), int(x.shape[-1])
    div = int(self.size_divisible)
    ph, pw = -(-oh // div) * div, -(-ow // div) * div
    if (ph, pw) != (oh, ow):
        x = F.pad(x, (0, pw - ow, 0, ph - oh))
    return ImageList(x, [(oh, ow)] * len(images)), targets


def _transform_postprocess(self, result, image_shapes, original_image_sizes):
    """GeneralizedRCNNTransform.postprocess (boxes and keypoints back to the original image scale) as one multiply over
    the concatenated detections instead of ~40 tiny ops and 8 host-to-device scalar copies per image.  The ratios are the
    float32 quotients torchvision forms (resize_boxes / resize_keypoints).  Mask pasting is not done here: models whose
    soft masks are pasted by torchvision go through the original loop."""
    if self.training or len(result) == 0:
        return result
    if any('masks' in r for r in result) and not getattr(self, '_msq_keep_soft_masks', False):
        return self._msq_postprocess(result, image_shapes, original_image_sizes)
    import numpy as np
    ratios = np.array([[np.float32(o[1]) / np.float32(s[1]), np.float32(o[0]) / np.float32(s[0])]
                       for s, o in zip(image_shapes, original_image_sizes)], dtype=np.float32)            # (n, 2): width, height
    counts = [int(r['boxes'].shape[0]) for r in result]
    dev = result[0]['boxes'].device
    if (ratios == ratios[0]).all():
        per_row = torch.tensor(ratios[0], device=dev)[None, :]
    else:
        per_row = torch.tensor(np.repeat(ratios, counts, axis=0), device=dev)
    boxes = torch.cat([r['boxes'] for r in result]) * torch.cat([per_row, per_row], dim=1)
    for r, b in zip(result, boxes.split(counts)):
        r['boxes'] = b
    if all('keypoints' in r for r in result):
        kp = torch.cat([r['keypoints'] for r in result]).clone()
        kp[..., :2] *= per_row[:, None, :]
        for r, k in zip(result, kp.split(counts)):
            r['keypoints'] = k
    return result


def _convert_to_roi_format(boxes):
    """torchvision.ops.poolers._convert_to_roi_format without one tiny kernel per image: (image index, x1, y1, x2, y2)."""
    counts = [int(b.shape[0]) for b in boxes]
    concat = torch.cat(boxes, dim=0)
    if len(boxes) == 0 or not concat.is_cuda:
        return _convert_to_roi_format.original(boxes)
    if len(set(counts)) == 1:
        ids = torch.arange(len(boxes), device=concat.device, dtype=concat.dtype).repeat_interleave(counts[0])
    else:
        ids = torch.repeat_interleave(torch.arange(len(boxes), device=concat.device, dtype=concat.dtype),
                                      torch.tensor(counts, device=concat.device))
    return torch.cat([ids[:, None], concat], dim=1)


def _level_mapper_call(self, boxlists):
    """torchvision.ops.poolers.LevelMapper.__call__ with the box areas taken on the concatenated boxes (the original
    computes them image by image: three tiny kernels per image and pooler)."""
    boxes = torch.cat(list(boxlists))
    if boxes.dtype in (torch.float16, torch.bfloat16) or not boxes.is_floating_point():
        return _level_mapper_call.original(self, boxlists)
    s = torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))
    target_lvls = torch.floor(self.lvl0 + torch.log2(s / self.s0) + torch.tensor(self.eps, dtype=s.dtype))
    target_lvls = torch.clamp(target_lvls, min=self.k_min, max=self.k_max)
    return (target_lvls.to(torch.int64) - self.k_min).to(torch.int64)


def _multiscale_roi_align(x_filtered, boxes, output_size, sampling_ratio, scales, mapper):
    """torchvision.ops.poolers._multiscale_roi_align as ONE launch of `msq_roi_align_levels` (csrc/roi_align.cu) on the
    channels-last feature maps: no per-level torch.where host syncs, gathers, float32 / NCHW conversions or index_put."""
    import ctypes
    feats = list(x_filtered)
    ok = (len(feats) >= 1 and feats[0].is_cuda and scales is not None and (mapper is not None or len(feats) == 1)
          and 1 <= int(sampling_ratio) <= 4 and len(feats) <= 8 and output_size[0] == output_size[1]
          and feats[0].dtype in (torch.bfloat16, torch.float32) and all(f.dtype == feats[0].dtype for f in feats)
          and feats[0].shape[1] % 8 == 0 and not torch.is_grad_enabled())
    if not ok:
        return _multiscale_roi_align.original(x_filtered, boxes, output_size, sampling_ratio, scales, mapper)
    rois = _convert_to_roi_format(boxes).float().contiguous()
    levels = mapper(boxes).contiguous() if len(feats) > 1 else None
    n_rois, channels, pooled = int(rois.shape[0]), int(feats[0].shape[1]), int(output_size[0])
    # (n, C, H, W) in channels-last memory is (n, H, W, C) contiguous: free for the autocast backbone's outputs
    nhwc = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    out = torch.empty((n_rois, channels, pooled, pooled), dtype=feats[0].dtype, device=feats[0].device)
    k = len(nhwc)
    _lib.call('msq_roi_align_levels', (ctypes.c_void_p * k)(*[f.data_ptr() for f in nhwc]), (ctypes.c_int * k)(*[int(f.shape[2]) for f in nhwc]),
              (ctypes.c_int * k)(*[int(f.shape[3]) for f in nhwc]), (ctypes.c_float * k)(*[float(s) for s in scales]), k, channels,
              int(feats[0].dtype == torch.bfloat16), _dev.ptr(rois), _dev.ptr(levels), n_rois, pooled, int(sampling_ratio), _dev.ptr(out),
              _dev.stream())
    return out


def _roi_heads_forward(self, features, proposals, image_shapes, targets=None):
    """RoIHeads.forward with the pyramid features laid out ONCE for the three poolers: channels-last in their own dtype
    for `_multiscale_roi_align` above; contiguous float32 when torchvision's roi_align is in use (it wants NCHW float32
    and otherwise casts / re-lays-out its whole input on every call, three poolers x four levels per batch)."""
    from torchvision.ops import poolers as tv_poolers
    if tv_poolers._multiscale_roi_align is _multiscale_roi_align:
        # our pooler reads channels-last maps of either dtype in place; make them channels-last once, not per pooler
        if not self.training:
            features = type(features)((k, v.contiguous(memory_format=torch.channels_last)) for k, v in features.items())
    elif not self.training and any(v.dtype != torch.float32 or not v.is_contiguous() for v in features.values()):
        features = type(features)((k, v.to(torch.float32, memory_format=torch.contiguous_format)) for k, v in features.items())
    return self._msq_forward(features, proposals, image_shapes, targets)


def enable_batched_heads(model) -> None:
    """Patch a torchvision detection model (RPN + RoIHeads) in place; idempotent."""
    rpn, heads = model.rpn, model.roi_heads
    if not hasattr(rpn, '_msq_filter_proposals'):
        rpn._msq_filter_proposals = rpn.filter_proposals
        rpn.filter_proposals = types.MethodType(_filter_proposals, rpn)
    if not hasattr(heads, '_msq_postprocess_detections'):
        heads._msq_postprocess_detections = heads.postprocess_detections
        heads.postprocess_detections = types.MethodType(_postprocess_detections, heads)
    tr = model.transform
    if not hasattr(tr, '_msq_forward'):
        tr._msq_forward = tr.forward
        tr.forward = types.MethodType(_transform_forward, tr)
    if not hasattr(heads, '_msq_forward'):
        heads._msq_forward = heads.forward
        heads.forward = types.MethodType(_roi_heads_forward, heads)
    if not hasattr(tr, '_msq_postprocess'):
        tr._msq_postprocess = tr.postprocess
        tr.postprocess = types.MethodType(_transform_postprocess, tr)
    from torchvision.ops import poolers as tv_poolers
    if tv_poolers._convert_to_roi_format is not _convert_to_roi_format:
        _convert_to_roi_format.original = tv_poolers._convert_to_roi_format
        tv_poolers._convert_to_roi_format = _convert_to_roi_format
    if tv_poolers._multiscale_roi_align is not _multiscale_roi_align:
        _multiscale_roi_align.original = tv_poolers._multiscale_roi_align
        tv_poolers._multiscale_roi_align = _multiscale_roi_align
    if tv_poolers.LevelMapper.__call__ is not _level_mapper_call:
        _level_mapper_call.original = tv_poolers.LevelMapper.__call__
        tv_poolers.LevelMapper.__call__ = _level_mapper_call
    from torchvision.models.detection import roi_heads as tv_heads
    if tv_heads.keypointrcnn_inference is not _keypointrcnn_inference:      # module-level function: patched process-wide
        _keypointrcnn_inference.original = tv_heads.keypointrcnn_inference
        tv_heads.keypointrcnn_inference = _keypointrcnn_inference


def disable_batched_heads(model) -> None:
    rpn, heads = model.rpn, model.roi_heads
    if hasattr(rpn, '_msq_filter_proposals'):
        rpn.filter_proposals = rpn._msq_filter_proposals
        del rpn._msq_filter_proposals
    if hasattr(heads, '_msq_postprocess_detections'):
        heads.postprocess_detections = heads._msq_postprocess_detections
        del heads._msq_postprocess_detections
    tr = model.transform
    if hasattr(tr, '_msq_forward'):
        tr.forward = tr._msq_forward
        del tr._msq_forward
    if hasattr(heads, '_msq_forward'):
        heads.forward = heads._msq_forward
        del heads._msq_forward
    if hasattr(tr, '_msq_postprocess'):
        tr.postprocess = tr._msq_postprocess
        del tr._msq_postprocess
    from torchvision.ops import poolers as tv_poolers
    if tv_poolers._convert_to_roi_format is _convert_to_roi_format:
        tv_poolers._convert_to_roi_format = _convert_to_roi_format.original
    if tv_poolers._multiscale_roi_align is _multiscale_roi_align:
        tv_poolers._multiscale_roi_align = _multiscale_roi_align.original
    if tv_poolers.LevelMapper.__call__ is _level_mapper_call:
        tv_poolers.LevelMapper.__call__ = _level_mapper_call.original
    from torchvision.models.detection import roi_heads as tv_heads
    if tv_heads.keypointrcnn_inference is _keypointrcnn_inference:
        tv_heads.keypointrcnn_inference = _keypointrcnn_inference.original
