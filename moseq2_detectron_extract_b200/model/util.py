"""Model-output helpers on the hot path (mirrors reference model/util.py:45-76)."""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from .. import _dev, _lib
from .instances import Boxes, Instances


def paste_masks(soft_masks, boxes, height: int, width: int, threshold: float = 0.5):
    """Paste (n, M, M) soft masks into (n, height, width) bool images, detectron2's
    `paste_masks_in_image` as used by `detector_postprocess` (ref: model/util.py:59).  CUDA kernel
    csrc/paste.cu; numpy in -> numpy out, CUDA tensors in -> CUDA tensor out."""
    soft = _dev.as_device(soft_masks, torch.float32)
    if soft.dim() == 4:                                 # (n, 1, M, M) as the mask head emits it
        soft = soft[:, 0]
    soft = soft.contiguous()
    bx = _dev.as_device(boxes.tensor if isinstance(boxes, Boxes) else boxes, torch.float32)
    n, M = int(soft.shape[0]), int(soft.shape[-1])
    out = _dev.empty((n, int(height), int(width)), torch.uint8)
    _lib.call('msq_paste_masks', _dev.ptr(soft), _dev.ptr(bx), n, M, int(height), int(width), float(threshold),
              _dev.ptr(out), _dev.stream())
    return _dev.give_back(out.to(torch.bool), soft_masks)


def detector_postprocess(results: Instances, output_height: int, output_width: int, mask_threshold: float = 0.5) -> Instances:
    """detectron2.modeling.postprocessing.detector_postprocess restated for the fields this model emits:
    rescale + clip boxes, drop empty ones, paste the soft masks, rescale keypoints (ref: model/util.py:59)."""
    in_h, in_w = results.image_size
    sx, sy = float(output_width) / in_w, float(output_height) / in_h
    out = Instances((int(output_height), int(output_width)), **results.get_fields())
    boxes = Boxes(out.pred_boxes.tensor.clone()) if isinstance(out.pred_boxes, Boxes) else Boxes(out.pred_boxes.clone())
    boxes.scale(sx, sy)
    boxes.clip(out.image_size)
    out.set('pred_boxes', boxes)
    keep = boxes.nonempty()
    out = out[keep]
    if out.has('pred_masks'):
        masks = out.pred_masks
        if masks.dim() == 4 or masks.dtype != torch.bool:      # still the soft (n,1,M,M) head output
            if len(out) > 0:
                out.set('pred_masks', paste_masks(masks.cuda() if not masks.is_cuda else masks, out.pred_boxes,
                                                  output_height, output_width, mask_threshold))
            else:
                out.set('pred_masks', torch.zeros((0, int(output_height), int(output_width)), dtype=torch.bool,
                                                  device=masks.device))
    if out.has('pred_keypoints'):
        kp = out.pred_keypoints.clone()
        kp[:, :, 0] *= sx
        kp[:, :, 1] *= sy
        out.set('pred_keypoints', kp)
    return out


def outputs_to_instances(inputs: List[Dict[str, torch.Tensor]], outputs: List[Dict[str, torch.Tensor]]) -> List[dict]:
    """TorchScript model outputs (list of dicts of tensors) -> [{'instances': Instances}] (ref: model/util.py:45-62)."""
    instances = []
    for i, o in zip(inputs, outputs):
        height = int(i['height']) if 'height' in i else int(i['image'].shape[-2])
        width = int(i['width']) if 'width' in i else int(i['image'].shape[-1])
        o = dict(o)
        ins = Instances((height, width), pred_boxes=Boxes(o.pop('pred_boxes')), **o)
        instances.append({'instances': detector_postprocess(ins, height, width)})
    return instances


def create_empty_instances(width: int, height: int, nkeypoints: int, device='cpu') -> Instances:
    """An `Instances` holding zero detections (ref: model/util.py:65-76)."""
    return Instances(
        (height, width),
        pred_boxes=Boxes(torch.empty((0, 4), dtype=torch.float32, device=device)),
        scores=torch.empty((0,), dtype=torch.float32, device=device),
        pred_classes=torch.empty((0,), dtype=torch.int64, device=device),
        pred_masks=torch.empty((0, height, width), dtype=torch.bool, device=device),
        pred_keypoints=torch.empty((0, nkeypoints, 3), dtype=torch.float32, device=device),
        pred_keypoints_heatmaps=torch.empty((0, nkeypoints, 28, 28), dtype=torch.float32, device=device),
    )
