"""Predictor (mirrors reference model/predict.py:12-106).

The R-CNN itself is a dense graph run by PyTorch (tensor cores via cuDNN/cuBLAS): either the repo's own
TorchScript export (`from_torchscript`, ref: model/predict.py:47-51) or a randomly initialised
Keypoint+Mask R-CNN R50-FPN built from torchvision parts with the reference's head layout
(`from_random_init`; ref: model/config.py:21-94 -- 1 class, 8 keypoints, 240/250 px, no resize).
Everything around it is ours: intensity scaling + channel replication (csrc/prep.cu), mask pasting
(csrc/paste.cu) and the `Instances` container; detections never leave the GPU.
"""
from __future__ import annotations

from contextlib import ExitStack
from typing import Any, Dict, List

import numpy as np
import torch

from .. import _dev, _lib
from .util import outputs_to_instances


def build_random_keypoint_mask_rcnn(num_keypoints: int = 8, image_size: int = 250, detections_per_img: int = 1,
                                    rpn_post_nms_top_n: int = 100):
    """Keypoint + Mask R-CNN with a ResNet-50-FPN backbone and random weights (no network access needed)."""
    import torchvision
    from torchvision.models.detection import MaskRCNN
    from torchvision.models.detection.backbone_utils import resnet_fpn_backbone
    from torchvision.models.detection.keypoint_rcnn import KeypointRCNNHeads, KeypointRCNNPredictor
    from torchvision.ops import MultiScaleRoIAlign

    backbone = resnet_fpn_backbone(backbone_name='resnet50', weights=None, trainable_layers=5)
    model = MaskRCNN(backbone, num_classes=2, min_size=image_size, max_size=image_size,
                     image_mean=[103.53, 116.28, 123.675], image_std=[57.375, 57.12, 58.395],
                     rpn_pre_nms_top_n_test=1000, rpn_post_nms_top_n_test=rpn_post_nms_top_n,
                     box_detections_per_img=detections_per_img, box_score_thresh=0.0)
    model.roi_heads.keypoint_roi_pool = MultiScaleRoIAlign(featmap_names=['0', '1', '2', '3'], output_size=14, sampling_ratio=2)
    model.roi_heads.keypoint_head = KeypointRCNNHeads(backbone.out_channels, tuple(512 for _ in range(8)))
    model.roi_heads.keypoint_predictor = KeypointRCNNPredictor(512, num_keypoints)

    # keep the raw 28x28 mask probabilities: pasting is done by our kernel (detector_postprocess)
    transform = model.transform
    original_postprocess = transform.postprocess

    def postprocess_keep_soft_masks(result, image_shapes, original_image_sizes):
        soft = [r.pop('masks') if 'masks' in r else None for r in result]
        result = original_postprocess(result, image_shapes, original_image_sizes)
        for r, m in zip(result, soft):
            if m is not None:
                r['masks'] = m
        return result

    transform.postprocess = postprocess_keep_soft_masks
    transform._msq_keep_soft_masks = True                  # tells model/batched_heads.py its batched postprocess may skip pasting
    return model


def fold_batchnorm_into_convs(module: torch.nn.Module) -> int:
    """Inference-time folding of every (Conv2d, BatchNorm2d / FrozenBatchNorm2d) pair registered next to each other in a
    module (ResNet stem, bottleneck conv1-3, downsample branches): w' = w * s, b' = (b - mean) * s + beta with
    s = gamma / sqrt(var + eps).  The normalisation layer becomes an Identity, which removes one or two full-tensor
    elementwise kernels per convolution (tools/rcnn_kernels.py: as much device time as the convolutions themselves).
    Returns the number of pairs folded; only valid for eval-mode models."""
    from torchvision.ops.misc import FrozenBatchNorm2d
    folded = 0
    for parent in module.modules():
        names = list(parent._modules.keys())
        for a, b in zip(names, names[1:]):
            conv, bn = parent._modules[a], parent._modules[b]
            if not isinstance(conv, torch.nn.Conv2d) or not isinstance(bn, (torch.nn.BatchNorm2d, FrozenBatchNorm2d)):
                continue
            if isinstance(bn, torch.nn.BatchNorm2d) and (bn.training or bn.running_var is None):
                continue
            with torch.no_grad():
                gamma = bn.weight if bn.weight is not None else torch.ones_like(bn.running_var)
                beta = bn.bias if bn.bias is not None else torch.zeros_like(bn.running_var)
                scale = gamma * torch.rsqrt(bn.running_var + bn.eps)
                bias = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
                conv.weight.mul_(scale.reshape(-1, 1, 1, 1))
                conv.bias = torch.nn.Parameter((bias - bn.running_mean) * scale + beta, requires_grad=False)
            parent._modules[b] = torch.nn.Identity()
            folded += 1
    return folded


class _TorchvisionAdapter(torch.nn.Module):
    """Gives a torchvision detector the I/O contract of the reference's TorchScript export
    (ref: model/deploy.py:73-102): list of {'image': CHW} in, list of dicts with pred_* keys out."""

    input_format = 'RGB'

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, inputs: List[Dict[str, torch.Tensor]]):
        outs = self.model([i['image'] for i in inputs])
        results = []
        for o in outs:
            results.append({'pred_boxes': o['boxes'], 'scores': o['scores'], 'pred_classes': o['labels'] - 1,
                            'pred_masks': o['masks'], 'pred_keypoints': o['keypoints']})
        return results


class Predictor:
    def __init__(self, model: Any, is_torchscript: bool = False, amp: bool = False):
        self.model = model
        self.is_torchscript = is_torchscript
        self.amp = amp                      # bf16 autocast for the dense graph (tensor cores)
        self.exit_stack = ExitStack()
        self.exit_stack.enter_context(torch.no_grad())

    @property
    def device(self):
        return next(self.model.parameters()).device

    @classmethod
    def from_torchscript(cls, path: str):
        _dev.require_cuda()
        model = torch.jit.load(path, map_location='cuda')
        model.eval()
        return cls(model, is_torchscript=True)

    @classmethod
    def from_random_init(cls, device: str = 'cuda', seed: int = 0, amp: bool = False, batched_heads: bool = True, **kwargs):
        _dev.require_cuda()
        torch.manual_seed(seed)
        fold_bn = kwargs.pop('fold_batchnorm', True)
        fused_convs = kwargs.pop('fused_convs', True)
        model = _TorchvisionAdapter(build_random_keypoint_mask_rcnn(**kwargs)).to(device).eval()
        if fold_bn:
            fold_batchnorm_into_convs(model.model.backbone)
        if fused_convs:                 # conv + bias + ReLU (+ residual) as single cuDNN calls
            from .fused_convs import enable_fused_convs
            enable_fused_convs(model.model)
        if batched_heads:               # one segmented NMS launch per batch instead of torchvision's per-image loops
            from .batched_heads import enable_batched_heads
            enable_batched_heads(model.model)
        if amp:
            model = model.to(memory_format=torch.channels_last)
        return cls(model, is_torchscript=True, amp=amp)

    # ---- reference-shaped entry point (ref: model/predict.py:53-106) ---------------------------------
    def __call__(self, original_image):
        single = original_image.ndim == 3
        if single:
            original_image = original_image[None]
        img = _dev.as_device(original_image)                     # (N, H, W, C) uint8
        if img.shape[3] == 1:
            img = img.expand(-1, -1, -1, 3)
        chw = img.permute(0, 3, 1, 2).to(torch.float32).contiguous()
        preds = self._forward(chw)
        return preds[0] if single else preds

    # ---- device fast path used by InferenceStep: scale + replicate + CHW in one kernel -----------------
    def _fused_input_ok(self) -> bool:
        """The one-kernel input path serves the torchvision detectors built here (a chunk's images all have one size)."""
        if not isinstance(self.model, _TorchvisionAdapter):
            return False
        tr = self.model.model.transform
        return bool(getattr(tr, '_msq_keep_soft_masks', False)) and tr.fixed_size is None

    def predict_prepared(self, chunk_u8: torch.Tensor, vmin, vmax) -> List[dict]:
        n, h, w = (int(v) for v in chunk_u8.shape)
        if self._fused_input_ok():
            with torch.no_grad():
                with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.amp):
                    dets = self._detect_from_chunk(chunk_u8, vmin, vmax)
                outputs = [{'pred_boxes': o['boxes'].float(), 'scores': o['scores'].float(), 'pred_classes': o['labels'] - 1,
                            'pred_masks': o['masks'].float(), 'pred_keypoints': o['keypoints'].float()} for o in dets]
                size = {'height': h, 'width': w}
                return outputs_to_instances([size] * n, outputs)
        chw = _dev.empty((n, 3, h, w), torch.float32)
        _lib.call('msq_scale_frames_chw3_f32', _dev.ptr(chunk_u8), _dev.ptr(chw), n, h, w, float(vmin), float(vmax),
                  int(isinstance(vmin, (int, np.integer))), _dev.stream())
        return self._forward(chw)

    # ---- dense fast path: what the feature step needs from the detections, without per-image Instances ----------
    def predict_dense(self, chunk_u8: torch.Tensor, vmin, vmax):
        """`predict_prepared` + `detector_postprocess` + `mask_and_keypoints_from_model_output` (ref: model/util.py:45-62,
        proc/proc.py:657-685) for the FIRST instance of every frame, batched: returns (masks (n,h,w) u8, keypoints (n,K,3)
        f32 with NaN where a frame has no instance, num_instances (n,) int64 on the device).  One concatenation per field,
        one clip, ONE mask-paste launch for the whole batch -- the per-image form costs ~0.6 ms of launch latency a frame."""
        n, h, w = (int(v) for v in chunk_u8.shape)
        fused_input = self._fused_input_ok()
        if not fused_input:
            chw = _dev.empty((n, 3, h, w), torch.float32)
            _lib.call('msq_scale_frames_chw3_f32', _dev.ptr(chunk_u8), _dev.ptr(chw), n, h, w, float(vmin), float(vmax),
                      int(isinstance(vmin, (int, np.integer))), _dev.stream())
        with torch.no_grad():
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.amp):
                if fused_input:                                          # scale + normalise + resize + pad in ONE kernel
                    outputs = [{'scores': o['scores'], 'pred_boxes': o['boxes'], 'pred_masks': o['masks'], 'pred_keypoints': o['keypoints']}
                               for o in self._detect_from_chunk(chunk_u8, vmin, vmax)]
                elif isinstance(self.model, _TorchvisionAdapter):        # straight to the detector: no per-image dicts
                    outputs = [{'scores': o['scores'], 'pred_boxes': o['boxes'], 'pred_masks': o['masks'], 'pred_keypoints': o['keypoints']}
                               for o in self.model.model(list(chw.unbind(0)))]
                else:
                    outputs = self.model([{'image': chw[i]} for i in range(n)])
            counts = [int(o['scores'].shape[0]) for o in outputs]                  # host-known sizes: no synchronisation
            have = [i for i, c in enumerate(counts) if c > 0]
            k = int(outputs[have[0]]['pred_keypoints'].shape[1]) if have else _lib.NUM_KEYPOINTS
            dev = chunk_u8.device
            masks = torch.zeros((n, h, w), dtype=torch.uint8, device=dev)
            kpts = torch.full((n, k, 3), float('nan'), dtype=torch.float32, device=dev)
            ninst = torch.zeros((n,), dtype=torch.int64, device=dev)
            if have:
                sel = torch.tensor(have, device=dev)
                # first instance of every frame that has one: one concatenation per field, then one row gather
                first = torch.tensor(np.cumsum([0] + counts[:-1])[have], device=dev)
                boxes = torch.cat([o['pred_boxes'] for o in outputs])[first].float()
                soft = torch.cat([o['pred_masks'] for o in outputs])[first].float()
                kp = torch.cat([o['pred_keypoints'] for o in outputs])[first].float()
                if soft.dim() == 4:
                    soft = soft[:, 0]
                # detector_postprocess at scale 1: clip to the image, keep non-empty boxes, paste at 0.5
                boxes[:, 0::2] = boxes[:, 0::2].clamp(0, w)
                boxes[:, 1::2] = boxes[:, 1::2].clamp(0, h)
                ok = ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)
                pasted = _dev.empty((len(have), h, w), torch.uint8)
                soft = soft.contiguous()
                _lib.call('msq_paste_masks', _dev.ptr(soft), _dev.ptr(boxes.contiguous()), len(have), int(soft.shape[-1]), h, w, 0.5,
                          _dev.ptr(pasted), _dev.stream())
                # frames whose first box is empty fall back to their next surviving instance in the reference; with one
                # detection per image (the extract configuration) they simply have no instance
                masks[sel] = pasted * ok[:, None, None].to(torch.uint8)
                kpts[sel] = torch.where(ok[:, None, None], kp, torch.full_like(kp, float('nan')))
                totals = torch.tensor([counts[i] for i in have], device=dev)
                ninst[sel] = torch.where(ok, totals, totals - 1)
        return masks, kpts, ninst

    def detector_input(self, chunk_u8: torch.Tensor, vmin, vmax):
        """What GeneralizedRCNNTransform.forward would hand the backbone for this chunk -- scaled, replicated to 3 channels,
        normalised, resized, zero-padded to the stride -- from ONE kernel (`msq_detector_input`), channels-last, bf16 under
        autocast.  Returns (tensor (n, 3, ph, pw), (oh, ow))."""
        import ctypes
        import math
        tr = self.model.model.transform
        n, h, w = (int(v) for v in chunk_u8.shape)
        scale = min(tr.min_size[-1] / min(h, w), tr.max_size / max(h, w))
        oh, ow = int(math.floor(h * scale)), int(math.floor(w * scale))       # interpolate(recompute_scale_factor=True)
        div = int(tr.size_divisible)
        ph, pw = -(-oh // div) * div, -(-ow // div) * div
        dtype = torch.bfloat16 if self.amp else torch.float32
        x = torch.empty((n, 3, ph, pw), dtype=dtype, device=chunk_u8.device, memory_format=torch.channels_last)
        _lib.call('msq_detector_input', _dev.ptr(chunk_u8), _dev.ptr(x), int(self.amp), n, h, w, oh, ow, ph, pw,
                  (ctypes.c_float * 3)(*[float(v) for v in tr.image_mean]), (ctypes.c_float * 3)(*[float(v) for v in tr.image_std]),
                  float(vmin), float(vmax), int(isinstance(vmin, (int, np.integer))), _dev.stream())
        return x, (oh, ow)

    def _detect_from_chunk(self, chunk_u8: torch.Tensor, vmin, vmax):
        """GeneralizedRCNN.forward (eval) with the transform replaced by `detector_input`."""
        from collections import OrderedDict
        from torchvision.models.detection.image_list import ImageList
        net = self.model.model
        n, h, w = (int(v) for v in chunk_u8.shape)
        x, size = self.detector_input(chunk_u8.contiguous(), vmin, vmax)
        images = ImageList(x, [size] * n)
        features = net.backbone(x)
        if isinstance(features, torch.Tensor):
            features = OrderedDict([('0', features)])
        proposals, _ = net.rpn(images, features, None)
        detections, _ = net.roi_heads(features, proposals, images.image_sizes, None)
        return net.transform.postprocess(detections, images.image_sizes, [(h, w)] * n)

    def _forward(self, chw: torch.Tensor) -> List[dict]:
        with torch.no_grad():
            inputs = [{'image': chw[i], 'height': torch.tensor(chw.shape[2]), 'width': torch.tensor(chw.shape[3])}
                      for i in range(chw.shape[0])]
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=self.amp):
                outputs = self.model(inputs)
            outputs = [{k: (v.float() if v.is_floating_point() else v) for k, v in o.items()} for o in outputs]
            return outputs_to_instances(inputs, outputs)
