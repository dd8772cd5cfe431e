"""Predictor (mirrors reference model/predict.py:12-106).

The R-CNN is the repo's own graph (model/rcnn.py), built to the reference's detectron2 configuration (ref:
model/config.py:21-94) and exported to TorchScript with the I/O contract of ref model/deploy.py:65-110:

    Predictor.from_torchscript(path)       load a `model.ts` (ref: model/predict.py:47-51).  A graph exported by this repo
                                           (`model.rcnn.export_torchscript`) also carries the batched `forward_dense` entry
                                           the extract pipeline uses; a foreign `.ts` is served through the reference-shaped
                                           per-image path only.
    Predictor.from_random_init(...)        random weights of that architecture (BASELINE: no network, no checkpoints)
    Predictor.from_detectron2_state_dict   tensors of a detectron2 checkpoint (FrozenBN folded, heads re-laid out)

Everything around the graph is ours: intensity scaling + channel replication + normalisation + padding in one kernel
(csrc/prep.cu), batched mask pasting (csrc/paste.cu) and the `Instances` container; detections never leave the GPU.
"""
from __future__ import annotations

from contextlib import ExitStack
from typing import Any, Dict, List

import numpy as np
import torch

from .. import _dev, _lib
from .util import outputs_to_instances


class Predictor:
    def __init__(self, model: Any, is_torchscript: bool = False):
        self.model = model
        self.is_torchscript = is_torchscript
        self.exit_stack = ExitStack()
        self.exit_stack.enter_context(torch.no_grad())
        self.has_dense_entry = hasattr(model, 'forward_dense')

    @property
    def device(self):
        return next(self.model.parameters()).device

    # ---- constructors ---------------------------------------------------------------------------------------------------
    @classmethod
    def from_torchscript(cls, path: str):
        """ref: model/predict.py:47-51.  The repo's own exports name `torch.ops.msq.*`: register them before loading."""
        _dev.require_cuda()
        from . import ops  # noqa: F401
        model = torch.jit.load(path, map_location='cuda')
        model.eval()
        return cls(model, is_torchscript=True)

    @classmethod
    def from_random_init(cls, device: str = 'cuda', seed: int = 0, dtype: torch.dtype = torch.bfloat16, scripted: bool = False, **kwargs):
        """Random-initialised Keypoint + Mask R-CNN R50-FPN of the reference configuration.  kwargs go to `MoseqRCNN`
        (e.g. post_nms_topk=100).  scripted=True runs the graph through torch.jit.script, as an exported model would."""
        _dev.require_cuda()
        from . import rcnn
        model = rcnn.build_random(seed=seed, dtype=dtype, device=device, **kwargs)
        if scripted:
            model = torch.jit.script(model)
        return cls(model, is_torchscript=True)

    @classmethod
    def from_detectron2_state_dict(cls, state: Dict[str, Any], device: str = 'cuda', dtype: torch.dtype = torch.bfloat16, **kwargs):
        _dev.require_cuda()
        from . import rcnn
        return cls(rcnn.from_detectron2_state_dict(state, dtype=dtype, device=device, **kwargs), is_torchscript=True)

    # ---- reference-shaped entry point (ref: model/predict.py:53-106) ---------------------------------
    def __call__(self, original_image):
        single = original_image.ndim == 3
        if single:
            original_image = original_image[None]
        img = _dev.as_device(original_image)                     # (N, H, W, C) uint8
        if img.shape[3] == 1:                                    # ref :74-77: grey -> 3 identical channels
            img = img.expand(-1, -1, -1, 3)
        chw = img.permute(0, 3, 1, 2).contiguous()
        preds = self._forward(chw)
        return preds[0] if single else preds

    def _forward(self, chw: torch.Tensor) -> List[dict]:
        with torch.no_grad():
            inputs = [{'image': chw[i], 'height': torch.tensor(chw.shape[2]), 'width': torch.tensor(chw.shape[3])}
                      for i in range(chw.shape[0])]
            outputs = self.model(inputs)
            outputs = [{k: (v.float() if v.is_floating_point() else v) for k, v in o.items()} for o in outputs]
            return outputs_to_instances(inputs, outputs)

    # ---- chunk entry used by InferenceStep: prepared uint8 frames -> [{'instances': Instances}] ---------------------------
    def predict_prepared(self, chunk_u8: torch.Tensor, vmin, vmax) -> List[dict]:
        n, h, w = (int(v) for v in chunk_u8.shape)
        if self.has_dense_entry:
            with torch.no_grad():
                boxes, scores, has, soft, kpts, heat = self.model.forward_dense(chunk_u8.contiguous(), float(vmin), float(vmax),
                                                                                 isinstance(vmin, (int, np.integer)))
                counts = has.to(torch.int64).cpu().tolist()
                outputs = [{'pred_boxes': boxes[i:i + c], 'scores': scores[i:i + c],
                            'pred_classes': torch.zeros((c,), dtype=torch.int64, device=boxes.device), 'pred_masks': soft[i:i + c],
                            'pred_keypoints': kpts[i:i + c], 'pred_keypoint_heatmaps': heat[i:i + c]} for i, c in enumerate(counts)]
                return outputs_to_instances([{'height': h, 'width': w}] * n, outputs)
        scaled = _dev.empty((n, h, w), torch.uint8)
        _lib.call('msq_scale_frames', _dev.ptr(chunk_u8.contiguous()), _dev.ptr(scaled), n * h * w, float(vmin), float(vmax),
                  int(isinstance(vmin, (int, np.integer))), _dev.stream())
        return self._forward(scaled[:, None].expand(-1, 3, -1, -1).contiguous())

    # ---- dense fast path: what the feature step needs from the detections, without per-image Instances ----------
    def predict_dense(self, chunk_u8: torch.Tensor, vmin, vmax):
        """`predict_prepared` + `detector_postprocess` + `mask_and_keypoints_from_model_output` (ref: model/util.py:45-62,
        proc/proc.py:657-685) for the detection of every frame, batched: returns (masks (n,h,w) u8, keypoints (n,K,3) f32 with
        NaN where a frame has no instance, num_instances (n,) int64), all on the device, no host synchronisation.  One graph
        call, one clip, ONE mask-paste launch for the whole batch."""
        n, h, w = (int(v) for v in chunk_u8.shape)
        if not self.has_dense_entry:
            return self._dense_from_instances(self.predict_prepared(chunk_u8, vmin, vmax), n, h, w)
        with torch.no_grad():
            boxes, scores, has, soft, kpts, _ = self.model.forward_dense(chunk_u8.contiguous(), float(vmin), float(vmax),
                                                                          isinstance(vmin, (int, np.integer)))
            # detector_postprocess at scale 1: clip to the image (done by the graph), keep non-empty boxes, paste at 0.5
            ok = (has != 0) & ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)
            pasted = _dev.empty((n, h, w), torch.uint8)
            soft = soft.reshape(n, soft.shape[-2], soft.shape[-1]).float().contiguous()
            _lib.call('msq_paste_masks', _dev.ptr(soft), _dev.ptr(boxes.contiguous()), n, int(soft.shape[-1]), h, w, 0.5,
                      _dev.ptr(pasted), _dev.stream())
            masks = pasted * ok[:, None, None].to(torch.uint8)
            kpts = torch.where(ok[:, None, None], kpts, torch.full_like(kpts, float('nan')))
            return masks, kpts, ok.to(torch.int64)

    @staticmethod
    def _dense_from_instances(outputs: List[dict], n: int, h: int, w: int):
        masks = torch.zeros((n, h, w), dtype=torch.uint8, device='cuda')
        kpts = torch.full((n, _lib.NUM_KEYPOINTS, 3), float('nan'), dtype=torch.float32, device='cuda')
        ninst = torch.zeros((n,), dtype=torch.int64, device='cuda')
        for i, o in enumerate(outputs):
            ins = o['instances']
            if len(ins) > 0:
                masks[i] = ins.pred_masks[0].to(torch.uint8)
                kpts[i] = ins.pred_keypoints[0]
                ninst[i] = len(ins)
        return masks, kpts, ninst
