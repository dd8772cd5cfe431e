"""Inference step (ref: pipeline/inference_step.py:16-72): intensity-scale the chunk batch by batch and run
the Keypoint/Mask R-CNN.  Unlike the reference the Instances stay on the GPU (no `.to('cpu')`, :68)."""
import os

import torch

from .. import _dev, _lib
from ..model.instances import Boxes, Instances
from ..model.predict import Predictor
from .pipeline_step import ProcessPipelineStep


class InferenceStep(ProcessPipelineStep):
    def initialize(self):
        model_path = self.config['model']
        self.write_message('Loading model....')
        if isinstance(model_path, str) and os.path.isfile(model_path) and model_path.endswith('.ts'):
            self.predictor = Predictor.from_torchscript(model_path)
        elif model_path is None or model_path == 'random':
            import torch as _torch
            kw = {k: self.config[k] for k in ('post_nms_topk', 'pre_nms_topk') if k in self.config}
            self.predictor = Predictor.from_random_init(device=self.config.get('device', 'cuda'),
                                                        dtype=_torch.bfloat16 if self.config.get('amp', True) else _torch.float32,
                                                        scripted=bool(self.config.get('scripted', False)), **kw)
        else:
            raise NotImplementedError('InferenceStep: only TorchScript (.ts) models or model="random" are supported; '
                                      'detectron2 checkpoints need detectron2 (not part of this build)')
        self.write_message(f' -> Actually using device "{self.predictor.device}"')

    def process(self, data):
        raw = _dev.as_device(data['chunk'], torch.uint8)
        n = int(raw.shape[0])
        batch_size = min(self.config['batch_size'], n)
        if self.config.get('dense_inference', False) and self.config.get('expected_instances', 1) == 1:
            # batched hand-over of the first instance of every frame (what ProcessFeaturesStep consumes); the reference's
            # data['inference'] list of Instances is not built
            parts = []
            for i in range(0, n, batch_size):
                parts.append(self.predictor.predict_dense(raw[i:i + batch_size], self.config['min_height'], self.config['max_height']))
                self.update_progress(min(batch_size, n - i))
            data['_dense_instances'] = tuple(torch.cat([p[j] for p in parts]) for j in range(3))
            data['inference'] = None
            return data
        outputs = []
        for i in range(0, n, batch_size):
            pred = self.predictor.predict_prepared(raw[i:i + batch_size], self.config['min_height'], self.config['max_height'])
            outputs.extend(pred)
            self.update_progress(min(batch_size, n - i))
        data['inference'] = outputs
        return data


class SyntheticInferenceStep(ProcessPipelineStep):
    """Stand-in for InferenceStep in the no-R-CNN configuration (BASELINE.json configs[1]): turns the
    synthetic session's ground-truth instances into the same `data['inference']` structure."""

    def process(self, data):
        gt = data['synthetic_instances']
        masks = _dev.as_device(gt.masks).to(torch.bool)
        kpts = _dev.as_device(gt.keypoints, torch.float32)
        h, w = int(masks.shape[1]), int(masks.shape[2])
        outputs = []
        for i in range(int(masks.shape[0])):
            if gt.num_instances[i] > 0:
                inst = Instances((h, w), pred_boxes=Boxes(torch.tensor([[0., 0., float(w), float(h)]], device='cuda')),
                                 scores=torch.ones((1,), device='cuda'), pred_classes=torch.zeros((1,), dtype=torch.int64, device='cuda'),
                                 pred_masks=masks[i:i + 1], pred_keypoints=kpts[i:i + 1])
            else:
                from ..model.util import create_empty_instances
                inst = create_empty_instances(w, h, _lib.NUM_KEYPOINTS, device='cuda')
            outputs.append({'instances': inst})
        data['inference'] = outputs
        data['_dense_instances'] = (masks.to(torch.uint8), kpts, gt.num_instances)   # fast path for ProcessFeaturesStep
        self.update_progress(int(masks.shape[0]))
        return data
