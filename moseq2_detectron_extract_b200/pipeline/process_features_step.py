"""Feature step (ref: pipeline/process_features_step.py:21-199): instance selection, moment features, flips,
angle filter, scalars, keypoint table and the egocentric crops -- one `msq_extract_chunk` call per chunk."""
from typing import List

import numpy as np
import torch

from .. import _dev
from ..engine import ChunkEngine
from ..model.instances import Instances
from ..model.util import create_empty_instances
from ..proc.keypoints import keypoints_from_table
from ..proc.proc import _gather_instances
from ..proc.scalars import scalars_from_table
from .pipeline_step import ProcessPipelineStep


class ProcessFeaturesStep(ProcessPipelineStep):
    def initialize(self):
        self.crop = self.config['crop_size']
        if self.config.get('use_tracking', False):
            raise NotImplementedError('ProcessFeaturesStep: use_tracking=True (Kalman branch, ref proc/kalman.py) is out of '
                                      'scope of this build (SURVEY.md section 8f row f1); run with use_tracking=False')
        self.engine = ChunkEngine()
        self.to_host = bool(self.config.get('results_to_host', True))

    def process(self, data: dict):
        data = self._select_instances(data)
        return self._features_and_crops(data)

    # ---- ref: process_features_step.py:63-113 (mask-IoU NMS); norfair tracking (:140) needs >1 instance --------
    @staticmethod
    def _nms_mask_instances(instances: Instances, iou_threshold: float = 0.5) -> Instances:
        if len(instances) <= 1:
            return instances
        keep_nonempty = instances.pred_masks.flatten(1).any(dim=1)
        instances = instances[keep_nonempty]
        order = torch.argsort(instances.scores, descending=True).tolist()
        flat = instances.pred_masks.flatten(1).float()
        inter = flat @ flat.T
        area = flat.sum(dim=1)
        iou = inter / (area[:, None] + area[None, :] - inter)
        picked: List[int] = []
        alive = set(order)
        for i in order:
            if i not in alive:
                continue
            picked.append(i)
            for j in list(alive):
                if j != i and float(iou[i, j]) > iou_threshold:
                    alive.discard(j)
            alive.discard(i)
        return instances[picked]

    def _select_instances(self, data: dict) -> dict:
        if '_dense_instances' in data:
            return data
        for frame in data['inference']:
            inst = self._nms_mask_instances(frame['instances'])
            if len(inst) > self.config['expected_instances']:
                top = torch.argsort(inst.scores, descending=True)[: self.config['expected_instances']]
                inst = inst[top.tolist()]
            frame['instances'] = inst
        return data

    # ---- ref: process_features_step.py:163-199 ----------------------------------------------------------
    def _features_and_crops(self, data: dict) -> dict:
        chunk = _dev.as_device(data['chunk'], torch.uint8)
        if '_dense_instances' in data:
            masks, kpts, ninst = data.pop('_dense_instances')
        else:
            masks, kpts, ninst = _gather_instances(data['inference'])
        n = int(chunk.shape[0])
        res = self.engine.extract(chunk, masks, kpts, chunk_size=max(n, 1), min_height=self.config['min_height'],
                                  max_height=self.config['max_height'], true_depth=self.config['true_depth'],
                                  crop_size=self.crop)
        for i in np.flatnonzero(np.asarray(ninst) <= 0):
            self.write_message(f"WARN: No instances found for frame {data['frame_idxs'][i]}")
        conv = (lambda t: t.cpu().numpy()) if self.to_host else (lambda t: t.clone())
        features = {
            'cleaned_frames': conv(res['cleaned']), 'masks': conv(masks),
            'features': {'centroid': conv(res['centroid']), 'orientation': conv(res['angle_deg']),
                         'axis_length': conv(res['axis_length']), 'contour': []},
            'flips': conv(res['flips'].to(torch.bool)), 'keypoints': conv(kpts.to(torch.float64)),
            'num_instances': np.asarray(ninst),
        }
        data['features'] = features
        data['scalars'] = scalars_from_table(conv(res['scalars']))
        data['keypoints'] = keypoints_from_table(conv(res['kpt_cols']))
        data['depth_frames'] = conv(res['depth_crops'])
        data['mask_frames'] = conv(res['mask_crops'])
        self.update_progress(n)
        return data
