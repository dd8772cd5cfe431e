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
from .. import _lib
from ..proc.kalman import KalmanTracker, KalmanTrackerAngle, KalmanTrackerNPoints2D, KalmanTrackerPoint2D
from ..proc.proc import _gather_instances, _tracked_angles_and_flips, crop_and_rotate_frames_batch
from ..proc.scalars import scalars_from_table
from ..proc.sort_tracker import Detection, Tracker
from .pipeline_step import ProcessPipelineStep


class ProcessFeaturesStep(ProcessPipelineStep):
    def initialize(self):
        self.crop = self.config['crop_size']
        if self.config.get('use_tracking', False):
            # ref: process_features_step.py:41-50 -- centroid + 8 keypoints, and the angle, at order 3
            self.point_tracker = KalmanTracker([KalmanTrackerPoint2D(order=3, delta_t=1.0),
                                                KalmanTrackerNPoints2D(8, order=3, delta_t=1.0)])
            self.angle_tracker = KalmanTracker([KalmanTrackerAngle(order=3, delta_t=1.0, degrees=True)])
        else:
            self.point_tracker = None
            self.angle_tracker = None
        # ref: process_features_step.py:35-38 -- identities of individuals across frames (norfair's configuration, restated)
        self.instance_tracker = Tracker(distance_function='euclidean', distance_threshold=50, initialization_delay=0, hit_counter_max=3)
        self.engine = ChunkEngine()
        self.to_host = bool(self.config.get('results_to_host', True))

    def process(self, data: dict):
        data = self._select_instances(data)
        return self._features_and_crops(data)

    # ---- ref: process_features_step.py:63-113 (mask-IoU NMS); norfair tracking (:140) needs >1 instance --------
    @staticmethod
    def _nms_mask_instances(instances: Instances, iou_threshold: float = 0.5) -> Instances:
        """Mask-IoU suppression exactly as the reference runs it (process_features_step.py:63-113), which is NOT the textbook
        greedy NMS: every round keeps the best remaining instance and then drops EVERY remaining instance that overlaps a
        better-scored remaining one (not only those overlapping the one just kept).  The IoU matrix comes from one matmul on
        the device; the rounds run on its few-by-few host copy."""
        if len(instances) <= 1:
            return instances
        instances = instances[instances.pred_masks.flatten(1).any(dim=1)]
        if len(instances) == 0:
            return instances
        flat = instances.pred_masks.flatten(1).float()
        inter = flat @ flat.T
        area = flat.sum(dim=1)
        iou = (inter / (area[:, None] + area[None, :] - inter)).cpu().numpy()
        idxs = np.argsort(instances.scores.detach().cpu().numpy())           # ascending: the best instance is last
        picked: List[int] = []
        while len(idxs) > 0:
            last = len(idxs) - 1
            picked.append(int(idxs[last]))
            rows = np.where(np.triu(iou[np.ix_(idxs, idxs)], k=1) > iou_threshold)[0]   # the lower-scored member of each pair
            idxs = np.delete(idxs, np.unique(np.concatenate(([last], rows))))
        return instances[picked]

    @staticmethod
    def _instances_to_detections(instances: Instances) -> List[Detection]:
        """ref: process_features_step.py:115-129 -- one Detection per instance at the centre of mass of its mask (row, column),
        the box centre when the mask is empty.  Centres are computed on the device in one pass; the tracker is host-side."""
        n = len(instances)
        if n == 0:
            return []
        masks = instances.pred_masks.float()
        area = masks.sum(dim=(1, 2))
        ys = torch.arange(masks.shape[1], device=masks.device, dtype=torch.float32)
        xs = torch.arange(masks.shape[2], device=masks.device, dtype=torch.float32)
        cy = (masks.sum(dim=2) * ys).sum(dim=1) / area
        cx = (masks.sum(dim=1) * xs).sum(dim=1) / area
        centres = torch.stack([cy, cx], dim=1).cpu().numpy()
        boxes = instances.pred_boxes.get_centers().cpu().numpy()
        out = []
        for i in range(n):
            centre = centres[i] if np.isfinite(centres[i]).all() else boxes[i]
            out.append(Detection(np.asarray(centre, dtype=float), data={'index': i, 'instance': instances[i]}))
        return out

    def _select_instances(self, data: dict) -> dict:
        """ref: process_features_step.py:132-160 -- mask-IoU suppression, then SORT-style identity tracking: when more than one
        tracked object is alive, keep the `expected_instances` OLDEST ones that are live in this frame.  The dense hand-over
        (one detection per frame by construction, TEST.DETECTIONS_PER_IMAGE = 1) has nothing to select."""
        if '_dense_instances' in data:
            return data
        expected = int(self.config.get('expected_instances', 1))
        for frame in data['inference']:
            inst = self._nms_mask_instances(frame['instances'])
            frame['instances'] = inst
            tracked = self.instance_tracker.update(detections=self._instances_to_detections(inst))
            if len(tracked) <= 1:
                continue
            by_age = sorted((t for t in tracked if t.live_points.any()), key=lambda t: t.age)
            selected = []
            while len(selected) < expected and len(by_age) > 0:
                selected.append(by_age.pop().last_detection.data['instance'])
            if selected:
                frame['instances'] = Instances.cat(selected)
            else:
                h, w = inst.image_size
                frame['instances'] = create_empty_instances(w, h, _lib.NUM_KEYPOINTS, device='cuda')
        return data

    # ---- use_tracking=True: the same outputs through the Kalman branch (ref: proc/proc.py:730-826) ---------------
    def _tracked(self, chunk, masks, kpts):
        n, h, w = (int(v) for v in chunk.shape)
        res = self.engine.clean_and_features(chunk, masks)
        centroid, kp64, angles, flips = _tracked_angles_and_flips(res['centroid'], res['orientation_rad'], res['axis_length'],
                                                                  kpts, self.point_tracker, self.angle_tracker)
        e = _dev.empty
        scalars, kcols = e((_lib.NUM_SCALARS, n), torch.float64), e((_lib.NUM_KPT_COLS, n), torch.float64)
        scratch = e((int(_lib.load().msq_scalars_scratch_bytes(n)) + 8,), torch.uint8)
        _lib.call('msq_scalars_and_keypoints_f64', _dev.ptr(chunk), _dev.ptr(masks), _dev.ptr(res['cleaned']), _dev.ptr(centroid),
                  _dev.ptr(angles), _dev.ptr(res['axis_length']), _dev.ptr(kp64), n, h, w, max(n, 1),
                  float(self.config['min_height']), float(self.config['max_height']), float(self.config['true_depth']),
                  _dev.ptr(scalars), _dev.ptr(kcols), _dev.ptr(scratch), scratch.numel(), _dev.stream())
        depth, mask_crops = crop_and_rotate_frames_batch(chunk, centroid, angles, self.crop, frames2=masks)
        out = {'cleaned': res['cleaned'], 'centroid': centroid, 'angle_deg': angles, 'axis_length': res['axis_length'],
               'flips': flips, 'scalars': scalars, 'kpt_cols': kcols, 'depth_crops': depth, 'mask_crops': mask_crops}
        return out, kp64

    # ---- ref: process_features_step.py:163-199 ----------------------------------------------------------
    def _features_and_crops(self, data: dict) -> dict:
        chunk = _dev.as_device(data['chunk'], torch.uint8)
        if '_dense_instances' in data:
            masks, kpts, ninst = data.pop('_dense_instances')
        else:
            masks, kpts, ninst = _gather_instances(data['inference'])
        if isinstance(ninst, torch.Tensor):
            ninst = ninst.cpu().numpy()
        n = int(chunk.shape[0])
        if self.point_tracker is not None:
            res, kpts = self._tracked(chunk, masks, kpts)
        else:
            res = self.engine.extract(chunk, masks, kpts, chunk_size=max(n, 1), min_height=self.config['min_height'],
                                      max_height=self.config['max_height'], true_depth=self.config['true_depth'],
                                      crop_size=self.crop, positive_bits=data.get('positive_bits'))
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
