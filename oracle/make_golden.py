"""TEST INFRASTRUCTURE -- generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in a container that has /root/reference:   python oracle/make_golden.py
Inputs are NOT stored (they are regenerated bit-identically from the seeds by
moseq2_detectron_extract_b200.synthetic); only the reference's outputs are.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_import  # noqa: E402
from moseq2_detectron_extract_b200 import synthetic  # noqa: E402

CASES = {
    # name: (geometry, generate_chunk kwargs, background kwargs)
    'kinect_clean': ('kinect_v2', dict(n_frames=48, seed=0, t0=90), dict(dtype='float32')),
    'kinect_missing_holes': ('kinect_v2', dict(n_frames=40, seed=1, t0=300, missing_every=13, mask_holes=True),
                             dict(dtype='float64', half_steps=True)),
    'kinect_invalid': ('kinect_v2', dict(n_frames=6, seed=2, t0=10, invalid_rate=0.002), dict(dtype='uint16')),
    'azure_clean': ('azure', dict(n_frames=8, seed=3, t0=40), dict(dtype='float32')),
}


class FakeInstances:
    """Duck-type of detectron2 Instances sufficient for ref proc/proc.py:672-684."""

    def __init__(self, mask, kpts, present):
        if present:
            self.pred_masks = torch.from_numpy(mask[None].astype(bool))
            self.pred_keypoints = torch.from_numpy(kpts[None].astype(np.float32))
        else:
            self.pred_masks = torch.zeros((0,) + mask.shape, dtype=torch.bool)
            self.pred_keypoints = torch.zeros((0,) + kpts.shape, dtype=torch.float32)

    def __len__(self):
        return int(self.pred_masks.shape[0])


def build_case(name):
    geom_name, gen_kw, bg_kw = CASES[name]
    geom = getattr(synthetic.SessionGeometry, geom_name)()
    chunk = synthetic.generate_chunk(geom=geom, **gen_kw)
    roi = synthetic.make_roi(geom)
    bg = synthetic.make_background(geom, dtype=np.dtype(bg_kw['dtype']), half_steps=bg_kw.get('half_steps', False))
    return geom, chunk, roi, bg


def run_reference(name):
    ref = ref_import.load()
    geom, chunk, roi, bg = build_case(name)
    cfg = synthetic.default_config(geom)
    out = {}
    prepped = ref.proc.prep_raw_frames(chunk.frames.copy(), bground_im=bg, roi=roi,
                                       vmin=cfg['min_height'], vmax=cfg['max_height'])
    out['prep'] = prepped
    out['prep_nofix'] = ref.proc.prep_raw_frames(chunk.frames.copy(), bground_im=bg, roi=roi, vmin=cfg['min_height'],
                                                 vmax=cfg['max_height'], fix_invalid_pixels=False)
    out['scale'] = ref.proc.scale_raw_frames(prepped[:4, :, :, None], vmin=cfg['min_height'], vmax=cfg['max_height'])
    outputs = [{'instances': FakeInstances(chunk.masks[i], chunk.keypoints[i], chunk.num_instances[i] > 0)}
               for i in range(prepped.shape[0])]
    feats = ref.proc.instances_to_features(outputs, prepped, None, None, debug=False)
    out['cleaned_frames'] = feats['cleaned_frames']
    out['masks'] = feats['masks']
    out['centroid'] = feats['features']['centroid']
    out['orientation'] = feats['features']['orientation']
    out['axis_length'] = feats['features']['axis_length']
    out['flips'] = feats['flips']
    out['keypoints'] = feats['keypoints']
    out['num_instances'] = feats['num_instances']
    # raw (pre-flip) moment features straight from get_frame_features
    raw, _ = ref.proc.get_frame_features(feats['cleaned_frames'], mask=feats['masks'], use_cc=True, frame_threshold=3,
                                         progress_bar=False)
    out['raw_orientation'] = raw['orientation']
    scal = ref.scalars.compute_scalars(prepped * feats['masks'], feats['features'], min_height=cfg['min_height'],
                                       max_height=cfg['max_height'], true_depth=cfg['true_depth'])
    for k, v in scal.items():
        out['scalars/' + k] = np.asarray(v)
    kd = ref.keypoints.keypoints_to_dict(feats['keypoints'], feats['cleaned_frames'], feats['features']['centroid'],
                                         feats['features']['orientation'], true_depth=cfg['true_depth'])
    for k, v in kd.items():
        out['keypoints/' + k] = np.asarray(v)
    n = prepped.shape[0]
    crop = cfg['crop_size']
    dc = np.zeros((n, crop[0], crop[1]), dtype='uint8')
    mc = np.zeros((n, crop[0], crop[1]), dtype='uint8')
    for i in range(n):
        dc[i] = ref.proc.crop_and_rotate_frame(prepped[i], out['centroid'][i], out['orientation'][i], crop)
        mc[i] = ref.proc.crop_and_rotate_frame(feats['masks'][i], out['centroid'][i], out['orientation'][i], crop)
    out['depth_frames'] = dc
    out['mask_frames'] = mc
    return out


TRACKING_CASE = dict(n_frames=180, seed=11, t0=20, missing_every=23, chunks=(100, 80))


def run_reference_tracking():
    """The Kalman tracking branch (ref proc/proc.py:730-826) over two consecutive chunks with fresh trackers built like
    pipeline/process_features_step.py:41-50.  pykalman is absent here: the reference runs on oracle/pykalman_standin.py."""
    ref = ref_import.load()
    K = ref.kalman
    geom = synthetic.SessionGeometry.kinect_v2()
    kw = {k: v for k, v in TRACKING_CASE.items() if k != 'chunks'}
    chunk = synthetic.generate_chunk(geom=geom, **kw)
    roi, bg = synthetic.make_roi(geom), synthetic.make_background(geom)
    cfg = synthetic.default_config(geom)
    prepped = ref.proc.prep_raw_frames(chunk.frames.copy(), bground_im=bg, roi=roi, vmin=cfg['min_height'], vmax=cfg['max_height'])
    pt = K.KalmanTracker([K.KalmanTrackerPoint2D(order=3, delta_t=1.0), K.KalmanTrackerNPoints2D(8, order=3, delta_t=1.0)])
    at = K.KalmanTracker([K.KalmanTrackerAngle(order=3, delta_t=1.0, degrees=True)])
    outputs = [{'instances': FakeInstances(chunk.masks[i], chunk.keypoints[i], chunk.num_instances[i] > 0)}
               for i in range(prepped.shape[0])]
    out, start = {}, 0
    for c, n in enumerate(TRACKING_CASE['chunks']):
        sl = slice(start, start + n)
        feats = ref.proc.instances_to_features(outputs[sl], prepped[sl], pt, at, debug=False)
        out[f'c{c}/centroid'] = feats['features']['centroid']
        out[f'c{c}/orientation'] = feats['features']['orientation']
        out[f'c{c}/flips'] = feats['flips']
        out[f'c{c}/keypoints'] = feats['keypoints']
        out[f'c{c}/point_last_mean'] = np.array(pt.last_mean)
        out[f'c{c}/angle_last_mean'] = np.array(at.last_mean)
        scal = ref.scalars.compute_scalars(prepped[sl] * feats['masks'], feats['features'], min_height=cfg['min_height'],
                                           max_height=cfg['max_height'], true_depth=cfg['true_depth'])
        for k, v in scal.items():
            out[f'c{c}/scalars/' + k] = np.asarray(v)
        kd = ref.keypoints.keypoints_to_dict(feats['keypoints'], feats['cleaned_frames'], feats['features']['centroid'],
                                             feats['features']['orientation'], true_depth=cfg['true_depth'])
        for k, v in kd.items():
            out[f'c{c}/keypoints/' + k] = np.asarray(v)
        crop = cfg['crop_size']
        dc = np.zeros((n, crop[0], crop[1]), dtype='uint8')
        for i in range(n):
            dc[i] = ref.proc.crop_and_rotate_frame(prepped[sl][i], feats['features']['centroid'][i],
                                                   feats['features']['orientation'][i], crop)
        out[f'c{c}/depth_frames'] = dc
        start += n
    for name in ('transition_covariance', 'observation_covariance', 'initial_state_covariance'):
        out['point_' + name] = np.array(getattr(pt.kalman_filter, name))
    out['angle_transition_covariance'] = np.array(at.kalman_filter.transition_covariance)
    out['angle_observation_covariance'] = np.array(at.kalman_filter.observation_covariance)
    return out


ROI_CASES = {
    # name: (synthetic_bground kwargs, np.random seed, get_roi kwargs as a function of cv2)
    'default': (dict(seed=0), 3, lambda cv2: dict()),
    'ellipse_erode_nofill': (dict(seed=1, h=300, w=250), 5,
                             lambda cv2: dict(strel_dilate=cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (9, 7)),
                                              strel_erode=cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5)), fill_holes=False,
                                              weights=(1, .5, .2), noise_tolerance=20, iters=300)),
    'gradient_filter': (dict(seed=2, h=200, w=260), 6, lambda cv2: dict(gradient_filter=True, gradient_kernel=5, gradient_threshold=400,
                                                                         strel_dilate=cv2.getStructuringElement(cv2.MORPH_RECT, (7, 7)))),
}
ROI_KEEP = 4            # ranked masks stored per case (all bounding boxes are)


def run_reference_roi():
    """Session ROI detection (ref proc/roi.py:14-212) on oracle/roi_oracle.synthetic_bground images.  scikit-image is
    absent here: the reference runs on oracle/skimage_standin.py."""
    import cv2
    import roi_oracle
    ref = ref_import.load()
    out = {}
    for name, (bg_kw, seed, kw) in ROI_CASES.items():
        bg = roi_oracle.synthetic_bground(**bg_kw)
        np.random.seed(seed)
        rois, plane, bboxes, label_im, ranks, shape_index = ref.roi.get_roi(bg, progress_bar=False, **kw(cv2))
        out[name + '/plane'] = plane
        out[name + '/label_im'] = label_im.astype(np.int32)
        out[name + '/ranks'] = ranks
        out[name + '/shape_index'] = shape_index
        out[name + '/bboxes'] = np.stack(bboxes)
        out[name + '/rois'] = np.packbits(np.stack([np.asarray(r) > 0 for r in rois[:ROI_KEEP]]), axis=-1)
    return out


def main():
    import cv2
    os.makedirs(os.path.join(ROOT, 'tests', 'golden'), exist_ok=True)
    for name in CASES:
        out = run_reference(name)
        out['_versions'] = np.array([f'cv2={cv2.__version__}', f'numpy={np.__version__}'])
        path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
        np.savez_compressed(path, **out)
        print(name, '->', path, f'{os.path.getsize(path) / 1e6:.2f} MB')
    out = run_reference_tracking()
    out['_versions'] = np.array([f'cv2={cv2.__version__}', f'numpy={np.__version__}', 'pykalman=stand-in (oracle/pykalman_standin.py)'])
    path = os.path.join(ROOT, 'tests', 'golden', 'kinect_tracking.npz')
    np.savez_compressed(path, **out)
    print('kinect_tracking ->', path, f'{os.path.getsize(path) / 1e6:.2f} MB')

    import scipy
    out = run_reference_roi()
    out['_versions'] = np.array([f'cv2={cv2.__version__}', f'numpy={np.__version__}', f'scipy={scipy.__version__}',
                                 'skimage=stand-in (oracle/skimage_standin.py)'])
    path = os.path.join(ROOT, 'tests', 'golden', 'session_roi.npz')
    np.savez_compressed(path, **out)
    print('session_roi ->', path, f'{os.path.getsize(path) / 1e6:.2f} MB')


if __name__ == '__main__':
    main()
