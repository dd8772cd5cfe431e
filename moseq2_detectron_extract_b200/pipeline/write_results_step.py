"""Result writer step (mirrors reference pipeline/write_results_step.py:14-73 and the dataset
layout of io/result.py:14-130): per chunk, the cropped frames / masks, the 17 scalars, the 96 keypoint columns and the
flips go to `results_XX.h5` when h5py is importable, otherwise to `results_XX.npz` with the SAME dataset paths as keys
(`frames`, `frames_mask`, `scalars/<name>`, `keypoints/<name>`, `metadata/extraction/<...>`), and the per-frame keypoint
table goes to `keypoints_XX.tsv` with the reference's columns.  Host-side I/O, not bandwidth-critical: the reference
rewrites the whole TSV after every chunk, here rows are appended.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np

from ..proc.keypoints import keypoint_attributes
from ..proc.scalars import scalar_attributes
from .pipeline_step import ProcessPipelineStep

try:                                    # pragma: no cover - h5py is not part of this image
    import h5py
except Exception:                       # pylint: disable=broad-except
    h5py = None


def _plain(obj):
    """numpy / tuple values -> plain Python for YAML."""
    if isinstance(obj, dict):
        return {str(k): _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, np.generic):
        return obj.item()
    return obj if isinstance(obj, (str, int, float, bool)) or obj is None else str(obj)


def _host(x):
    return x.detach().cpu().numpy() if hasattr(x, 'detach') else np.asarray(x)


class ResultStore:
    """Datasets of one extraction result (ref: io/result.py:14-102), filled chunk by chunk (ref: :105-130)."""

    def __init__(self, path_without_ext: str, config: dict):
        self.config = config
        n, crop = int(config['nframes']), tuple(config['crop_size'])
        self.arrays: Dict[str, np.ndarray] = {}
        self.attrs: Dict[str, str] = {}
        for name, text in scalar_attributes().items():
            self._create(f'scalars/{name}', (n,), 'float32', text)
        for name, text in keypoint_attributes().items():
            self._create(f'keypoints/{name}', (n,), 'float32', text)
        self._create('frames', (n, crop[0], crop[1]), config.get('frame_dtype', 'uint8'),
                     '3D Numpy array of depth frames (nframes x w x h, in mm)')
        self._create('frames_mask', (n, crop[0], crop[1]), 'bool', 'Boolean mask, false=not mouse, true=mouse')
        self._create('metadata/extraction/flips', (n,), 'bool', 'Output from flip classifier, false=no flip, true=flip')
        for key, text in (('timestamps', 'Depth video timestamps'), ('true_depth', 'Detected true depth of arena floor in mm'),
                          ('roi', 'ROI mask'), ('first_frame', 'First frame of depth dataset'),
                          ('bground_im', 'Computed background image')):
            if config.get(key) is not None:
                path = 'timestamps' if key == 'timestamps' else 'metadata/extraction/' + ('background' if key == 'bground_im' else key)
                self.arrays[path] = _host(config[key])
                self.attrs[path] = text
        self.arrays['metadata/extraction/extract_version'] = np.array('moseq2-detectron-extract_b200')
        # ref: io/result.py:31,88-102 -- uuid, the extraction parameters and the acquisition metadata of the session
        status = config.get('status_dict') or {}
        if status.get('uuid') is not None:
            self.arrays['metadata/uuid'] = np.array(str(status['uuid']))
        for key, value in (status.get('parameters') or {}).items():
            if value is not None and not callable(value):
                try:
                    self.arrays[f'metadata/extraction/parameters/{key}'] = np.asarray(value)
                except Exception:          # pylint: disable=broad-except
                    self.arrays[f'metadata/extraction/parameters/{key}'] = np.array(str(value))
        for key, value in (status.get('metadata') or {}).items():
            self.arrays[f'metadata/acquisition/{key}'] = np.asarray([] if value is None else value)
        self.path = path_without_ext + ('.h5' if h5py is not None else '.npz')
        self.closed = False
        self.frames_written = 0

    def _create(self, path, shape, dtype, description):
        self.arrays[path] = np.zeros(shape, dtype=dtype)
        self.attrs[path] = description

    def write_chunk(self, results: dict) -> None:
        """ref: io/result.py:105-130 -- the first `offset` frames of a chunk overlap the previous chunk and are dropped."""
        off = int(results.get('offset', 0))
        rng = list(results['frame_idxs'])[off:]
        for name, col in results['scalars'].items():
            self.arrays[f'scalars/{name}'][rng] = _host(col)[off:]
        self.arrays['frames'][rng] = _host(results['depth_frames'])[off:]
        self.arrays['frames_mask'][rng] = _host(results['mask_frames'])[off:] != 0
        self.arrays['metadata/extraction/flips'][rng] = _host(results['features']['flips'])[off:]
        for name, col in results['keypoints'].items():
            self.arrays[f'keypoints/{name}'][rng] = _host(col)[off:]
        self.frames_written += len(rng)

    def close(self, complete: bool = True) -> str:
        """Write the file.  A run that did not reach its end (an upstream step failed, the pipeline was shut down) is NOT written
        under the result's name: its partial contents go to `<name>.partial<ext>` so that a zero-padded file can never be
        mistaken for a finished extraction."""
        if self.closed:
            return self.path
        if not complete:
            root, ext = os.path.splitext(self.path)
            self.path = root + '.partial' + ext
        if h5py is not None:            # pragma: no cover
            with h5py.File(self.path, 'w') as f:
                for path, arr in self.arrays.items():
                    ds = f.create_dataset(path, data=arr, compression='gzip' if arr.ndim > 0 and arr.dtype.kind != 'U' else None)
                    if path in self.attrs:
                        ds.attrs['description'] = self.attrs[path]
        else:
            np.savez_compressed(self.path, **self.arrays)
        self.closed = True
        return self.path


class ResultWriterStep(ProcessPipelineStep):
    """Writes results_XX.{h5|npz} and keypoints_XX.tsv (ref: pipeline/write_results_step.py:14-73)."""

    def initialize(self):
        idx = int(self.config.get('bg_roi_index', 0))
        out_dir = self.config['output_dir']
        os.makedirs(out_dir, exist_ok=True)
        self.store = ResultStore(os.path.join(out_dir, f'results_{idx:02d}'), self.config)
        self.keypoint_data_dest = os.path.join(out_dir, f'keypoints_{idx:02d}.tsv')
        self.status_filename = os.path.join(out_dir, f'results_{idx:02d}.yaml')
        self._tsv_header: Optional[list] = None

    def process(self, data):
        self._write_tsv(data)
        self.store.write_chunk(data)
        self.update_progress(int(data['chunk'].shape[0]))
        return data

    def finalize(self):
        """After an error in this step."""
        if getattr(self, 'store', None) is not None:
            self.store.close(complete=False)
            self._write_status(False)

    def shutdown(self):
        """End of the step's loop: a complete result only when the end-of-stream sentinel arrived and nothing failed."""
        if getattr(self, 'store', None) is not None and not self.store.closed:
            complete = self.is_complete.is_set() and self.error is None
            self.store.close(complete=complete)
            self._write_status(complete)

    def _write_status(self, complete: bool) -> None:
        """results_XX.yaml (ref: extract.py:55-62,131-132): uuid, parameters, acquisition metadata and the `complete` flag."""
        status = dict(self.config.get('status_dict') or {})
        status['complete'] = bool(complete)
        status['frames_written'] = int(self.store.frames_written)
        try:
            import yaml
            with open(self.status_filename, 'w', encoding='utf-8') as fh:
                yaml.safe_dump(_plain(status), fh)
        except Exception:              # pylint: disable=broad-except
            pass

    # ---- keypoints_XX.tsv: Frame_Idx, Flip, Centroid_X, Centroid_Y, Angle, then the 96 keypoint columns -------------------
    def _write_tsv(self, data) -> None:
        feats = data['features']
        cen = _host(feats['features']['centroid'])
        ang = _host(feats['features']['orientation'])
        flips = _host(feats['flips'])
        kp = {k: _host(v) for k, v in data['keypoints'].items()}
        header = ['Frame_Idx', 'Flip', 'Centroid_X', 'Centroid_Y', 'Angle'] + list(kp)
        first = self._tsv_header is None
        self._tsv_header = header
        with open(self.keypoint_data_dest, 'w' if first else 'a', encoding='utf-8') as fh:
            if first:
                fh.write('\t'.join(header) + '\n')
            for i, frame_idx in enumerate(data['frame_idxs']):
                row = [str(int(frame_idx)), str(bool(flips[i])), repr(float(cen[i][0])), repr(float(cen[i][1])), repr(float(ang[i]))]
                row += [repr(float(kp[k][i])) for k in header[5:]]
                fh.write('\t'.join(row) + '\n')
