"""Raw depth decode-to-tensor (mirrors reference io/video.py:33-127).

The wire format is headerless little-endian int16, `width*height*2` bytes per frame, optionally a member of a
tar archive.  "Decoding" is therefore a byte reinterpretation: frames are read with `readinto` straight into ONE
destination buffer -- optionally page-locked, so that the prep kernel can consume the ROI box of the frames
directly from host memory (zero-copy, see bench.py e2e) or a single cudaMemcpyAsync can move them.
"""
from __future__ import annotations

import os
import tarfile
from typing import Iterable, List, Optional, Tuple, Union

import numpy as np


def get_raw_info(filename: Union[str, tarfile.TarInfo], bit_depth: int = 16, frame_dims: Tuple[int, int] = (512, 424)) -> dict:
    """Size bookkeeping of a raw depth file (ref: io/video.py:33-58)."""
    bytes_per_frame = int(frame_dims[0] * frame_dims[1] * bit_depth / 8)
    file_bytes = filename.size if isinstance(filename, tarfile.TarInfo) else os.stat(filename).st_size
    return {'bytes': file_bytes, 'nframes': int(file_bytes / bytes_per_frame), 'dims': frame_dims,
            'bytes_per_frame': bytes_per_frame}


def _consecutive_runs(values: List[int]) -> List[Tuple[int, int]]:
    """[(start, count)] of runs of consecutive integers in a sorted list (ref: io/video.py:130-150)."""
    runs: List[Tuple[int, int]] = []
    for v in values:
        if runs and v == runs[-1][0] + runs[-1][1]:
            runs[-1] = (runs[-1][0], runs[-1][1] + 1)
        else:
            runs.append((v, 1))
    return runs


def read_frames_raw(filename: Union[str, tarfile.TarInfo], frames: Optional[Union[int, Iterable[int]]] = None,
                    frame_dims: Tuple[int, int] = (512, 424), bit_depth: int = 16, dtype="<i2",
                    tar_object: Optional[tarfile.TarFile] = None, pinned: bool = False, as_tensor: bool = False):
    """Read frames of a raw binary depth file into a `(nframes, height, width)` array (ref: io/video.py:67-127).

    Same arguments as the reference plus `pinned`: allocate the result in page-locked host memory (a numpy view of a
    pinned torch tensor); with `as_tensor` the pinned torch tensor itself is returned, which `prep_raw_frames` consumes
    zero-copy.  Requested indices may be unordered or repeated: like the reference, runs of consecutive frames are read
    with one seek + one read each."""
    info = get_raw_info(filename, frame_dims=frame_dims, bit_depth=bit_depth)
    if isinstance(frames, (int, np.integer)):
        frames = [int(frames)]
    elif frames is not None:
        frames = [int(i) for i in frames]
    if frames is None or len(frames) == 0:
        frames = list(range(info['nframes']))
    dt = np.dtype(dtype)
    shape = (len(frames), frame_dims[1], frame_dims[0])
    if pinned:
        import torch
        holder = torch.empty(shape, dtype=torch.int16, pin_memory=True)      # page-locked from the start: no pageable copy
        out = holder.numpy().view(dt)
    else:
        out = np.empty(shape, dtype=dt)
    first_slot = {}
    for slot, idx in enumerate(frames):
        first_slot.setdefault(idx, slot)

    if isinstance(tar_object, tarfile.TarFile):
        handle = tar_object.extractfile(filename)
        if handle is None:
            name = filename.name if isinstance(filename, tarfile.TarInfo) else filename
            raise FileNotFoundError(f'Could not open tar member: {name}')
    elif isinstance(filename, str):
        handle = open(filename, 'rb')
    else:
        raise ValueError('Could not read!')
    contiguous = frames == list(range(frames[0], frames[0] + len(frames)))
    with handle:
        if contiguous:                                  # the chunk iterator's case: one read into the destination
            handle.seek(max(0, frames[0] * info['bytes_per_frame']))
            got = handle.readinto(memoryview(out).cast('B'))
            if got != out.nbytes:
                raise EOFError(f'{filename}: wanted {out.nbytes} bytes, got {got}')
        else:
            for start, count in _consecutive_runs(sorted(set(frames))):
                handle.seek(max(0, start * info['bytes_per_frame']))
                block = np.empty((count, frame_dims[1], frame_dims[0]), dtype=dt)
                got = handle.readinto(memoryview(block).cast('B'))
                if got != block.nbytes:
                    raise EOFError(f'{filename}: wanted {block.nbytes} bytes, got {got}')
                for k in range(count):
                    out[first_slot[start + k]] = block[k]
            for slot, idx in enumerate(frames):          # repeated indices
                if first_slot[idx] != slot:
                    out[slot] = out[first_slot[idx]]
    return holder if (pinned and as_tensor) else out


class RawDepthSession:
    """The slice of `io.session.Session` that the extract pipeline uses (ref: io/session.py:24-110, :112-178, :352-466),
    backed by a raw `depth.dat` file or a `.tar.gz` / `.tgz` session archive holding one: `.bground_im`, `.roi`,
    `.true_depth`, `.nframes`, `.iterate(chunk_size, chunk_overlap)`, `.load_metadata()`, `.load_timestamps()`.
    Background, ROI and true depth are either supplied by the caller or estimated from the file by `find_roi()`
    (ref io/session.py:181-264).  `frame_trim=(head, tail)` drops frames like the reference's `__trim_frames`;
    `frame_dims=None` takes `DepthResolution` from the session's `metadata.json`."""

    def __init__(self, depth_file: str, bground_im: Optional[np.ndarray] = None, roi: Optional[np.ndarray] = None,
                 true_depth: float = float('nan'), frame_dims: Optional[Tuple[int, int]] = (512, 424), pinned: bool = True,
                 frame_trim: Tuple[int, int] = (0, 0)):
        self.pinned = pinned
        self.bground_im, self.roi, self.true_depth = bground_im, roi, float(true_depth)
        self.dirname = os.path.dirname(os.path.abspath(depth_file))
        if depth_file.endswith('.tar.gz') or depth_file.endswith('.tgz'):         # ref io/session.py:53-70
            self.tar: Optional[tarfile.TarFile] = tarfile.open(depth_file, mode='r:*')
            self.tar_members = self.tar.getmembers()
            self.tar_names = [m.name for m in self.tar_members]
            if 'depth.dat' not in self.tar_names:
                raise FileNotFoundError(f'{depth_file}: the archive holds no depth.dat')
            self.depth_file: Union[str, tarfile.TarInfo] = self.tar_members[self.tar_names.index('depth.dat')]
            self.session_id = os.path.basename(depth_file).split('.')[0]
        else:
            self.tar, self.tar_members, self.tar_names = None, None, []
            self.depth_file = depth_file
            self.session_id = os.path.basename(self.dirname)
        if frame_dims is None:
            frame_dims = tuple(int(v) for v in self.load_metadata()['DepthResolution'])
        self.frame_dims = frame_dims
        total = get_raw_info(self.depth_file, frame_dims=frame_dims)['nframes']
        # ref io/session.py:85-99: a trim that would leave nothing is ignored
        self.frame_trim = tuple(frame_trim)
        self.first_frame_idx = frame_trim[0] if 0 < frame_trim[0] < total else 0
        self.last_frame_idx = total - frame_trim[1] if total - frame_trim[1] > self.first_frame_idx else total
        self.nframes = self.last_frame_idx - self.first_frame_idx

    @property
    def is_compressed(self) -> bool:
        return self.tar is not None

    def _open_member(self, name: str):
        """File object of a sibling of the depth file (inside the archive, or next to depth.dat), or None."""
        if self.tar is not None:
            return self.tar.extractfile(self.tar_members[self.tar_names.index(name)]) if name in self.tar_names else None
        path = os.path.join(self.dirname, name)
        return open(path, 'rb') if os.path.exists(path) else None

    def load_metadata(self) -> dict:
        """The session's metadata.json (ref: io/session.py:112-128, io/util.py:66-81)."""
        import json
        handle = self._open_member('metadata.json')
        if handle is None:
            raise ValueError('Could not find metadata.json for this session')
        with handle:
            return json.load(handle)

    def load_timestamps(self) -> np.ndarray:
        """Depth timestamps (ref: io/session.py:131-178): first column of `depth_ts.txt`, else of `timestamps.csv` scaled by
        1000, trimmed to `[first_frame_idx, last_frame_idx)`."""
        for name, factor in (('depth_ts.txt', 1.0), ('timestamps.csv', 1000.0)):
            handle = self._open_member(name)
            if handle is None:
                continue
            with handle:
                stamps = np.array([float(line.decode().split()[0]) for line in handle if line.strip()])
            return stamps[self.first_frame_idx:self.last_frame_idx] * factor
        raise ValueError('Could not locate timestamp file!')

    def read_frames(self, frame_idxs, pinned: Optional[bool] = None, as_tensor: bool = False):
        """Frames by session index (0 = first frame after the head trim)."""
        idxs = [int(i) + self.first_frame_idx for i in frame_idxs]
        pin = self.pinned if pinned is None else pinned
        return read_frames_raw(self.depth_file, idxs, frame_dims=self.frame_dims, tar_object=self.tar, pinned=pin, as_tensor=as_tensor)

    def iterate(self, chunk_size: int = 1000, chunk_overlap: int = 0):
        return _RawIterator(self, chunk_size, chunk_overlap)

    def compute_bground(self, frame_stride: int = 500, med_scale: int = 5) -> np.ndarray:
        """Background image from every `frame_stride`-th frame (ref: io/session.py:217-218 -> proc/roi.py:293-307), on
        the GPU (`proc.get_bground_im`); stores it as `.bground_im` and returns it (float64, (height, width))."""
        from ..proc.roi import get_bground_im
        idxs = list(range(0, self.nframes, max(1, int(frame_stride))))
        frames = self.read_frames(idxs, pinned=False)
        self.bground_im = get_bground_im(frames, med_scale=med_scale)
        return self.bground_im


    def find_roi(self, bg_roi_dilate: Tuple[int, int] = (10, 10), bg_roi_shape: str = 'ellipse', bg_roi_index: int = 0,
                 bg_roi_weights=(1, .1, 1), bg_roi_depth_range: Tuple[int, int] = (650, 750), bg_roi_gradient_filter: bool = False,
                 bg_roi_gradient_threshold: int = 3000, bg_roi_gradient_kernel: int = 7, bg_roi_fill_holes: bool = True,
                 use_plane_bground: bool = False, verbose: bool = False, frame_stride: int = 500, cache_dir: Optional[str] = None):
        """Per-session set-up (ref: io/session.py:181-264): first frame, background image (unless one was given to the
        constructor), arena ROI = the `bg_roi_index`-th ranked region of `proc.get_roi`, and `true_depth` = median background
        depth inside the ROI.  Stores `.bground_im`, `.roi`, `.true_depth` and returns `(first_frame, bground_im, roi,
        true_depth)`.  Background and ROI are computed on the GPU.  With `cache_dir` the reference's tiff cache is kept
        (`first_frame.tiff`, `bground.tiff`, `roi_XX.tiff`, ref :194-257): existing files are loaded instead of recomputed --
        the background then comes back as uint16, as in the reference (SURVEY trap 8)."""
        from ..proc.roi import get_roi
        from ..proc.util import select_strel
        from .image import read_tiff_image, write_image
        use_cache = cache_dir is not None
        cache_dir = cache_dir or ''
        ff_filename = os.path.join(cache_dir, 'first_frame.tiff')
        if use_cache and os.path.exists(ff_filename):
            first_frame = read_tiff_image(ff_filename, scale=True)[None]
        else:
            first_frame = self.read_frames([0], pinned=False)
            if use_cache:
                write_image(ff_filename, first_frame[0], scale=True, scale_factor=tuple(bg_roi_depth_range))
        bg_filename = os.path.join(cache_dir, 'bground.tiff')
        if use_cache and os.path.exists(bg_filename):
            self.bground_im = read_tiff_image(bg_filename, scale=True)
        else:
            if self.bground_im is None:
                self.compute_bground(frame_stride=frame_stride)
            if use_cache and not use_plane_bground:
                write_image(bg_filename, np.asarray(self.bground_im), scale=True)
        bground_im = np.asarray(self.bground_im)
        roi_filename = os.path.join(cache_dir, f'roi_{int(bg_roi_index):02d}.tiff')
        if use_cache and os.path.exists(roi_filename):
            roi = read_tiff_image(roi_filename, scale=True) > 0
        else:
            rois, plane, _, _, _, _ = get_roi(bground_im, strel_dilate=select_strel(bg_roi_shape, bg_roi_dilate), weights=bg_roi_weights,
                                              depth_range=bg_roi_depth_range, gradient_filter=bg_roi_gradient_filter,
                                              gradient_threshold=bg_roi_gradient_threshold, gradient_kernel=bg_roi_gradient_kernel,
                                              fill_holes=bg_roi_fill_holes, progress_bar=verbose)
            if use_plane_bground:                           # io/session.py:245-252: the fitted plane replaces the median image
                yy, xx = np.mgrid[:bground_im.shape[0], :bground_im.shape[1]]
                bground_im = ((xx * plane[0] + yy * plane[1]) + plane[3]) / -plane[2]
                if use_cache:
                    write_image(bg_filename, bground_im, scale=True)
            roi = rois[bg_roi_index]
            if use_cache:
                write_image(roi_filename, roi, scale=True, dtype='uint8')
        self.bground_im, self.roi = bground_im, roi
        self.true_depth = float(np.median(bground_im[roi > 0]))
        return first_frame, bground_im, roi, self.true_depth


class _RawIterator:
    def __init__(self, session: RawDepthSession, chunk_size: int, chunk_overlap: int):
        from ..shard import chunk_ranges
        self.session = session
        self.batches = chunk_ranges(session.nframes, chunk_size, chunk_overlap)
        self.filters = []
        self._pos = 0

    def attach_filter(self, stream=None, filterer=None):
        self.filters.append(filterer)

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return self

    def __next__(self):
        if self._pos >= len(self.batches):
            raise StopIteration
        idxs = self.batches[self._pos]
        self._pos += 1
        frames = self.session.read_frames(idxs, as_tensor=self.session.pinned)
        for f in self.filters:
            frames = f(frames)
        return list(idxs), frames
