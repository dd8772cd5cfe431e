"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy + OpenCV, the reference's own arithmetic libraries) of the per-frame
extract hot path of tischfieldlab/moseq2-detectron-extract.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
file; the product package never does (it fails loudly if its CUDA library is missing).

Parity pin: the reference ships NO golden vectors or numeric tests for this path
(reference tests/test_entry_points.py:26-40 only smoke-tests `--help`).  The pin is therefore
(1) outputs of the UNMODIFIED reference functions, imported from /root/reference by
`oracle/ref_import.py` and stored by `oracle/make_golden.py` under tests/golden/*.npz, against
which every function below is checked in tests/test_oracle_vs_golden.py, and (2) OpenCV 4.13.0
(as installed; the reference does not pin a version, reference setup.py:13-16) for the
functions that are restated from OpenCV's published algorithms (3x3 median, flat morphological
opening, polygon moments of the outer contour, fixed-point bilinear warpAffine): each of those
has a `*_cv2` twin that calls OpenCV exactly like the reference does and a `*_np` restatement,
and the tests assert the two agree bit-for-bit on random inputs.  Third-party pieces that are not
installed anywhere here XX
from their published behaviour: "parity unpinned" for those two.

Every function cites the reference lines it follows as `ref: <file>:<lines>`.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

try:  # OpenCV is the reference's arithmetic library; it is present in this image
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

KEYPOINT_NAMES = ['Nose', 'Left Ear', 'Right Ear', 'Neck', 'Left Hip', 'Right Hip', 'TailBase', 'TailTip']

# cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (9, 9)) in OpenCV 4.13.0 (ref: proc/proc.py:481)
ELLIPSE9_HALF_WIDTHS = (0, 3, 3, 4, 4, 4, 3, 3, 0)   # half-width of the set run in rows dy = -4..4


# ------------------------------------------------------------------------------------------
# a2  prep_raw_frames                                             ref: proc/proc.py:129-172
# ------------------------------------------------------------------------------------------
def bbox_of_roi(roi: np.ndarray) -> Optional[Tuple[int, int, int, int]]:
    """(y0, x0, y1, x1), max-exclusive exactly like the slicing in ref: proc/roi.py:235 on the
    min/max of ref: proc/roi.py:248-254 (so the last ROI row and column are dropped)."""
    rows = np.flatnonzero(roi.any(axis=1))
    cols = np.flatnonzero(roi.any(axis=0))
    if rows.size == 0:
        return None
    return int(rows[0]), int(cols[0]), int(rows[-1]), int(cols[-1])


def invalid_pixel_mask(frames: np.ndarray) -> np.ndarray:
    """ref: proc/proc.py:175-186 -- Kinect v2 marks bad pixels with raw value 0."""
    return (frames == 0).astype(np.uint8)


def prep_frames(frames: np.ndarray, bground: Optional[np.ndarray], roi: Optional[np.ndarray],
                vmin: Optional[float], vmax: Optional[float], fix_invalid: bool = True,
                return_invalid: bool = False):
    """ref: proc/proc.py:129-172 (+ apply_roi proc/roi.py:215-236, fill proc/proc.py:189-210).

    Arithmetic order matters for bit-exactness: subtract in numpy's promoted dtype
    (float32 background - int16 frame -> float32; float64 -> float64; uint16 -> int32), multiply by
    the ROI, crop, clamp (below vmin -> 0, above vmax -> vmax), then C-style truncating cast to u8.
    """
    bad = invalid_pixel_mask(frames) if fix_invalid else None
    work = frames
    if bground is not None:
        work = bground - work
    if roi is not None:
        box = bbox_of_roi(roi)
        work = work * roi
        if bad is not None:
            bad = bad * roi
        if box is not None:
            y0, x0, y1, x1 = box
            work = work[:, y0:y1, x0:x1]
            if bad is not None:
                bad = bad[:, y0:y1, x0:x1]
    work = np.array(work, copy=True)
    if vmin is not None:
        work[work < vmin] = 0
    if vmax is not None:
        work[work > vmax] = vmax
    out = work.astype(np.uint8)
    if fix_invalid:
        out = inpaint_invalid(out, bad)
    if return_invalid:
        return out, bad
    return out


def inpaint_invalid(frames_u8: np.ndarray, bad: np.ndarray) -> np.ndarray:
    """ref: proc/proc.py:189-210 -- per-frame cv2.inpaint(radius 3, Navier-Stokes)."""
    out = frames_u8.copy()
    for i in range(out.shape[0]):
        if bad[i].any():
            out[i] = cv2.inpaint(out[i], np.ascontiguousarray(bad[i]), 3, cv2.INPAINT_NS)
    return out


# ------------------------------------------------------------------------------------------
# a3  scale_raw_frames                                            ref: proc/proc.py:214-234
# ------------------------------------------------------------------------------------------
def scale_frames(frames: np.ndarray, vmin: float, vmax: float) -> np.ndarray:
    """ref: proc/proc.py:214-234 with dtype='uint8' (dmin 0, dmax 255), float64 arithmetic."""
    gain = (255.0 - 0.0) / (vmax - vmin)
    return ((frames - vmin) * gain + 0.0).astype(np.uint8)


def scale_lut(vmin: float, vmax: float) -> np.ndarray:
    """The same map as a 256-entry table (what the kernel evaluates per input byte)."""
    return scale_frames(np.arange(256, dtype=np.uint8), vmin, vmax)


# ------------------------------------------------------------------------------------------
# a6  clean_frames(iters_tail=3)                                  ref: proc/proc.py:480-515
# ------------------------------------------------------------------------------------------
def clean_frames_cv2(frames: np.ndarray) -> np.ndarray:
    """ref: proc/proc.py:496-509 with the defaults instances_to_features uses (proc/proc.py:715):
    medianBlur(3) then morphologyEx(OPEN, 9x9 ellipse).  `iters_tail` lands in the positional
    `dst` slot of cv2.morphologyEx (proc/proc.py:509), so exactly ONE opening is applied."""
    strel = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (9, 9))
    out = np.empty_like(frames, dtype=np.uint8)
    for i in range(frames.shape[0]):
        med = cv2.medianBlur(np.ascontiguousarray(frames[i], dtype=np.uint8), 3)
        out[i] = cv2.morphologyEx(med, cv2.MORPH_OPEN, strel)
    return out


def median3_np(img: np.ndarray) -> np.ndarray:
    """OpenCV medianBlur(ksize=3) on 8-bit: 3x3 median with BORDER_REPLICATE."""
    p = np.pad(img, 1, mode='edge')
    h, w = img.shape
    stack = np.stack([p[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)], axis=0)
    return np.sort(stack, axis=0)[4]


def _morph_ellipse9(img: np.ndarray, take_min: bool) -> np.ndarray:
    """Flat erosion/dilation by the 9x9 ellipse; pixels outside the image never win
    (OpenCV's default morphology border value is +max for erode and -max for dilate)."""
    fill = 255 if take_min else 0
    p = np.pad(img, 4, mode='constant', constant_values=fill)
    h, w = img.shape
    acc = np.full_like(img, fill)
    op = np.minimum if take_min else np.maximum
    for dy in range(-4, 5):
        hw = ELLIPSE9_HALF_WIDTHS[dy + 4]
        for dx in range(-hw, hw + 1):
            acc = op(acc, p[4 + dy:4 + dy + h, 4 + dx:4 + dx + w])
    return acc


def clean_frames_np(frames: np.ndarray) -> np.ndarray:
    """Pure-numpy restatement of `clean_frames_cv2` (the arithmetic the CUDA kernel follows)."""
    out = np.empty_like(frames, dtype=np.uint8)
    for i in range(frames.shape[0]):
        med = median3_np(frames[i].astype(np.uint8))
        out[i] = _morph_ellipse9(_morph_ellipse9(med, True), False)
    return out


# ------------------------------------------------------------------------------------------
# a7  get_frame_features + im_moment_features           ref: proc/proc.py:237-302, 518-549
# ------------------------------------------------------------------------------------------
def moment_features_from_sums(a00: float, a10: float, a01: float, a20: float, a11: float, a02: float):
    """From OpenCV's contour accumulators (a00 = 2*area, a10 = 6*Int x, a01 = 6*Int y,
    a20 = 12*Int x^2, a11 = 24*Int xy, a02 = 12*Int y^2; all >= 0 here) to the reference's
    features, in the same float64 operation order as cv::contourMoments/completeMomentState
    followed by ref: proc/proc.py:529-547."""
    nan = float('nan')
    if not (abs(a00) > np.finfo(np.float32).eps):
        return (nan, nan), nan, (nan, nan)
    m00 = a00 * 0.5
    m10 = a10 * 0.16666666666666666666666666666667
    m01 = a01 * 0.16666666666666666666666666666667
    m20 = a20 * 0.083333333333333333333333333333333
    m11 = a11 * 0.041666666666666666666666666666667
    m02 = a02 * 0.083333333333333333333333333333333
    inv = 1.0 / m00
    cx, cy = m10 * inv, m01 * inv
    mu20 = m20 - m10 * cx
    mu11 = m11 - m10 * cy
    mu02 = m02 - m01 * cy
    num = 2 * mu11
    den = mu20 - mu02
    common = np.sqrt(4 * np.square(mu11) + np.square(den))
    orientation = -.5 * np.arctan2(num, den)
    centroid = (m10 / m00, m01 / m00)
    with np.errstate(invalid='ignore'):
        axes = (2 * np.sqrt(2) * np.sqrt((mu20 + mu02 + common) / m00),
                2 * np.sqrt(2) * np.sqrt((mu20 + mu02 - common) / m00))
    return centroid, float(orientation), axes


def frame_features_cv2(cleaned: np.ndarray, masks: np.ndarray, frame_threshold: float = 3):
    """ref: proc/proc.py:237-302 as called from proc/proc.py:716 (mask given, use_cc=True).
    `get_largest_cc((frame > -30))` on uint8 input is all-True (SURVEY.md trap 2), so
    cc_mask is dropped; tests/test_oracle_vs_golden.py checks the result against the real call."""
    n = cleaned.shape[0]
    centroid = np.full((n, 2), np.nan)
    orientation = np.full((n,), np.nan)
    axis_length = np.full((n, 2), np.nan)
    for i in range(n):
        fm = np.logical_and(cleaned[i] > frame_threshold, masks[i]).astype(np.uint8)
        cnts, _ = cv2.findContours(fm, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        if len(cnts) == 0:
            continue
        areas = np.array([cv2.contourArea(c) for c in cnts])
        m = cv2.moments(cnts[int(areas.argmax())])
        if m['m00'] == 0:
            continue
        num = 2 * m['mu11']
        den = m['mu20'] - m['mu02']
        common = np.sqrt(4 * np.square(m['mu11']) + np.square(den))
        orientation[i] = -.5 * np.arctan2(num, den)
        centroid[i] = (m['m10'] / m['m00'], m['m01'] / m['m00'])
        with np.errstate(invalid='ignore'):
            axis_length[i] = (2 * np.sqrt(2) * np.sqrt((m['mu20'] + m['mu02'] + common) / m['m00']),
                              2 * np.sqrt(2) * np.sqrt((m['mu20'] + m['mu02'] - common) / m['m00']))
    return {'centroid': centroid, 'orientation': orientation, 'axis_length': axis_length}


# 24 x (A, U, V, UU, UV, VV) of the part of a unit cell covered by the outer-contour polygon,
# in cell-local coordinates (u to the right, v down); key = which corner is NOT set.
_CELL24 = {
    'full': (24, 12, 12, 8, 6, 8),
    'no_br': (12, 4, 4, 2, 1, 2),
    'no_tl': (12, 8, 8, 6, 5, 6),
    'no_tr': (12, 4, 8, 2, 3, 6),
    'no_bl': (12, 8, 4, 6, 3, 2),
}


def polygon_sums_of_blob(blob: np.ndarray) -> Tuple[int, ...]:
    """Exact integer 24x area integrals of the polygon that OpenCV's border follower traces
    around a hole-free 8-connected blob (vertices at pixel centres): every 2x2 block of pixel
    centres contributes its unit square when all four are set, the corner triangle when exactly
    three are set, nothing otherwise.  Returns (S00, S10, S01, S20, S11, S02) = 24 * Int(1, x, y, x^2, xy, y^2)."""
    b = blob.astype(bool)
    tl, tr, bl, br = b[:-1, :-1], b[:-1, 1:], b[1:, :-1], b[1:, 1:]
    classes = {
        'full': tl & tr & bl & br,
        'no_br': tl & tr & bl & ~br,
        'no_tl': ~tl & tr & bl & br,
        'no_tr': tl & ~tr & bl & br,
        'no_bl': tl & tr & ~bl & br,
    }
    s = [0, 0, 0, 0, 0, 0]
    for name, sel in classes.items():
        jj, ii = np.nonzero(sel)            # row j (y), column i (x) of the cell's top-left corner
        if ii.size == 0:
            continue
        ii = ii.astype(np.int64)
        jj = jj.astype(np.int64)
        A, U, V, UU, UV, VV = _CELL24[name]
        n = ii.size
        s[0] += A * n
        s[1] += A * ii.sum() + U * n
        s[2] += A * jj.sum() + V * n
        s[3] += A * (ii * ii).sum() + 2 * U * ii.sum() + UU * n
        s[4] += A * (ii * jj).sum() + V * ii.sum() + U * jj.sum() + UV * n
        s[5] += A * (jj * jj).sum() + 2 * V * jj.sum() + VV * n
    return tuple(int(v) for v in s)


def _label(mask: np.ndarray, eight: bool):
    from scipy import ndimage
    st = np.ones((3, 3), dtype=int) if eight else None
    return ndimage.label(mask, structure=st)


def frame_features_np(cleaned: np.ndarray, masks: np.ndarray, frame_threshold: float = 3,
                      return_sums: bool = False):
    """Restatement of `frame_features_cv2` without contour tracing (what the CUDA kernel does):
    1. fm = (cleaned > thr) & mask
    2. flood the background 4-connectedly from a virtual ring outside the frame; everything not
       reached (foreground + enclosed holes + whatever sits inside holes) is the filled set F
    3. 8-connected components of F are the filled top-level blobs = the outer contours that can win
       `argmax(contourArea)` (a hole or nested contour never has a larger area than its encloser)
    4. per blob exact integer polygon sums; largest polygon area wins; on ties OpenCV 4.13 lists
       sibling contours in REVERSE raster order of their first pixel, so the LAST blob in raster
       order wins (np.argmax takes the first of the list)
    5. float64 epilogue identical to OpenCV's + ref: proc/proc.py:529-547."""
    n = cleaned.shape[0]
    centroid = np.full((n, 2), np.nan)
    orientation = np.full((n,), np.nan)
    axis_length = np.full((n, 2), np.nan)
    sums = np.zeros((n, 6), dtype=np.int64)
    for i in range(n):
        fm = np.logical_and(cleaned[i] > frame_threshold, masks[i] != 0)
        if not fm.any():
            continue
        padded_bg = np.pad(~fm, 1, mode='constant', constant_values=True)
        lab, _ = _label(padded_bg, eight=False)
        outside = (lab == lab[0, 0])[1:-1, 1:-1]
        filled = ~outside
        lab8, ncomp = _label(filled, eight=True)
        best, best_sums = -1, None
        for c in range(1, ncomp + 1):            # scipy labels in raster order of first pixel
            s = polygon_sums_of_blob(lab8 == c)
            if s[0] >= best:                      # '>=' : later blob wins ties
                best, best_sums = s[0], s
        sums[i] = best_sums
        S00, S10, S01, S20, S11, S02 = best_sums
        cen, ori, axes = moment_features_from_sums(S00 / 12.0, S10 / 4.0, S01 / 4.0, S20 / 2.0, float(S11), S02 / 2.0)
        centroid[i], orientation[i], axis_length[i] = cen, ori, axes
    out = {'centroid': centroid, 'orientation': orientation, 'axis_length': axis_length}
    if return_sums:
        out['sums24'] = sums
    return out


# ------------------------------------------------------------------------------------------
# a8-a10  angles, keypoint flips, iterative angle filter   ref: proc/proc.py:688-724, 827-889, 600-654
# ------------------------------------------------------------------------------------------
def clamp_deg(a: np.ndarray) -> np.ndarray:
    """ref: proc/proc.py:688-691"""
    return np.where(a < 0, 360 + a, a) % 360


def rotate_about(points_xy: np.ndarray, centers: np.ndarray, angles_deg: np.ndarray) -> np.ndarray:
    """ref: proc/keypoints.py:11-64 -- rotate (N, K, 2) points about (N, 2) centres by -angle."""
    t = np.deg2rad(-np.asarray(angles_deg, dtype=np.float64))
    c, s = np.cos(t)[:, None], np.sin(t)[:, None]
    dx = points_xy[..., 0] - centers[:, None, 0]
    dy = points_xy[..., 1] - centers[:, None, 1]
    out = np.empty(points_xy.shape, dtype=np.float64)
    out[..., 0] = (c * dx + (-s) * dy) + centers[:, None, 0]
    out[..., 1] = (s * dx + c * dy) + centers[:, None, 1]
    return out


def keypoint_flips(keypoints: np.ndarray, centroids: np.ndarray, angles: np.ndarray, lengths: np.ndarray):
    """ref: proc/proc.py:851-889 -- front (0..3) and rear (4..6) keypoints vote for the nearer end."""
    rot = rotate_about(keypoints[..., :2].astype(np.float64), centroids, angles)
    lo = centroids[:, 0] - lengths / 2
    hi = centroids[:, 0] + lengths / 2
    with np.errstate(invalid='ignore'):
        votes = np.where(np.abs(lo[:, None] - rot[..., 0]) < np.abs(hi[:, None] - rot[..., 0]), -1, 1)
    front = votes[:, 0:4].mean(axis=1)
    rear = votes[:, 4:7].mean(axis=1)
    flips = front < rear
    want_front = np.where(flips, -1, 1)[:, None]
    agree = (votes[:, 0:4] == want_front).sum(axis=1) + (votes[:, 4:7] == -want_front).sum(axis=1)
    return flips, agree / 7


def trailing_median3(a: np.ndarray, window: int = 3) -> np.ndarray:
    """bottleneck.move_median(a, window, min_count=1) (ref: proc/proc.py:618): median of the
    non-NaN values among a[i-window+1 .. i]; NaN when there are none.  bottleneck is not installed
    in this image; semantics restated from its documentation (parity unpinned)."""
    n = a.shape[0]
    out = np.full((n,), np.nan)
    for i in range(n):
        win = a[max(0, i - window + 1):i + 1]
        win = win[~np.isnan(win)]
        if win.size:
            out[i] = np.median(win)
    return out


def angle_filter_pass(angles: np.ndarray, window: int = 3, tolerance: float = 60) -> np.ndarray:
    """ref: proc/proc.py:600-624"""
    out = angles.copy()
    med = trailing_median3(angles, min(window, angles.shape[0]))
    with np.errstate(invalid='ignore'):
        d = out - med
        ad = np.abs(d)
        hit = (ad > (180 - tolerance)) & (ad < (180 + tolerance))
    out[hit] = out[hit] + (-180 * np.sign(d[hit]))
    return out


def iterative_angle_filter(angles: np.ndarray, window: int = 3, tolerance: float = 60, max_iters: int = 1000):
    """ref: proc/proc.py:627-654 -- note the loop runs max_iters+1 passes when it never
    converges (any NaN makes np.allclose false, SURVEY.md trap 16)."""
    last = angles.copy()
    done = 0
    while True:
        if done > max_iters:
            break
        done += 1
        cur = angle_filter_pass(last, window, tolerance)
        if np.allclose(cur, last):
            break
        last = cur
    with np.errstate(invalid='ignore'):
        flips = np.isclose(np.abs(cur - angles), 180)
    return cur, flips, done


# ------------------------------------------------------------------------------------------
# a11  keypoints_to_dict                                        ref: proc/keypoints.py:93-165
# ------------------------------------------------------------------------------------------
def px_to_mm(coords: np.ndarray, true_depth: float) -> np.ndarray:
    """ref: proc/util.py:29-61 -- always Kinect-v2 resolution/FOV (SURVEY.md trap 11)."""
    fw = 512 / (2 * np.deg2rad(70.6 / 2))
    fh = 424 / (2 * np.deg2rad(60 / 2))
    out = np.zeros_like(coords, dtype=np.float64)
    out[:, 0] = true_depth * (coords[:, 0] - 256) / fw
    out[:, 1] = true_depth * (coords[:, 1] - 212) / fh
    return out


def keypoint_table(keypoints: np.ndarray, cleaned: np.ndarray, centers: np.ndarray, angles: np.ndarray,
                   true_depth: float) -> Dict[str, np.ndarray]:
    """ref: proc/keypoints.py:93-165; returns the same 96 arrays under the same names."""
    kp = keypoints.astype(np.float64)
    n, k = kp.shape[:2]
    with np.errstate(invalid='ignore'):
        xi = np.clip(np.floor(kp[:, :, 0]).astype(np.int64), 0, cleaned.shape[2] - 1)
        yi = np.clip(np.floor(kp[:, :, 1]).astype(np.int64), 0, cleaned.shape[1] - 1)
    z = cleaned[np.arange(n)[:, None], yi, xi].astype(np.float64)
    ref_mm = np.zeros_like(kp)
    ref_mm[:, :, 2] = kp[:, :, 2]
    for j in range(k):
        ref_mm[:, j, :2] = px_to_mm(kp[:, j, :2], true_depth)
    rot_px = rotate_about(kp[:, :, :2], centers, angles) - centers[:, None, :]
    cen_mm = px_to_mm(centers, true_depth)
    rot_mm = rotate_about(ref_mm[:, :, :2], cen_mm, angles) - cen_mm[:, None, :]
    out = {}
    for j, name in enumerate(KEYPOINT_NAMES):
        out[f'reference/{name}_x_px'] = kp[:, j, 0]
        out[f'reference/{name}_y_px'] = kp[:, j, 1]
        out[f'reference/{name}_score'] = kp[:, j, 2]
        out[f'reference/{name}_x_mm'] = ref_mm[:, j, 0]
        out[f'reference/{name}_y_mm'] = ref_mm[:, j, 1]
        out[f'reference/{name}_z_mm'] = z[:, j]
        out[f'rotated/{name}_x_px'] = rot_px[:, j, 0]
        out[f'rotated/{name}_y_px'] = rot_px[:, j, 1]
        out[f'rotated/{name}_score'] = kp[:, j, 2]
        out[f'rotated/{name}_x_mm'] = rot_mm[:, j, 0]
        out[f'rotated/{name}_y_mm'] = rot_mm[:, j, 1]
        out[f'rotated/{name}_z_mm'] = z[:, j]
    return out


# ------------------------------------------------------------------------------------------
# a12  compute_scalars                                          ref: proc/scalars.py:36-120
# ------------------------------------------------------------------------------------------
def scalar_table(masked_frames: np.ndarray, centroid: np.ndarray, orientation_deg: np.ndarray,
                 axis_length: np.ndarray, min_height: float, max_height: float, true_depth: float):
    """ref: proc/scalars.py:36-120.  `masked_frames` is chunk*mask (uint8).  Output dtypes follow
    what the reference actually returns (most entries are re-bound to float64 arrays; only
    height_ave_mm stays float32)."""
    n = masked_frames.shape[0]
    cen_mm = px_to_mm(centroid, true_depth)
    step_mm = np.abs(px_to_mm(centroid + 1, true_depth) - cen_mm)
    inside = (masked_frames > min_height) & (masked_frames < max_height)
    count = inside.sum(axis=(1, 2))
    total = (masked_frames.astype(np.int64) * inside).sum(axis=(1, 2))
    height = np.zeros((n,), dtype=np.float32)
    nz = count > 0
    height[nz] = (total[nz] / count[nz]).astype(np.float32)

    def first_diff(v):
        return np.diff(np.concatenate((v[:1], v)))

    out = {
        'centroid_x_px': centroid[:, 0], 'centroid_y_px': centroid[:, 1],
        'centroid_x_mm': cen_mm[:, 0], 'centroid_y_mm': cen_mm[:, 1],
        'width_px': axis_length.min(axis=1), 'length_px': axis_length.max(axis=1),
        'area_px': count, 'height_ave_mm': height, 'angle': np.deg2rad(orientation_deg),
    }
    out['width_mm'] = out['width_px'] * step_mm[:, 1]
    out['length_mm'] = out['length_px'] * step_mm[:, 0]
    out['area_mm'] = out['area_px'] * step_mm.mean(axis=1)
    vx, vy, vz = first_diff(out['centroid_x_px']), first_diff(out['centroid_y_px']), first_diff(height)
    out['velocity_2d_px'] = np.hypot(vx, vy)
    out['velocity_3d_px'] = np.sqrt(np.square(vx) + np.square(vy) + np.square(vz))
    vx, vy = first_diff(out['centroid_x_mm']), first_diff(out['centroid_y_mm'])
    out['velocity_2d_mm'] = np.hypot(vx, vy)
    out['velocity_3d_mm'] = np.sqrt(np.square(vx) + np.square(vy) + np.square(vz))
    out['velocity_theta'] = np.arctan2(vy, vx)
    return out


# ------------------------------------------------------------------------------------------
# a13  crop_and_rotate_frame                                    ref: proc/proc.py:305-335
# ------------------------------------------------------------------------------------------
def crop_rotate_cv2(frame: np.ndarray, center: Sequence[float], angle: float, crop: Tuple[int, int] = (80, 80)):
    """ref: proc/proc.py:305-335, calling OpenCV the way the reference does."""
    cw, ch = int(crop[0]), int(crop[1])
    blank = np.zeros((cw, ch), dtype=frame.dtype)      # the reference's shape=crop_size
    if np.isnan(angle) or np.isnan(center[0]) or np.isnan(center[1]):
        return blank
    if center[0] < 0 or center[1] < 0:
        return blank
    xa, xb = int(center[0] - cw // 2) + cw, int(center[0] + cw // 2) + cw
    ya, yb = int(center[1] - ch // 2) + ch, int(center[1] + ch // 2) + ch
    padded = cv2.copyMakeBorder(frame, ch, ch, cw, cw, cv2.BORDER_CONSTANT, 0)
    rot = cv2.getRotationMatrix2D((cw // 2, ch // 2), angle, 1)
    return cv2.warpAffine(padded[ya:yb, xa:xb], rot, (cw, ch))


def rotation_coeffs(angle_deg: float, cw: int, ch: int) -> Tuple[float, ...]:
    """cv::getRotationMatrix2D((cw//2, ch//2), angle, 1) followed by the inversion cv::warpAffine
    applies when WARP_INVERSE_MAP is not set; float64 throughout, same operation order."""
    cx, cy = float(cw // 2), float(ch // 2)
    rad = angle_deg * (np.pi / 180.0)
    alpha, beta = float(np.cos(rad)), float(np.sin(rad))
    m00, m01, m02 = alpha, beta, (1 - alpha) * cx - beta * cy
    m10, m11, m12 = -beta, alpha, beta * cx + (1 - alpha) * cy
    det = m00 * m11 - m01 * m10
    det = 1.0 / det if det != 0 else 0.0
    a11, a22 = m11 * det, m00 * det
    i00, i01, i10, i11 = a11, m01 * (-det), m10 * (-det), a22
    b1 = -i00 * m02 - i01 * m12
    b2 = -i10 * m02 - i11 * m12
    return i00, i01, b1, i10, i11, b2


def crop_rotate_np(frame: np.ndarray, center: Sequence[float], angle: float, crop: Tuple[int, int] = (80, 80)):
    """Restatement of `crop_rotate_cv2` with OpenCV's 8-bit INTER_LINEAR fixed-point scheme
    (AB_BITS 10, INTER_BITS 5, coefficient scale 2^15) gathered straight from the un-padded frame."""
    cw, ch = int(crop[0]), int(crop[1])
    out = np.zeros((ch, cw), dtype=np.uint8)
    if np.isnan(angle) or np.isnan(center[0]) or np.isnan(center[1]) or center[0] < 0 or center[1] < 0:
        return np.zeros((cw, ch), dtype=np.uint8)
    ox, oy = int(center[0] - cw // 2), int(center[1] - ch // 2)        # sub-image origin in frame px
    sw = int(center[0] + cw // 2) - ox                                 # sub-image size (79 or 80 ...)
    sh = int(center[1] + ch // 2) - oy
    # numpy slicing clips the padded frame: the sub-image cannot extend past the padded canvas
    H, W = frame.shape
    sw = min(sw, W + cw - ox)
    sh = min(sh, H + ch - oy)
    if sw <= 0 or sh <= 0:
        return np.zeros((cw, ch), dtype=np.uint8)
    i00, i01, b1, i10, i11, b2 = rotation_coeffs(angle, cw, ch)
    xs = np.arange(cw, dtype=np.float64)
    adelta = np.rint(i00 * xs * 1024).astype(np.int64)
    bdelta = np.rint(i10 * xs * 1024).astype(np.int64)
    f = frame.astype(np.int64)
    for y in range(ch):
        X0 = int(np.rint((i01 * y + b1) * 1024)) + 16
        Y0 = int(np.rint((i11 * y + b2) * 1024)) + 16
        X = (X0 + adelta) >> 5
        Y = (Y0 + bdelta) >> 5
        sx, sy = X >> 5, Y >> 5
        fx, fy = X & 31, Y & 31
        acc = np.zeros((cw,), dtype=np.int64)
        for dy, dx, wgt in ((0, 0, (32 - fx) * (32 - fy)), (0, 1, fx * (32 - fy)),
                            (1, 0, (32 - fx) * fy), (1, 1, fx * fy)):
            px, py = sx + dx, sy + dy
            ok = (px >= 0) & (px < sw) & (py >= 0) & (py < sh)
            gx, gy = px + ox, py + oy
            ok &= (gx >= 0) & (gx < W) & (gy >= 0) & (gy < H)
            vals = np.where(ok, f[np.clip(gy, 0, H - 1), np.clip(gx, 0, W - 1)], 0)
            acc += vals * wgt * 32
        out[y] = ((acc + 16384) >> 15).astype(np.uint8)
    return out


# ------------------------------------------------------------------------------------------
# a4  mask paste (detectron2 detector_postprocess)             ref: model/util.py:45-62
# ------------------------------------------------------------------------------------------
def paste_masks_np(soft: np.ndarray, boxes: np.ndarray, height: int, width: int, threshold: float = 0.5):
    """detectron2.layers.mask_ops._do_paste_mask + `>= threshold` (third-party, unpinned, absent
    here -- restated from the published algorithm: sample the MxM soft mask at pixel centres
    mapped into the box, bilinear, zero outside, align_corners=False; float32 arithmetic)."""
    n, M = soft.shape[0], soft.shape[-1]
    out = np.zeros((n, height, width), dtype=bool)
    ys = np.arange(height, dtype=np.float32) + np.float32(0.5)
    xs = np.arange(width, dtype=np.float32) + np.float32(0.5)
    for i in range(n):
        x0, y0, x1, y1 = (np.float32(v) for v in boxes[i])
        gy = (ys - y0) / (y1 - y0) * np.float32(2) - np.float32(1)
        gx = (xs - x0) / (x1 - x0) * np.float32(2) - np.float32(1)
        # grid_sample, align_corners=False: pixel = ((g + 1) * M - 1) / 2
        py = ((gy + np.float32(1)) * np.float32(M) - np.float32(1)) / np.float32(2)
        px = ((gx + np.float32(1)) * np.float32(M) - np.float32(1)) / np.float32(2)
        y_lo = np.floor(py).astype(np.int64)
        x_lo = np.floor(px).astype(np.int64)
        wy1 = (py - y_lo.astype(np.float32)).astype(np.float32)
        wx1 = (px - x_lo.astype(np.float32)).astype(np.float32)
        wy0 = ((y_lo.astype(np.float32) + np.float32(1)) - py).astype(np.float32)     # torch: (iy_se - iy)
        wx0 = ((x_lo.astype(np.float32) + np.float32(1)) - px).astype(np.float32)
        src = soft[i].astype(np.float32)

        def tap(yy, xx):
            ok = (yy[:, None] >= 0) & (yy[:, None] < M) & (xx[None, :] >= 0) & (xx[None, :] < M)
            v = src[np.clip(yy, 0, M - 1)[:, None], np.clip(xx, 0, M - 1)[None, :]]
            return np.where(ok, v, np.float32(0))

        # torch's grid_sampler accumulates the four taps in the order nw, ne, sw, se
        val = tap(y_lo, x_lo) * (wx0[None, :] * wy0[:, None])
        val = val + tap(y_lo, x_lo + 1) * (wx1[None, :] * wy0[:, None])
        val = val + tap(y_lo + 1, x_lo) * (wx0[None, :] * wy1[:, None])
        val = val + tap(y_lo + 1, x_lo + 1) * (wx1[None, :] * wy1[:, None])
        out[i] = val >= np.float32(threshold)
    return out


# ------------------------------------------------------------------------------------------
# instances_to_features (non-tracking) + ProcessFeaturesStep glue
#                     ref: proc/proc.py:700-848 and pipeline/process_features_step.py:163-199
# ------------------------------------------------------------------------------------------
def bground_im(frames: np.ndarray, med_scale: int = 5) -> np.ndarray:
    """ref: proc/roi.py:293-307 -- cv2.medianBlur per frame (on a copy: the reference overwrites its input), then
    np.median over the frame axis."""
    blurred = np.stack([cv2.medianBlur(np.ascontiguousarray(f), med_scale) for f in frames])
    return np.median(blurred, axis=0)


def bground_im_np(frames: np.ndarray, med_scale: int = 5) -> np.ndarray:
    """The same without OpenCV: replicate border, middle of the sorted window (pins what medianBlur computes)."""
    r = med_scale // 2
    n, h, w = frames.shape
    p = np.pad(frames, ((0, 0), (r, r), (r, r)), mode='edge')
    win = np.stack([p[:, i:i + h, j:j + w] for i in range(med_scale) for j in range(med_scale)], axis=0)
    blurred = np.sort(win, axis=0)[(med_scale * med_scale) // 2]
    return np.median(blurred, axis=0)


def extract_chunk(chunk_u8: np.ndarray, masks: np.ndarray, keypoints: np.ndarray, num_instances: np.ndarray,
                  min_height: float = 0, max_height: float = 100, true_depth: float = 673.0,
                  crop: Tuple[int, int] = (80, 80), use_cv2: bool = True) -> dict:
    """Everything ProcessFeaturesStep.process does for one chunk with use_tracking=False and <=1
    instance per frame.  `masks` (N,h,w) u8 and `keypoints` (N,8,3) (NaN where no instance) are
    what mask_and_keypoints_from_model_output (ref: proc/proc.py:657-685) hands over."""
    kp = keypoints.astype(np.float64)
    cleaned = clean_frames_cv2(chunk_u8) if use_cv2 else clean_frames_np(chunk_u8)
    feats = frame_features_cv2(cleaned, masks) if use_cv2 else frame_features_np(cleaned, masks)
    with np.errstate(invalid='ignore'):
        lengths = np.max(feats['axis_length'], axis=1)
        angles = clamp_deg(-np.rad2deg(feats['orientation']))
    flips, _conf = keypoint_flips(kp, feats['centroid'], angles, lengths)
    angles = angles.copy()
    angles[flips] += 180
    angles, filt_flips, n_pass = iterative_angle_filter(angles)
    flips = np.logical_xor(flips, filt_flips)
    feats = dict(feats, orientation=np.array(angles))
    scalars = scalar_table(chunk_u8 * masks, feats['centroid'], feats['orientation'], feats['axis_length'],
                           min_height, max_height, true_depth)
    kp_table = keypoint_table(kp, cleaned, feats['centroid'], feats['orientation'], true_depth)
    n = chunk_u8.shape[0]
    depth_crops = np.zeros((n, crop[0], crop[1]), dtype=np.uint8)
    mask_crops = np.zeros((n, crop[0], crop[1]), dtype=np.uint8)
    fn = crop_rotate_cv2 if use_cv2 else crop_rotate_np
    for i in range(n):
        depth_crops[i] = fn(chunk_u8[i], feats['centroid'][i], feats['orientation'][i], crop)
        mask_crops[i] = fn(masks[i], feats['centroid'][i], feats['orientation'][i], crop)
    return {
        'cleaned_frames': cleaned, 'masks': masks, 'features': feats, 'flips': flips, 'keypoints': kp,
        'num_instances': num_instances, 'scalars': scalars, 'keypoint_table': kp_table,
        'depth_frames': depth_crops, 'mask_frames': mask_crops, 'filter_passes': n_pass,
    }


def nms_mask_instances(masks, scores, iou_threshold=0.5):
    """reference pipeline/process_features_step.py:63-113 (__nms_mask_instances), restated on numpy arrays: returns the
    indices (into `masks`) of the kept instances, best first.  Instances with empty masks are dropped first; then, round by
    round, the best remaining instance is kept and every remaining instance that is the lower-scored member of ANY pair with
    IoU above the threshold is removed together with it."""
    masks = np.asarray(masks).astype(bool)
    scores = np.asarray(scores)
    if len(masks) <= 1:
        return list(range(len(masks)))
    alive = np.flatnonzero(masks.reshape(len(masks), -1).any(axis=1))
    flat = masks[alive].reshape(len(alive), -1).astype(np.int64)
    idxs = np.argsort(scores[alive])
    pick = []
    while len(idxs) > 0:
        last = len(idxs) - 1
        pick.append(int(alive[idxs[last]]))
        m = flat[idxs]
        inter = m @ m.T
        areas = np.broadcast_to(m.sum(axis=1), inter.shape)
        ious = np.triu((inter / (areas + areas.T - inter)).astype(np.float32), k=1)
        drop = np.unique(np.concatenate(([last], np.where(ious > iou_threshold)[0])))
        idxs = np.delete(idxs, drop)
    return pick
