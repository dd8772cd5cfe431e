"""TEST INFRASTRUCTURE -- restatement of cv2.inpaint(src, mask, 3, cv2.INPAINT_NS) for 8-bit single-channel images
(OpenCV photo/src/inpaint.cpp: icvInpaint + icvNSInpaintFMM + FastMarching_solve + CvPriorityQueueFloat).
Used by the reference at proc/proc.py:209.  Pinned against cv2.inpaint itself in tests/test_oracle_vs_golden.py.
Pure-Python loops: small cases only."""
import numpy as np

KNOWN, BAND, INSIDE = 0, 1, 2
f32 = np.float32


def _solve(i1, j1, i2, j2, f, t):
    a11, a22 = float(t[i1, j1]), float(t[i2, j2])
    m12 = min(a11, a22)
    if f[i1, j1] != INSIDE:
        if f[i2, j2] != INSIDE:
            if abs(a11 - a22) >= 1.0:
                sol = 1 + m12
            else:
                sol = (a11 + a22 + np.sqrt(2 - (a11 - a22) * (a11 - a22))) * 0.5
        else:
            sol = 1 + a11
    elif f[i2, j2] != INSIDE:
        sol = 1 + a22
    else:
        sol = 1 + m12
    return f32(sol)


def inpaint_ns(src: np.ndarray, mask: np.ndarray, radius: int = 3) -> np.ndarray:
    rows, cols = src.shape
    er, ec = rows + 2, cols + 2
    out = src.astype(np.int64).copy()
    f = np.zeros((er, ec), dtype=np.uint8)
    f[1:-1, 1:-1][mask != 0] = INSIDE
    t = np.full((er, ec), 1.0e6, dtype=np.float32)
    # band = dilate(mask, 3x3 cross) - mask, border cleared
    ins = f == INSIDE
    dil = ins.copy()
    dil[1:, :] |= ins[:-1, :]
    dil[:-1, :] |= ins[1:, :]
    dil[:, 1:] |= ins[:, :-1]
    dil[:, :-1] |= ins[:, 1:]
    band = dil & ~ins
    band[0, :] = band[-1, :] = False
    band[:, 0] = band[:, -1] = False
    if not dil.any():
        return src.copy()
    t[band] = 0
    # priority list: stable insertion after all elements with T <= new T
    heap = [(0.0, int(i), int(j)) for i, j in zip(*np.nonzero(band))]

    def push(i, j, T):
        pos = len(heap)
        while pos > 0 and heap[pos - 1][0] > T:
            pos -= 1
        heap.insert(pos, (T, i, j))

    R2 = radius * radius
    while heap:
        _, ii, jj = heap.pop(0)
        f[ii, jj] = KNOWN
        for (i, j) in ((ii - 1, jj), (ii, jj - 1), (ii + 1, jj), (ii, jj + 1)):
            if i <= 0 or j <= 0 or i > er - 1 or j > ec - 1:
                continue
            if f[i, j] != INSIDE:
                continue
            dist = min(_solve(i - 1, j, i, j - 1, f, t), _solve(i + 1, j, i, j - 1, f, t),
                       _solve(i - 1, j, i, j + 1, f, t), _solve(i + 1, j, i, j + 1, f, t))
            t[i, j] = dist
            Ia, s = f32(0), f32(1.0e-20)
            for k in range(i - radius, i + radius + 1):
                km = k - 1 + (k == 1)
                kp = k - 1 - (k == er - 2)
                for l in range(j - radius, j + radius + 1):
                    lm = l - 1 + (l == 1)
                    lp = l - 1 - (l == ec - 2)
                    if not (k > 0 and l > 0 and k < er - 1 and l < ec - 1):
                        continue
                    if f[k, l] == INSIDE or (l - j) * (l - j) + (k - i) * (k - i) > R2:
                        continue
                    ry, rx = f32(i - k), f32(j - l)
                    len_r = f32(rx * rx + ry * ry)
                    dst = f32(f32(1) / f32(len_r * len_r + f32(1)))
                    if f[k + 1, l] != INSIDE:
                        if f[k - 1, l] != INSIDE:
                            gx = f32(abs(out[kp + 1, lm] - out[kp, lm]) + abs(out[kp, lm] - out[km - 1, lm]))
                        else:
                            gx = f32(f32(abs(out[kp + 1, lm] - out[kp, lm])) * f32(2))
                    else:
                        if f[k - 1, l] != INSIDE:
                            gx = f32(f32(abs(out[kp, lm] - out[km - 1, lm])) * f32(2))
                        else:
                            gx = f32(0)
                    if f[k, l + 1] != INSIDE:
                        if f[k, l - 1] != INSIDE:
                            gy = f32(abs(out[km, lp + 1] - out[km, lm]) + abs(out[km, lm] - out[km, lm - 1]))
                        else:
                            gy = f32(f32(abs(out[km, lp + 1] - out[km, lm])) * f32(2))
                    else:
                        if f[k, l - 1] != INSIDE:
                            gy = f32(f32(abs(out[km, lm] - out[km, lm - 1])) * f32(2))
                        else:
                            gy = f32(0)
                    gx = f32(-gx)
                    dot = f32(f32(rx * gx) + f32(ry * gy))
                    if abs(dot) <= 0.01:
                        dirv = f32(0.000001)
                    else:
                        len_g = f32(f32(gx * gx) + f32(gy * gy))
                        dirv = f32(abs(dot / np.sqrt(f32(len_r * len_g))))
                    w = f32(dst * dirv)
                    Ia = f32(Ia + f32(w * f32(out[k - 1, l - 1])))
                    s = f32(s + w)
            val = float(Ia) / float(s)
            out[i - 1, j - 1] = int(min(255, max(0, np.rint(val))))
            f[i, j] = BAND
            push(i, j, float(dist))
    return out.astype(np.uint8)
