"""Lane-by-lane numpy emulation of csrc/features.cu (one warp = arrays of 32 uint32 lanes).

Test infrastructure: lets the CPU suite check the bit-parallel flood / peel / cell-sum algorithm of the
CUDA feature kernel against the oracle without a GPU.  Mirrors the kernel statement by statement."""
import numpy as np

U32 = np.uint32
FULL = U32(0xFFFFFFFF)


def brev(x):
    x = np.asarray(x, dtype=np.uint32)
    x = ((x >> U32(1)) & U32(0x55555555)) | ((x & U32(0x55555555)) << U32(1))
    x = ((x >> U32(2)) & U32(0x33333333)) | ((x & U32(0x33333333)) << U32(2))
    x = ((x >> U32(4)) & U32(0x0F0F0F0F)) | ((x & U32(0x0F0F0F0F)) << U32(4))
    x = ((x >> U32(8)) & U32(0x00FF00FF)) | ((x & U32(0x00FF00FF)) << U32(8))
    return (x >> U32(16)) | (x << U32(16))


def fill_up(seed, op):
    return (((op + seed) ^ op) & op) | seed


def fill_local(seed, op):
    return fill_up(seed, op) | brev(fill_up(brev(seed), brev(op)))


def trailing_ones(op):
    return op & ~(op + U32(1))


def leading_ones(op):
    return brev(trailing_ones(brev(op)))


def ballot(pred):
    return U32(int(sum((1 << i) for i in range(32) if pred[i])))


def brev1(x):
    return brev(np.array([x], dtype=np.uint32))[0]


def fill_row(seed, op):
    with np.errstate(over='ignore'):
        f = fill_local(seed & op, op)
        full = ballot(op == FULL)
        g_up = ballot((f >> U32(31)) != 0)
        g_dn = ballot((f & U32(1)) != 0)
        xu = g_up | full
        cin_up = (xu + g_up) ^ xu ^ g_up
        gr = brev1(g_dn)
        xr = gr | brev1(full)
        cin_dn = brev1((xr + gr) ^ xr ^ gr)
        lanes = np.arange(32, dtype=np.uint32)
        f = np.where((cin_up >> lanes) & U32(1), f | trailing_ones(op), f)
        f = np.where((cin_dn >> lanes) & U32(1), f | leading_ones(op), f)
    return f.astype(np.uint32)


def spread3(x):
    below = np.concatenate(([U32(0)], x[:-1]))
    above = np.concatenate((x[1:], [U32(0)]))
    return x | (x << U32(1)) | (x >> U32(1)) | (below >> U32(31)) | (above << U32(31))


CELL24 = [(24, 12, 12, 8, 6, 8), (12, 8, 8, 6, 5, 6), (12, 4, 8, 2, 3, 6), (12, 8, 4, 6, 3, 2), (12, 4, 4, 2, 1, 2)]


def pack_rows(fm):
    h, w = fm.shape
    rows = np.zeros((h, 32), dtype=np.uint32)
    for r in range(h):
        for x in np.flatnonzero(fm[r]):
            rows[r, x >> 5] |= U32(1 << (x & 31))
    return rows


def emulate_frame(fm):
    """fm: (h, w) bool, w <= 1024.  Returns the 6 int sums (x24) of the winning blob (zeros if none)."""
    h, w = fm.shape
    wpr = (w + 31) >> 5
    P = pack_rows(fm)
    Q = np.zeros_like(P)
    lane = np.arange(32)
    lane_mask = np.where(lane < wpr - 1, FULL,
                         np.where(lane == wpr - 1, U32(((1 << (w & 31)) - 1) if (w & 31) else 0xFFFFFFFF), U32(0))).astype(np.uint32)
    nz = np.flatnonzero(P.any(axis=1))
    best = [0] * 6
    have = False
    if nz.size == 0:
        return best
    r_lo, r_hi = int(nz[0]), int(nz[-1])
    edge = np.zeros(32, dtype=np.uint32)
    edge[0] |= U32(1)
    edge[(w - 1) >> 5] |= U32(1 << ((w - 1) & 31))
    down = True
    sweep = 0
    while True:
        changed = False
        prev = lane_mask.copy()
        rng = range(r_lo, r_hi + 1) if down else range(r_hi, r_lo - 1, -1)
        for r in rng:
            op = ~P[r] & lane_mask
            old = Q[r].copy()
            now = fill_row((prev | old | edge) & op, op)
            Q[r] = now
            changed |= bool((now != old).any())
            prev = now
        if not changed and sweep > 0:
            break
        down = not down
        sweep += 1
    for r in range(r_lo, r_hi + 1):
        Q[r] = ~Q[r] & lane_mask
    scan = r_lo
    while True:
        r0 = -1
        for r in range(scan, r_hi + 1):
            if Q[r].any():
                kw = int(np.flatnonzero(Q[r])[0])
                word = int(Q[r, kw])
                seed = np.zeros(32, dtype=np.uint32)
                seed[kw] = U32(word & -word)
                r0 = r
                break
        if r0 < 0:
            break
        scan = r0
        rmax = r0
        P[r0] = fill_row(seed, Q[r0])
        go_down = True
        sweep = 0
        while True:
            changed = False
            if go_down:
                prev = P[r0]
                for r in range(r0 + 1, r_hi + 1):
                    rem = Q[r]
                    old = P[r].copy() if r <= rmax else np.zeros(32, dtype=np.uint32)
                    now = fill_row((spread3(prev) | old) & rem, rem)
                    any_now = bool(now.any())
                    if not any_now and r > rmax:
                        break
                    P[r] = now
                    if any_now:
                        rmax = max(rmax, r)
                    changed |= bool((now != old).any())
                    prev = now
            else:
                prev = P[rmax]
                for r in range(rmax - 1, r0 - 1, -1):
                    rem = Q[r]
                    old = P[r].copy()
                    now = fill_row((spread3(prev) | old) & rem, rem)
                    P[r] = now
                    changed |= bool((now != old).any())
                    prev = now
            if sweep > 0 and not changed:
                break
            if sweep == 0 and rmax == r0:
                break
            go_down = not go_down
            sweep += 1
        s = [0] * 6
        for r in range(r0, rmax):
            for k in range(wpr):
                a, b = int(P[r, k]), int(P[r + 1, k])
                an = int(P[r, k + 1]) if k + 1 < wpr else 0
                bn = int(P[r + 1, k + 1]) if k + 1 < wpr else 0
                if (a | b) == 0:
                    continue
                a1 = ((a >> 1) | (an << 31)) & 0xFFFFFFFF
                b1 = ((b >> 1) | (bn << 31)) & 0xFFFFFFFF
                na, na1, nb, nb1 = ~a & 0xFFFFFFFF, ~a1 & 0xFFFFFFFF, ~b & 0xFFFFFFFF, ~b1 & 0xFFFFFFFF
                cls = [a & a1 & b & b1, na & a1 & b & b1, a & na1 & b & b1, a & a1 & nb & b1, a & a1 & b & nb1]
                for c in range(5):
                    bits = [i for i in range(32) if (cls[c] >> i) & 1]
                    if not bits:
                        continue
                    cnt = len(bits)
                    x0 = 32 * k
                    si = sum(x0 + i for i in bits)
                    sii = sum((x0 + i) ** 2 for i in bits)
                    A, U, V, UU, UV, VV = CELL24[c]
                    s[0] += A * cnt
                    s[1] += A * si + U * cnt
                    s[2] += A * r * cnt + V * cnt
                    s[3] += A * sii + 2 * U * si + UU * cnt
                    s[4] += A * r * si + V * si + U * r * cnt + UV * cnt
                    s[5] += A * r * r * cnt + 2 * V * r * cnt + VV * cnt
        if (not have) or s[0] >= best[0]:
            best, have = s, True
        for r in range(r0, rmax + 1):
            Q[r] &= ~P[r]
    return best
