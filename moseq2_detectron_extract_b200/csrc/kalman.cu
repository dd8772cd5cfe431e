// a14 tracking branch: linear-Gaussian Kalman filter, RTS smoother and EM covariance estimation on the GPU.
//   ref proc/kalman.py:281-418 (KalmanTracker: initialize/em, smooth_update, filter_update, sample) and the
//   third-party pykalman "standard" filter it drives; proc/proc.py:730-826 (the per-frame angle heuristic).
//
// Model: x[t+1] = A x[t] + N(0,Q),  z[t] = H x[t] + N(0,R), time-invariant, S <= 64 states, O <= 32 observations
// (the reference tracks centroid + 8 keypoints at order 3: S = 54, O = 18; and an angle as (sin, cos): S = 6, O = 2).
// pykalman conventions kept: the first predicted state is (m0, P0); an observation with ANY non-finite component is
// skipped; smoother gains J[t] = Pf[t] A^T inv(Pp[t+1]).  A and H are block-sparse (3 and 1 non-zeros per row for the
// reference's trackers), so every product with them runs over per-row non-zero lists built in shared memory.
//
// Work split (float64 throughout):
//   kalman_filter_kernel    1 CTA   the Riccati recursion is sequential in t; ~100 K FMA per step from shared memory
//   kalman_gain_kernel      T-1 CTAs the gains depend on forward results only, so the S x S inversions (the expensive
//                                   part of a smoother) run in parallel over t, one CTA each
//   kalman_backward_kernel  1 CTA   mean recursion (mat-vec per step, next gain prefetched); with `want_cov` also the
//                                   smoothed covariances and lag-one covariances that EM needs
//   kalman_em_reduce/finish         the M-step: sums over t in parallel, then a handful of small products
//   track_angles_kernel     1 thread the reference's per-frame loop (predict, compare, maybe flip, filter_update) on the
//                                   6-state angle filter; sequential by construction
#include "common.cuh"
#include "angles.cuh"
#include <math.h>

namespace msq {
namespace {

constexpr int kKalThreads = 512;      // one CTA is alone on its SM in the sequential kernels: more warps hide the latencies
constexpr int kMaxS = 64, kMaxO = 32;

__device__ __forceinline__ bool finite_f64(double v) { return fabs(v) <= 1.79769313486231570e308; }   // false for NaN/Inf

// ---- per-row non-zero lists of a small dense matrix (rows x cols, row-major), ELL layout: every row holds `width`
// (= the largest non-zero count of any row) entries, short rows padded with (col 0, val 0.0) -------------------------
struct SparseRows {
    int *width;      // [1]
    int *col;        // [rows * cols] capacity, rows * width used
    double *val;     // [rows * cols] capacity
    int cols;
};
__device__ void build_sparse(const double *__restrict__ M, int rows, int cols, SparseRows sp) {
    if (threadIdx.x == 0) *sp.width = 1;
    __syncthreads();
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        int c = 0;
        for (int k = 0; k < cols; ++k) c += M[r * cols + k] != 0.0;
        atomicMax(sp.width, c);
    }
    __syncthreads();
    const int wd = *sp.width;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        int c = 0;
        for (int k = 0; k < cols; ++k) {
            const double v = M[r * cols + k];
            if (v != 0.0) { sp.col[r * wd + c] = k; sp.val[r * wd + c] = v; ++c; }
        }
        for (; c < wd; ++c) { sp.col[r * wd + c] = 0; sp.val[r * wd + c] = 0.0; }
    }
    __syncthreads();
}
__host__ __device__ constexpr size_t up16(size_t v) { return (v + 15) / 16 * 16; }
__device__ SparseRows carve_sparse(unsigned char *&p, int rows, int cols) {
    SparseRows sp;
    sp.val = reinterpret_cast<double *>(p); p += (size_t)rows * cols * sizeof(double);
    sp.col = reinterpret_cast<int *>(p); p += up16((size_t)rows * cols * sizeof(int));
    sp.width = reinterpret_cast<int *>(p); p += up16((size_t)rows * sizeof(int));
    sp.cols = cols;
    return sp;
}

// Loop over the elements e = threadIdx.x, +blockDim.x, ... of an (n_rows x n_cols) matrix with (i, j) kept incrementally
// (ncu on the first version of these kernels: a quarter of all instructions were the e / n_cols divisions).
#define MSQ_FOR_ELEMENTS(e, i, j, n_rows, n_cols)                                                              \
    for (int e = threadIdx.x, i = e / (n_cols), j = e - i * (n_cols), _di = (int)blockDim.x / (n_cols),           \
             _dj = (int)blockDim.x - _di * (n_cols);                                                              \
         e < (n_rows) * (n_cols); e += blockDim.x, i += _di, j += _dj, i += (j >= (n_cols)), j -= (j >= (n_cols)) ? (n_cols) : 0)

// C[n x m] = Asp[n x k] * B[k x m]   (B, C dense in shared memory, ld = m)
template <int W>
__device__ __forceinline__ void spmm_left_w(double *__restrict__ C, const SparseRows A, const double *__restrict__ B, int n, int m,
                                            const double *__restrict__ add) {
    const int wd = W > 0 ? W : *A.width;
    MSQ_FOR_ELEMENTS(e, i, j, n, m) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < wd; ++q) acc = fma(A.val[i * wd + q], B[A.col[i * wd + q] * m + j], acc);
        C[e] = add ? acc + add[e] : acc;
    }
}
__device__ void spmm_left(double *C, const SparseRows A, const double *B, int n, int m, const double *add = nullptr) {
    switch (*A.width) {
        case 1: spmm_left_w<1>(C, A, B, n, m, add); break;
        case 2: spmm_left_w<2>(C, A, B, n, m, add); break;
        case 3: spmm_left_w<3>(C, A, B, n, m, add); break;
        case 4: spmm_left_w<4>(C, A, B, n, m, add); break;
        default: spmm_left_w<0>(C, A, B, n, m, add); break;
    }
}
// C[n x r] = B[n x k] * Asp^T  (Asp is r x k)  + (add ? add[n x r] : 0)
template <int W>
__device__ __forceinline__ void spmm_right_t_w(double *__restrict__ C, const double *__restrict__ B, const SparseRows A, int n, int r,
                                               const double *__restrict__ add) {
    const int k = A.cols;
    const int wd = W > 0 ? W : *A.width;
    MSQ_FOR_ELEMENTS(e, i, j, n, r) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < wd; ++q) acc = fma(B[i * k + A.col[j * wd + q]], A.val[j * wd + q], acc);
        C[e] = add ? acc + add[e] : acc;
    }
}
__device__ void spmm_right_t(double *C, const double *B, const SparseRows A, int n, int r, const double *add) {
    switch (*A.width) {
        case 1: spmm_right_t_w<1>(C, B, A, n, r, add); break;
        case 2: spmm_right_t_w<2>(C, B, A, n, r, add); break;
        case 3: spmm_right_t_w<3>(C, B, A, n, r, add); break;
        case 4: spmm_right_t_w<4>(C, B, A, n, r, add); break;
        default: spmm_right_t_w<0>(C, B, A, n, r, add); break;
    }
}
// y[n] = Asp x
__device__ void spmv(double *y, const SparseRows A, const double *x, int n) {
    const int wd = *A.width;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double acc = 0.0;
        for (int q = 0; q < wd; ++q) acc = fma(A.val[i * wd + q], x[A.col[i * wd + q]], acc);
        y[i] = acc;
    }
}

// Dense C[n x m] (op)= A[n x k] * B   with B given as [k x m] (transB = false) or [m x k] (transB = true); 2x3 register
// tiles (486 tiles for 54 x 54: every thread of the CTA has one).  op: 0 store, +1 add to C, -1 subtract from C.
// fma() explicitly: the library is built with --fmad=false for the bit-exact float64 epilogues elsewhere.
__device__ void mm_dense(double *C, const double *A, const double *B, int n, int k, int m, bool transB, int op) {
    constexpr int TI = 2, TJ = 3;
    const int ti_n = (n + TI - 1) / TI, tj_n = (m + TJ - 1) / TJ;
    for (int tile = threadIdx.x; tile < ti_n * tj_n; tile += blockDim.x) {
        const int ti = tile / tj_n, tj = tile - ti * tj_n;
        const int i0 = ti * TI, j0 = tj * TJ;
        double acc[TI][TJ] = {};
        int ia[TI], jb[TJ];
#pragma unroll
        for (int a = 0; a < TI; ++a) ia[a] = min(i0 + a, n - 1) * k;
#pragma unroll
        for (int b = 0; b < TJ; ++b) jb[b] = transB ? min(j0 + b, m - 1) * k : min(j0 + b, m - 1);
        for (int kk = 0; kk < k; ++kk) {
            double av[TI], bv[TJ];
#pragma unroll
            for (int a = 0; a < TI; ++a) av[a] = A[ia[a] + kk];
#pragma unroll
            for (int b = 0; b < TJ; ++b) bv[b] = transB ? B[jb[b] + kk] : B[kk * m + jb[b]];
#pragma unroll
            for (int a = 0; a < TI; ++a)
#pragma unroll
                for (int b = 0; b < TJ; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < TI; ++a)
#pragma unroll
            for (int b = 0; b < TJ; ++b)
                if (i0 + a < n && j0 + b < m) {
                    double *c = C + (i0 + a) * m + j0 + b;
                    *c = op == 0 ? acc[a][b] : (op > 0 ? *c + acc[a][b] : *c - acc[a][b]);
                }
    }
}

// In-place Gauss-Jordan inverse of the symmetric positive-definite n x n matrix M (no pivoting).  Step k divides row k by
// the pivot (its own column becoming 1 / pivot) and eliminates column k from every other row (leaving -f / pivot there):
// the columns of the inverse grow in the place of the eliminated ones.  colk / rowk: n doubles of scratch each.
__device__ void gauss_jordan_inplace(double *M, int n, double *colk, double *rowk) {
    // (i, j) of this thread's first element and the step to its next one, computed once for all n elimination steps
    const int i0 = threadIdx.x / n, j0 = threadIdx.x - i0 * n, di = (int)blockDim.x / n, dj = (int)blockDim.x - di * n;
    for (int k = 0; k < n; ++k) {
        const double pivot = M[k * n + k];
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            rowk[j] = (j == k ? 1.0 : M[k * n + j]) / pivot;
            colk[j] = (j == k) ? 0.0 : M[j * n + k];
        }
        __syncthreads();
        for (int e = threadIdx.x, i = i0, j = j0; e < n * n; e += blockDim.x) {
            M[e] = (i == k) ? rowk[j] : fma(-colk[i], rowk[j], j == k ? 0.0 : M[e]);
            i += di; j += dj;
            if (j >= n) { j -= n; ++i; }
        }
        __syncthreads();
    }
}

// The same elimination for small fixed sizes inside ONE warp: lane r keeps row r in registers, the scaled pivot row is
// broadcast by shuffles -- no shared-memory traffic and no block barriers (the block version above needs 2 n of them, which
// was two thirds of a filter step for the 18 x 18 innovation covariance).  All 32 lanes of the calling warp must take part.
template <int N>
__device__ void gauss_jordan_warp(double *M) {
    const int lane = threadIdx.x & 31;
    double row[N];
#pragma unroll
    for (int j = 0; j < N; ++j) row[j] = lane < N ? M[lane * N + j] : 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double inv = 1.0 / __shfl_sync(0xffffffffu, row[k], k);
        const double f = (lane == k) ? 0.0 : row[k];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double rkj = __shfl_sync(0xffffffffu, (j == k ? 1.0 : row[j]) * inv, k);     // scaled pivot row
            row[j] = (lane == k) ? rkj : fma(-f, rkj, j == k ? 0.0 : row[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < N; ++j)
        if (lane < N) M[lane * N + j] = row[j];
}
// inverse of the O x O innovation covariance: warp version for the sizes the reference's trackers have, block version otherwise
__device__ void invert_innovation(double *Sm, int O, double *colk, double *rowk) {
    if (O == 18 || O == 2) {
        if (threadIdx.x < 32) {
            if (O == 18) gauss_jordan_warp<18>(Sm); else gauss_jordan_warp<2>(Sm);
        }
        __syncthreads();
    } else {
        gauss_jordan_inplace(Sm, O, colk, rowk);
    }
}

// =====================================================================================================================
// forward filter
// =====================================================================================================================
struct FilterArgs {
    const double *A, *H, *Q, *R, *m0, *P0, *obs;
    int T, S, O, predict_first;
    double *mp, *mf, *Pp, *Pf;       // [T][S], [T][S][S]
    unsigned char *valid;            // [T]
};

__global__ void __launch_bounds__(kKalThreads)
kalman_filter_kernel(FilterArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.S, O = a.O, T = a.T;
    unsigned char *p = smem_raw;
    double *P = reinterpret_cast<double *>(p); p += (size_t)S * S * 8;
    double *tmp = reinterpret_cast<double *>(p); p += (size_t)S * S * 8;
    double *Q = reinterpret_cast<double *>(p); p += (size_t)S * S * 8;
    double *Pps = reinterpret_cast<double *>(p); p += (size_t)S * S * 8;      // the last predicted covariance
    double *PHt = reinterpret_cast<double *>(p); p += (size_t)S * O * 8;
    double *K = reinterpret_cast<double *>(p); p += (size_t)S * O * 8;
    double *Sm = reinterpret_cast<double *>(p); p += (size_t)O * 2 * O * 8;
    double *R = reinterpret_cast<double *>(p); p += (size_t)O * O * 8;
    double *m = reinterpret_cast<double *>(p); p += (size_t)S * 8;
    double *m2 = reinterpret_cast<double *>(p); p += (size_t)S * 8;
    double *y = reinterpret_cast<double *>(p); p += (size_t)O * 8;
    double *colk = reinterpret_cast<double *>(p); p += (size_t)O * 8;
    double *rowk = reinterpret_cast<double *>(p); p += (size_t)2 * O * 8;
    SparseRows As = carve_sparse(p, S, S);
    SparseRows Hs = carve_sparse(p, O, S);

    build_sparse(a.A, S, S, As);
    build_sparse(a.H, O, S, Hs);
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) { P[e] = a.P0[e]; Q[e] = a.Q[e]; }
    for (int e = threadIdx.x; e < O * O; e += blockDim.x) R[e] = a.R[e];
    for (int e = threadIdx.x; e < S; e += blockDim.x) m[e] = a.m0[e];
    __syncthreads();

    // thread e < O keeps component e of the observation in a register; the next step's value is requested right after the
    // vote so that its global-memory latency is off the critical path of the (sequential) step
    double z_cur = (threadIdx.x < O) ? a.obs[threadIdx.x] : 0.0;
    // The covariances do not depend on the observed values, only on which steps have an observation, and the Riccati
    // recursion converges geometrically (for the reference's trackers: to rounding level within ~30 observed steps).  Once
    // the predicted covariance repeats to 1e-13 of its scale over two observed steps, further observed steps reuse P_pred,
    // K and P_filt and only move the mean -- until an observation is missing, which restarts the full recursion.
    bool steady = false, prev_valid = false;
    for (int t = 0; t < T; ++t) {
        // pykalman: any masked component -> the whole observation is skipped
        const int ok = finite_f64(z_cur) ? 1 : 0;
        const double z_now = z_cur;
        const int valid = __syncthreads_and(ok);         // also separates this step from the previous one
        if (threadIdx.x < O && t + 1 < T) z_cur = a.obs[(size_t)(t + 1) * O + threadIdx.x];
        if (threadIdx.x == 0) a.valid[t] = (unsigned char)valid;
        const bool predict = t > 0 || a.predict_first;
        if (steady && valid && predict) {
            spmv(m2, As, m, S);                           // predicted mean
            __syncthreads();
            if (threadIdx.x < O) {                        // innovation z - H m_pred
                const int wd = *Hs.width;
                double hm = 0.0;
                for (int q = 0; q < wd; ++q) hm = fma(Hs.val[threadIdx.x * wd + q], m2[Hs.col[threadIdx.x * wd + q]], hm);
                y[threadIdx.x] = z_now - hm;
            }
            for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
                a.Pp[(size_t)t * S * S + e] = Pps[e];
                a.Pf[(size_t)t * S * S + e] = P[e];
            }
            for (int e = threadIdx.x; e < S; e += blockDim.x) a.mp[(size_t)t * S + e] = m2[e];
            __syncthreads();
            for (int i = threadIdx.x; i < S; i += blockDim.x) {
                double acc = 0.0;
                for (int q = 0; q < O; ++q) acc = fma(K[i * O + q], y[q], acc);
                const double v = m2[i] + acc;
                m[i] = v;
                a.mf[(size_t)t * S + i] = v;
            }
            prev_valid = true;
            continue;
        }
        int moved = 1;
        if (predict) {
            spmv(m2, As, m, S);
            spmm_left(tmp, As, P, S, S);                 // A P
            __syncthreads();
            spmm_right_t(P, tmp, As, S, S, Q);           // (A P) A^T + Q
            for (int e = threadIdx.x; e < S; e += blockDim.x) m[e] = m2[e];
            __syncthreads();
            // has the predicted covariance stopped moving?  (element scale: sqrt(P_ii P_jj) bounds |P_ij|)
            moved = 0;
            MSQ_FOR_ELEMENTS(e, i, j, S, S) {
                const double pe = P[e];
                moved |= fabs(pe - Pps[e]) > 1e-13 * sqrt(fabs(P[i * S + i] * P[j * S + j]));
                Pps[e] = pe;
            }
        } else {
            for (int e = threadIdx.x; e < S * S; e += blockDim.x) Pps[e] = P[e];
        }
        for (int e = threadIdx.x; e < S * S; e += blockDim.x) a.Pp[(size_t)t * S * S + e] = P[e];
        for (int e = threadIdx.x; e < S; e += blockDim.x) a.mp[(size_t)t * S + e] = m[e];
        steady = false;
        if (valid) {
            spmm_right_t(PHt, P, Hs, S, O, nullptr);     // P H^T
            spmm_left(tmp, Hs, P, O, S);                 // H P   (O x S, in the scratch matrix)
            spmv(m2, Hs, m, O);                           // H m
            const int moved_any = __syncthreads_or(moved);
            steady = predict && prev_valid && !moved_any;
            if (threadIdx.x < O) y[threadIdx.x] = z_now - m2[threadIdx.x];
            spmm_left(Sm, Hs, PHt, O, O, R);             // H P H^T + R
            __syncthreads();
            invert_innovation(Sm, O, colk, rowk);
            mm_dense(K, PHt, Sm, S, O, O, false, 0);     // K = P H^T S^-1
            __syncthreads();
            for (int i = threadIdx.x; i < S; i += blockDim.x) {
                double acc = 0.0;
                for (int q = 0; q < O; ++q) acc = fma(K[i * O + q], y[q], acc);
                m[i] += acc;
            }
            // P -= K (H P).  NOT K (P H^T)^T: rounding leaves P slightly asymmetric, and with the transposed form the
            // antisymmetric part is multiplied by (I + K H) every step instead of (I - K H) -- it grows until overflow.
            mm_dense(P, K, tmp, S, O, S, false, -1);
            __syncthreads();
        }
        prev_valid = valid != 0;
        for (int e = threadIdx.x; e < S * S; e += blockDim.x) a.Pf[(size_t)t * S * S + e] = P[e];
        for (int e = threadIdx.x; e < S; e += blockDim.x) a.mf[(size_t)t * S + e] = m[e];
    }
}

// =====================================================================================================================
// smoother gains, one CTA per t:  J[t] = Pf[t] A^T inv(Pp[t+1])
// =====================================================================================================================
__global__ void __launch_bounds__(kKalThreads)
kalman_gain_kernel(const double *__restrict__ A, const double *__restrict__ Pf, const double *__restrict__ Pp, int S,
                   double *__restrict__ J) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *p = smem_raw;
    double *inv = reinterpret_cast<double *>(p); p += (size_t)S * S * 8;
    double *F = reinterpret_cast<double *>(p); p += (size_t)S * S * 8;
    double *FA = reinterpret_cast<double *>(p); p += (size_t)S * S * 8;
    double *colk = reinterpret_cast<double *>(p); p += (size_t)S * 8;
    double *rowk = reinterpret_cast<double *>(p); p += (size_t)S * 8;
    SparseRows As = carve_sparse(p, S, S);
    const int t = blockIdx.x;
    build_sparse(A, S, S, As);
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) {
        inv[e] = Pp[(size_t)(t + 1) * S * S + e];
        F[e] = Pf[(size_t)t * S * S + e];
    }
    __syncthreads();
    spmm_right_t(FA, F, As, S, S, nullptr);               // Pf A^T
    gauss_jordan_inplace(inv, S, colk, rowk);
    mm_dense(F, FA, inv, S, S, S, false, 0);
    __syncthreads();
    for (int e = threadIdx.x; e < S * S; e += blockDim.x) J[(size_t)t * S * S + e] = F[e];
}

// =====================================================================================================================
// backward pass (1 CTA)
// =====================================================================================================================
struct BackwardArgs {
    const double *mp, *mf, *Pp, *Pf, *J;
    int T, S, want_cov;
    double *ms;                      // [T][S]
    double *Ps, *pair;               // [T][S][S] (want_cov only)
};

__global__ void __launch_bounds__(kKalThreads)
kalman_backward_kernel(BackwardArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.S, T = a.T, SS = S * S;
    unsigned char *p = smem_raw;
    double *Jt = reinterpret_cast<double *>(p); p += (size_t)SS * 8;
    double *d = reinterpret_cast<double *>(p); p += (size_t)S * 8;
    double *msn = reinterpret_cast<double *>(p); p += (size_t)S * 8;
    double *msn2 = reinterpret_cast<double *>(p); p += (size_t)S * 8;
    double *mfs = reinterpret_cast<double *>(p); p += (size_t)S * 8;
    double *Psn = nullptr, *D = nullptr, *JD = nullptr;
    if (a.want_cov) {
        Psn = reinterpret_cast<double *>(p); p += (size_t)SS * 8;
        D = reinterpret_cast<double *>(p); p += (size_t)SS * 8;
        JD = reinterpret_cast<double *>(p); p += (size_t)SS * 8;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;

    for (int e = threadIdx.x; e < S; e += blockDim.x) { msn[e] = a.mf[(size_t)(T - 1) * S + e]; a.ms[(size_t)(T - 1) * S + e] = msn[e]; }
    if (a.want_cov) {
        for (int e = threadIdx.x; e < SS; e += blockDim.x) {
            Psn[e] = a.Pf[(size_t)(T - 1) * SS + e];
            a.Ps[(size_t)(T - 1) * SS + e] = Psn[e];
            a.pair[e] = 0.0;                                         // pair[0] is never defined
        }
    }
    // Everything step t reads from global memory (J[t], mp[t+1], mf[t], and for EM Pp[t+1], Pf[t]) is requested one step
    // ahead into registers: the recursion is sequential, so a global-memory latency on its critical path costs every step.
    constexpr int kPer = (kMaxS * kMaxS + kKalThreads - 1) / kKalThreads;
    double jreg[kPer], ppreg[kPer], pfreg[kPer], mp_r = 0.0, mf_r = 0.0;
    auto fetch = [&](int t) {
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const int e = threadIdx.x + q * kKalThreads;
            const bool on = t >= 0 && e < SS;
            jreg[q] = on ? a.J[(size_t)t * SS + e] : 0.0;
            if (a.want_cov) {
                ppreg[q] = on ? a.Pp[(size_t)(t + 1) * SS + e] : 0.0;
                pfreg[q] = on ? a.Pf[(size_t)t * SS + e] : 0.0;
            }
        }
        if (t >= 0 && threadIdx.x < S) {
            mp_r = a.mp[(size_t)(t + 1) * S + threadIdx.x];
            mf_r = a.mf[(size_t)t * S + threadIdx.x];
        }
    };
    fetch(T - 2);
    __syncthreads();
    for (int t = T - 2; t >= 0; --t) {
        double pf_now[kPer];
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const int e = threadIdx.x + q * kKalThreads;
            if (e < SS) {
                Jt[e] = jreg[q];
                if (a.want_cov) D[e] = Psn[e] - ppreg[q];
            }
            pf_now[q] = pfreg[q];
        }
        if (threadIdx.x < S) { d[threadIdx.x] = msn[threadIdx.x] - mp_r; mfs[threadIdx.x] = mf_r; }
        fetch(t - 1);
        __syncthreads();
        // ms[t] = mf[t] + J d : one warp per row, lanes stride the row
        for (int r = warp; r < S; r += warps) {
            double acc = 0.0;
            for (int c = lane; c < S; c += 32) acc = fma(Jt[r * S + c], d[c], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) {
                const double v = mfs[r] + acc;
                a.ms[(size_t)t * S + r] = v;
                msn2[r] = v;
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < S; e += blockDim.x) msn[e] = msn2[e];
        if (a.want_cov) {
            // pair[t+1] = Ps[t+1] J[t]^T
            mm_dense(JD, Psn, Jt, S, S, S, true, 0);
            __syncthreads();
            for (int e = threadIdx.x; e < SS; e += blockDim.x) a.pair[(size_t)(t + 1) * SS + e] = JD[e];
            __syncthreads();
            mm_dense(JD, Jt, D, S, S, S, false, 0);                   // J (Ps[t+1] - Pp[t+1])
#pragma unroll
            for (int q = 0; q < kPer; ++q) {
                const int e = threadIdx.x + q * kKalThreads;
                if (e < SS) Psn[e] = pf_now[q];                       // (Psn was last read by the first product)
            }
            __syncthreads();
            mm_dense(Psn, JD, Jt, S, S, S, true, +1);                 // Ps[t] = Pf[t] + (J D) J^T
            __syncthreads();
            for (int e = threadIdx.x; e < SS; e += blockDim.x) a.Ps[(size_t)t * SS + e] = Psn[e];
        }
        __syncthreads();
    }
}


// =====================================================================================================================
// EM M-step (pykalman _em for transition_covariance, observation_covariance, initial_state_covariance)
// =====================================================================================================================
// sums over t of the smoothed / lag-one covariances; one thread per matrix element, t ascending (deterministic)
__global__ void __launch_bounds__(128)
kalman_em_reduce_kernel(const double *__restrict__ Ps, const double *__restrict__ pair, const unsigned char *__restrict__ valid,
                        int T, int SS, double *__restrict__ sum_lo, double *__restrict__ sum_hi,
                        double *__restrict__ sum_pair, double *__restrict__ sum_valid) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= SS) return;
    double lo = 0.0, hi = 0.0, pr = 0.0, va = 0.0;
    for (int t = 0; t < T; ++t) {
        const double v = Ps[(size_t)t * SS + e];
        if (t < T - 1) lo += v;
        if (t > 0) { hi += v; pr += pair[(size_t)t * SS + e]; }
        if (valid[t]) va += v;
    }
    sum_lo[e] = lo; sum_hi[e] = hi; sum_pair[e] = pr; sum_valid[e] = va;
}

struct EmArgs {
    const double *A, *H, *m0, *obs, *ms, *Ps0;
    const unsigned char *valid;
    const double *sum_lo, *sum_hi, *sum_pair, *sum_valid;
    int T, S, O;
    double *Q, *R, *P0;              // updated in place
};

__global__ void __launch_bounds__(kKalThreads)
kalman_em_finish_kernel(EmArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.S, O = a.O, T = a.T, SS = S * S;
    unsigned char *p = smem_raw;
    double *X = reinterpret_cast<double *>(p); p += (size_t)SS * 8;
    double *Y = reinterpret_cast<double *>(p); p += (size_t)SS * 8;
    double *Z = reinterpret_cast<double *>(p); p += (size_t)SS * 8;
    double *err = reinterpret_cast<double *>(p); p += (size_t)S * 8;
    double *cur = reinterpret_cast<double *>(p); p += (size_t)S * 8;
    double *nxt = reinterpret_cast<double *>(p); p += (size_t)S * 8;
    double *ez = reinterpret_cast<double *>(p); p += (size_t)O * 8;
    SparseRows As = carve_sparse(p, S, S);
    SparseRows Hs = carve_sparse(p, O, S);
    build_sparse(a.A, S, S, As);
    build_sparse(a.H, O, S, Hs);
    __syncthreads();

    // ---- E = sum_t err err^T, err = ms[t+1] - A ms[t];  Ez = sum_valid (z - H ms)(z - H ms)^T ----------------------
    constexpr int kPer = (kMaxS * kMaxS + kKalThreads - 1) / kKalThreads;
    constexpr int kPerO = (kMaxO * kMaxO + kKalThreads - 1) / kKalThreads;
    double accE[kPer] = {}, accZ[kPerO] = {};
    int n_obs = 0;
    for (int t = 0; t < T; ++t) {
        for (int e = threadIdx.x; e < S; e += blockDim.x) { cur[e] = a.ms[(size_t)t * S + e]; nxt[e] = (t + 1 < T) ? a.ms[(size_t)(t + 1) * S + e] : 0.0; }
        __syncthreads();
        spmv(err, As, cur, S);
        spmv(ez, Hs, cur, O);
        __syncthreads();
        for (int e = threadIdx.x; e < S; e += blockDim.x) err[e] = nxt[e] - err[e];
        const bool v = a.valid[t] != 0;
        if (v) for (int e = threadIdx.x; e < O; e += blockDim.x) ez[e] = a.obs[(size_t)t * O + e] - ez[e];
        __syncthreads();
        if (t + 1 < T) {
#pragma unroll
            for (int q = 0; q < kPer; ++q) {
                const int e = threadIdx.x + q * kKalThreads;
                if (e < SS) accE[q] += err[e / S] * err[e % S];
            }
        }
        if (v) {
            ++n_obs;
#pragma unroll
            for (int q = 0; q < kPerO; ++q) {
                const int e = threadIdx.x + q * kKalThreads;
                if (e < O * O) accZ[q] += ez[e / O] * ez[e % O];
            }
        }
        __syncthreads();
    }
    // ---- Q = (E + A sum_lo A^T + sum_hi - V A^T - (V A^T)^T) / (T-1),  V = sum_pair ----------------------------------
    for (int e = threadIdx.x; e < SS; e += blockDim.x) { X[e] = a.sum_lo[e]; Z[e] = a.sum_pair[e]; }
    __syncthreads();
    spmm_left(Y, As, X, S, S);                            // A sum_lo
    __syncthreads();
    spmm_right_t(X, Y, As, S, S, nullptr);                // A sum_lo A^T
    __syncthreads();
    spmm_right_t(Y, Z, As, S, S, nullptr);                // V A^T
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int e = threadIdx.x + q * kKalThreads;
        if (e < SS) {
            const int i = e / S, j = e - i * S;
            a.Q[e] = (accE[q] + X[e] + a.sum_hi[e] - Y[e] - Y[j * S + i]) / (double)(T - 1);
        }
    }
    __syncthreads();
    // ---- R = (Ez + H sum_valid H^T) / n_obs ---------------------------------------------------------------------------
    for (int e = threadIdx.x; e < SS; e += blockDim.x) X[e] = a.sum_valid[e];
    __syncthreads();
    spmm_right_t(Y, X, Hs, S, O, nullptr);                // sum_valid H^T   (S x O)
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kPerO; ++q) {
        const int e = threadIdx.x + q * kKalThreads;
        if (e < O * O) {
            const int i = e / O, j = e - i * O;
            double acc = 0.0;
            const int c = *Hs.width;
            for (int k = 0; k < c; ++k) acc += Hs.val[i * c + k] * Y[Hs.col[i * c + k] * O + j];
            const double tot = accZ[q] + acc;
            a.R[e] = n_obs > 0 ? tot / (double)n_obs : tot;
        }
    }
    // ---- P0 = Ps[0] + x0 x0^T - mu x0^T - x0 mu^T + mu mu^T -----------------------------------------------------------
    for (int e = threadIdx.x; e < SS; e += blockDim.x) {
        const int i = e / S, j = e - i * S;
        const double x0i = a.ms[i], x0j = a.ms[j], mi = a.m0[i], mj = a.m0[j];
        a.P0[e] = (((a.Ps0[e] + x0i * x0j) - mi * x0j) - x0i * mj) + mi * mj;
    }
}

// =====================================================================================================================
// keypoint alignment scores (ref proc/proc.py:936-958) of the keypoints rotated into the egocentric frame
// =====================================================================================================================
__constant__ signed char kExpectedAlignment[7][7] = {      // ref proc/proc.py:961-984
    {0, 1, 1, 1, 1, 1, 1}, {-1, 0, 0, 1, 1, 1, 1}, {-1, 0, 0, 1, 1, 1, 1}, {-1, -1, -1, 0, 1, 1, 1},
    {-1, -1, -1, -1, 0, 0, 1}, {-1, -1, -1, -1, 0, 0, 1}, {-1, -1, -1, -1, -1, -1, 0}};

__device__ __forceinline__ double clamp_deg(double a) {            // ref proc/proc.py:688-691
    a = a < 0 ? 360.0 + a : a;
    double r = fmod(a, 360.0);
    if (r != 0.0 && r < 0) r += 360.0;
    return r;
}
__device__ __forceinline__ double angle_diff(double a1, double a2) {   // ref proc/kalman.py:93-98
    double d = fmod(a2 - a1, 360.0);
    if (d < 0) d += 360.0;
    return d > 180.0 ? -(360.0 - d) : d;
}

__device__ __forceinline__ double alignment_score(const double *__restrict__ kp, int kp_stride, double ox, double oy, double angle) {
    const double t = (-angle) * kPiOver180;
    const double c = cos(t), s = sin(t);
    double rx[7], ry;
#pragma unroll
    for (int k = 0; k < 7; ++k) rotate_about(kp[k * kp_stride], kp[k * kp_stride + 1], ox, oy, c, s, rx[k], ry);
    int met = 0, expectations = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int want = kExpectedAlignment[i][j];
            if (want == 0) continue;
            ++expectations;
            const double dist = rx[i] - rx[j];
            const int sign = dist > 0 ? 1 : (dist < 0 ? -1 : (dist == 0 ? 0 : 2));      // NaN never matches
            met += sign == want;
        }
    return (double)met / (double)expectations;
}

__global__ void __launch_bounds__(128)
keypoint_alignment_kernel(const double *__restrict__ kpts, int kp_stride, const double *__restrict__ centroid,
                          const double *__restrict__ angles, int n, double *__restrict__ scores) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    scores[f] = alignment_score(kpts + (size_t)f * 8 * kp_stride, kp_stride, centroid[2 * f], centroid[2 * f + 1], angles[f]);
}

// a8 + a9 of the tracking branch in one pass over the frames (ref proc/proc.py:720-724, 756-763): degrees + clamp,
// keypoint flip votes on the SMOOTHED float64 keypoints / centroids, `angles[flips] = clamp(angles + 180)`, and the
// alignment score of the keypoints rotated by the flipped angle.
__global__ void __launch_bounds__(128)
tracking_prepare_kernel(const double *__restrict__ orientation, const double *__restrict__ axis,
                        const double *__restrict__ centroid, const double *__restrict__ kpts, int n,
                        double *__restrict__ angles, uint8_t *__restrict__ flips, double *__restrict__ conf,
                        double *__restrict__ scores) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const double length = np_max2(axis[2 * f], axis[2 * f + 1]);
    double a = clamp_deg(-(orientation[f] * k180OverPi));
    double c;
    const bool flip = keypoint_flip_vote(kpts + (size_t)f * 24, centroid[2 * f], centroid[2 * f + 1], a, length, &c);
    if (flip) a = clamp_deg(a + 180.0);
    angles[f] = a;
    flips[f] = flip ? 1 : 0;
    if (conf) conf[f] = c;
    scores[f] = alignment_score(kpts + (size_t)f * 24, 3, centroid[2 * f], centroid[2 * f + 1], a);
}

// =====================================================================================================================
// the per-frame angle heuristic (ref proc/proc.py:771-796) on the angle tracker's small filter; one thread
// =====================================================================================================================
constexpr int kAngS = 8, kAngO = 2;     // (sin, cos) at order <= 4

struct AngleArgs {
    const double *A, *H, *Q, *R;
    double *mean, *cov;              // in/out: last filtered state
    double *angles;                  // in/out [n] degrees
    unsigned char *flips;            // in/out [n]
    const double *scores;            // [n] keypoint alignment scores
    int n, S, order;
};

// S is a template parameter so that the state lives in registers (with a run-time S the arrays went to local memory:
// 22 ms per 1000 frames); the constant model matrices and the last predicted covariance sit in shared memory.  The same
// steady-state shortcut as in kalman_filter_kernel: once the predicted covariance repeats over two observed frames, observed
// frames only move the mean (K, P_filt reused) until an angle is missing.
template <int S>
__global__ void track_angles_kernel(AngleArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    constexpr int O = kAngO;
    __shared__ double A[S][S], Q[S][S], H[O][S], R[O][O], Ppp[S][S];
    double P[S][S], K[S][O], m[S];
#pragma unroll
    for (int i = 0; i < S; ++i) {
        m[i] = a.mean[i];
#pragma unroll
        for (int j = 0; j < S; ++j) { A[i][j] = a.A[i * S + j]; Q[i][j] = a.Q[i * S + j]; P[i][j] = a.cov[i * S + j]; Ppp[i][j] = 0.0; }
#pragma unroll
        for (int o = 0; o < O; ++o) K[i][o] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < O; ++i) {
#pragma unroll
        for (int j = 0; j < S; ++j) H[i][j] = a.H[i * S + j];
#pragma unroll
        for (int j = 0; j < O; ++j) R[i][j] = a.R[i * O + j];
    }
    bool steady = false, prev_valid = false;
    for (int f = 0; f < a.n; ++f) {
        // KalmanTracker.sample(1): the last filtered state itself, read back as an angle (kalman.py:376, :236-242)
        double pred = atan2(m[0], m[a.order]);
        if (pred < 0) pred = 2 * 3.141592653589793 + pred;
        pred = pred * 57.29577951308232;
        double ang = a.angles[f];
        const double rel = angle_diff(pred, ang);
        if (a.scores[f] < 0.4) {
            ang = pred;
        } else if (fabs(rel) > 140.0) {
            ang = clamp_deg(ang + 180.0);
            a.flips[f] ^= 1;
        }
        a.angles[f] = ang;
        // filter_update: predict, then correct with (sin, cos) unless the angle is not finite
        const double rad = ang * 0.017453292519943295;
        const double z[2] = {sin(rad), cos(rad)};
        const bool valid = finite_f64(z[0]) && finite_f64(z[1]);
        double m2[S];
#pragma unroll
        for (int i = 0; i < S; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < S; ++k) acc = fma(A[i][k], m[k], acc);
            m2[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < S; ++i) m[i] = m2[i];
        double y[O];
#pragma unroll
        for (int o = 0; o < O; ++o) {
            double hm = 0.0;
#pragma unroll
            for (int k = 0; k < S; ++k) hm = fma(H[o][k], m[k], hm);
            y[o] = z[o] - hm;
        }
        if (steady && valid) {
#pragma unroll
            for (int i = 0; i < S; ++i) m[i] += K[i][0] * y[0] + K[i][1] * y[1];
            continue;
        }
        // ---- full covariance step ----
        double T1[S][S];
#pragma unroll
        for (int i = 0; i < S; ++i)
#pragma unroll
            for (int j = 0; j < S; ++j) {
                double t = 0.0;
#pragma unroll
                for (int k = 0; k < S; ++k) t = fma(A[i][k], P[k][j], t);
                T1[i][j] = t;
            }
#pragma unroll
        for (int i = 0; i < S; ++i)
#pragma unroll
            for (int j = 0; j < S; ++j) {
                double t = 0.0;
#pragma unroll
                for (int k = 0; k < S; ++k) t = fma(T1[i][k], A[j][k], t);
                P[i][j] = t + Q[i][j];
            }
        bool moved = false;
#pragma unroll
        for (int i = 0; i < S; ++i)
#pragma unroll
            for (int j = 0; j < S; ++j) {
                moved |= fabs(P[i][j] - Ppp[i][j]) > 1e-13 * sqrt(fabs(P[i][i] * P[j][j]));
                Ppp[i][j] = P[i][j];
            }
        steady = false;
        if (valid) {
            steady = prev_valid && !moved;
            double PHt[S][O], HP[O][S], Sm[O][O];
#pragma unroll
            for (int i = 0; i < S; ++i)
#pragma unroll
                for (int o = 0; o < O; ++o) {
                    double t = 0.0, u = 0.0;
#pragma unroll
                    for (int k = 0; k < S; ++k) { t = fma(P[i][k], H[o][k], t); u = fma(H[o][k], P[k][i], u); }
                    PHt[i][o] = t;
                    HP[o][i] = u;                              // H P (see kalman_filter_kernel: not (P H^T)^T)
                }
#pragma unroll
            for (int o = 0; o < O; ++o)
#pragma unroll
                for (int q = 0; q < O; ++q) {
                    double t = 0.0;
#pragma unroll
                    for (int k = 0; k < S; ++k) t = fma(H[o][k], PHt[k][q], t);
                    Sm[o][q] = t + R[o][q];
                }
            const double det = Sm[0][0] * Sm[1][1] - Sm[0][1] * Sm[1][0];
            const double Si[2][2] = {{Sm[1][1] / det, -Sm[0][1] / det}, {-Sm[1][0] / det, Sm[0][0] / det}};
#pragma unroll
            for (int i = 0; i < S; ++i)
#pragma unroll
                for (int o = 0; o < O; ++o) K[i][o] = PHt[i][0] * Si[0][o] + PHt[i][1] * Si[1][o];
#pragma unroll
            for (int i = 0; i < S; ++i) m[i] += K[i][0] * y[0] + K[i][1] * y[1];
#pragma unroll
            for (int i = 0; i < S; ++i)
#pragma unroll
                for (int j = 0; j < S; ++j) P[i][j] -= K[i][0] * HP[0][j] + K[i][1] * HP[1][j];
        }
        prev_valid = valid;
    }
#pragma unroll
    for (int i = 0; i < S; ++i) {
        a.mean[i] = m[i];
#pragma unroll
        for (int j = 0; j < S; ++j) a.cov[i * S + j] = P[i][j];
    }
}

// ---- shared-memory sizes ------------------------------------------------------------------------------------------------
size_t sparse_bytes_host(int rows, int cols) {
    return up16((size_t)rows * sizeof(int)) + up16((size_t)rows * cols * sizeof(int)) + (size_t)rows * cols * sizeof(double);
}
size_t filter_smem(int S, int O) {
    return (size_t)8 * (4 * S * S + 2 * S * O + 2 * O * O + O * O + 2 * S + 4 * O) + sparse_bytes_host(S, S) + sparse_bytes_host(O, S) + 64;
}
size_t gain_smem(int S) { return (size_t)8 * (3 * S * S + 2 * S) + sparse_bytes_host(S, S) + 64; }
size_t backward_smem(int S, bool cov) { return (size_t)8 * (S * S + 4 * S + (cov ? 3 * S * S : 0)) + 64; }
size_t em_smem(int S, int O) { return (size_t)8 * (3 * S * S + 3 * S + O) + sparse_bytes_host(S, S) + sparse_bytes_host(O, S) + 64; }

struct Workspace {
    double *mp, *mf, *ms, *Pp, *Pf, *J, *Ps, *pair, *sum_lo, *sum_hi, *sum_pair, *sum_valid;
    unsigned char *valid;
};
size_t carve_workspace(void *base, int T, int S, int O, bool em, Workspace *w) {
    size_t off = 0;
    auto take = [&](size_t bytes) { void *p = base ? (char *)base + off : nullptr; off += align_up(bytes, 256); return p; };
    const size_t TS = (size_t)T * S * 8, TSS = (size_t)T * S * S * 8, SS = (size_t)S * S * 8;
    Workspace tmp;
    tmp.mp = (double *)take(TS); tmp.mf = (double *)take(TS); tmp.ms = (double *)take(TS);
    tmp.Pp = (double *)take(TSS); tmp.Pf = (double *)take(TSS); tmp.J = (double *)take(TSS);
    tmp.valid = (unsigned char *)take((size_t)T);
    if (em) {
        tmp.Ps = (double *)take(TSS); tmp.pair = (double *)take(TSS);
        tmp.sum_lo = (double *)take(SS); tmp.sum_hi = (double *)take(SS); tmp.sum_pair = (double *)take(SS); tmp.sum_valid = (double *)take(SS);
    } else {
        tmp.Ps = tmp.pair = tmp.sum_lo = tmp.sum_hi = tmp.sum_pair = tmp.sum_valid = nullptr;
    }
    (void)O;
    if (w) *w = tmp;
    return off;
}

template <class K> int allow_smem(K kernel, size_t bytes, const char *name) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) { set_error("%s: %zu bytes of shared memory refused: %s", name, bytes, cudaGetErrorString(e)); return MSQ_ECUDA; }
    }
    return MSQ_OK;
}

// filter -> gains -> backward; leaves everything in the workspace
int run_smoother(const double *A, const double *H, const double *Q, const double *R, const double *m0, const double *P0,
                 const double *obs, int T, int S, int O, int predict_first, bool smooth, bool want_cov, const Workspace &w,
                 cudaStream_t st) {
    FilterArgs fa{A, H, Q, R, m0, P0, obs, T, S, O, predict_first, w.mp, w.mf, w.Pp, w.Pf, w.valid};
    int rc = allow_smem(kalman_filter_kernel, filter_smem(S, O), "kalman_filter");
    if (rc != MSQ_OK) return rc;
    {
        TimedLaunch timed(K_KALMAN, st);
        kalman_filter_kernel<<<1, kKalThreads, filter_smem(S, O), st>>>(fa);
        MSQ_LAUNCH_OK("kalman_filter");
    }
    if (!smooth) return MSQ_OK;
    if (T > 1) {
        rc = allow_smem(kalman_gain_kernel, gain_smem(S), "kalman_gain");
        if (rc != MSQ_OK) return rc;
        TimedLaunch timed(K_KALMAN, st);
        kalman_gain_kernel<<<T - 1, kKalThreads, gain_smem(S), st>>>(A, w.Pf, w.Pp, S, w.J);
        MSQ_LAUNCH_OK("kalman_gain");
    }
    BackwardArgs ba{w.mp, w.mf, w.Pp, w.Pf, w.J, T, S, want_cov ? 1 : 0, w.ms, w.Ps, w.pair};
    rc = allow_smem(kalman_backward_kernel, backward_smem(S, want_cov), "kalman_backward");
    if (rc != MSQ_OK) return rc;
    TimedLaunch timed(K_KALMAN, st);
    kalman_backward_kernel<<<1, kKalThreads, backward_smem(S, want_cov), st>>>(ba);
    MSQ_LAUNCH_OK("kalman_backward");
    return MSQ_OK;
}

}  // namespace
}  // namespace msq

using namespace msq;

extern "C" size_t msq_kalman_workspace_bytes(int T, int S, int O, int for_em) {
    if (T <= 0 || S <= 0 || O <= 0) return 0;
    return carve_workspace(nullptr, T, S, O, for_em != 0, nullptr);
}

static int check_dims(const char *fn, int T, int S, int O) {
    MSQ_REQUIRE(T >= 1, MSQ_EINVAL, "%s: need at least one time step (got %d)", fn, T);
    MSQ_REQUIRE(S >= 1 && S <= kMaxS && O >= 1 && O <= kMaxO && O <= S, MSQ_EUNSUPPORTED,
                "%s: supports 1 <= n_obs <= %d, n_obs <= n_state <= %d (got n_state=%d n_obs=%d)", fn, kMaxO, kMaxS, S, O);
    return MSQ_OK;
}

extern "C" int msq_kalman_smooth(const double *A, const double *H, const double *Q, const double *R, const double *m0,
                                 const double *P0, const double *obs, int T, int S, int O, int predict_first, int smooth,
                                 double *means_out, double *last_mean, double *last_cov, void *workspace,
                                 size_t workspace_bytes, void *stream) {
    int rc = check_dims("msq_kalman_smooth", T, S, O);
    if (rc != MSQ_OK) return rc;
    MSQ_REQUIRE(A && H && Q && R && m0 && P0 && obs, MSQ_EINVAL, "msq_kalman_smooth: null model or observation pointer");
    MSQ_REQUIRE(workspace && (uintptr_t)workspace % 256 == 0 && workspace_bytes >= msq_kalman_workspace_bytes(T, S, O, 0),
                MSQ_ENOMEM, "msq_kalman_smooth: workspace must be 256-byte aligned and >= %zu bytes",
                msq_kalman_workspace_bytes(T, S, O, 0));
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w;
    carve_workspace(workspace, T, S, O, false, &w);
    rc = run_smoother(A, H, Q, R, m0, P0, obs, T, S, O, predict_first, smooth != 0, false, w, st);
    if (rc != MSQ_OK) return rc;
    const double *means = smooth ? w.ms : w.mf;
    if (means_out) MSQ_CUDA_OK(cudaMemcpyAsync(means_out, means, (size_t)T * S * 8, cudaMemcpyDeviceToDevice, st));
    // the smoother's last state is the filter's last state (kalman.py:399-400 keeps it as the next chunk's prior)
    if (last_mean) MSQ_CUDA_OK(cudaMemcpyAsync(last_mean, w.mf + (size_t)(T - 1) * S, (size_t)S * 8, cudaMemcpyDeviceToDevice, st));
    if (last_cov) MSQ_CUDA_OK(cudaMemcpyAsync(last_cov, w.Pf + (size_t)(T - 1) * S * S, (size_t)S * S * 8, cudaMemcpyDeviceToDevice, st));
    return MSQ_OK;
}

extern "C" int msq_kalman_em(const double *A, const double *H, double *Q, double *R, const double *m0, double *P0,
                             const double *obs, int T, int S, int O, int n_iter, void *workspace, size_t workspace_bytes,
                             void *stream) {
    int rc = check_dims("msq_kalman_em", T, S, O);
    if (rc != MSQ_OK) return rc;
    MSQ_REQUIRE(T >= 2, MSQ_EINVAL, "msq_kalman_em: needs at least two time steps (got %d)", T);
    MSQ_REQUIRE(A && H && Q && R && m0 && P0 && obs, MSQ_EINVAL, "msq_kalman_em: null model or observation pointer");
    MSQ_REQUIRE(workspace && (uintptr_t)workspace % 256 == 0 && workspace_bytes >= msq_kalman_workspace_bytes(T, S, O, 1),
                MSQ_ENOMEM, "msq_kalman_em: workspace must be 256-byte aligned and >= %zu bytes",
                msq_kalman_workspace_bytes(T, S, O, 1));
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w;
    carve_workspace(workspace, T, S, O, true, &w);
    rc = allow_smem(kalman_em_finish_kernel, em_smem(S, O), "kalman_em_finish");
    if (rc != MSQ_OK) return rc;
    for (int it = 0; it < n_iter; ++it) {
        rc = run_smoother(A, H, Q, R, m0, P0, obs, T, S, O, 0, true, true, w, st);
        if (rc != MSQ_OK) return rc;
        TimedLaunch timed(K_KALMAN, st);
        kalman_em_reduce_kernel<<<(S * S + 127) / 128, 128, 0, st>>>(w.Ps, w.pair, w.valid, T, S * S, w.sum_lo, w.sum_hi, w.sum_pair, w.sum_valid);
        MSQ_LAUNCH_OK("kalman_em_reduce");
        EmArgs ea{A, H, m0, obs, w.ms, w.Ps, w.valid, w.sum_lo, w.sum_hi, w.sum_pair, w.sum_valid, T, S, O, Q, R, P0};
        kalman_em_finish_kernel<<<1, kKalThreads, em_smem(S, O), st>>>(ea);
        MSQ_LAUNCH_OK("kalman_em_finish");
    }
    return MSQ_OK;
}

extern "C" int msq_keypoint_alignment_scores(const double *kpts, int kp_stride, const double *centroid, const double *angles,
                                             int n, double *scores, void *stream) {
    MSQ_REQUIRE(n >= 0 && (kp_stride == 2 || kp_stride == 3), MSQ_EINVAL, "msq_keypoint_alignment_scores: bad sizes");
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(kpts && centroid && angles && scores, MSQ_EINVAL, "msq_keypoint_alignment_scores: null pointer");
    TimedLaunch timed(K_KALMAN, (cudaStream_t)stream);
    keypoint_alignment_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(kpts, kp_stride, centroid, angles, n, scores);
    MSQ_LAUNCH_OK("keypoint_alignment");
    return MSQ_OK;
}

extern "C" int msq_tracking_prepare(const double *orientation, const double *axis, const double *centroid, const double *kpts,
                                    int n, double *angles, uint8_t *flips, double *conf, double *scores, void *stream) {
    MSQ_REQUIRE(n >= 0, MSQ_EINVAL, "msq_tracking_prepare: bad n");
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(orientation && axis && centroid && kpts && angles && flips && scores, MSQ_EINVAL, "msq_tracking_prepare: null pointer");
    TimedLaunch timed(K_KALMAN, (cudaStream_t)stream);
    tracking_prepare_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(orientation, axis, centroid, kpts, n, angles, flips, conf, scores);
    MSQ_LAUNCH_OK("tracking_prepare");
    return MSQ_OK;
}

extern "C" int msq_track_angles(const double *A, const double *H, const double *Q, const double *R, double *mean, double *cov,
                                int S, double *angles, uint8_t *flips, const double *scores, int n, void *stream) {
    MSQ_REQUIRE(n >= 0, MSQ_EINVAL, "msq_track_angles: bad n");
    MSQ_REQUIRE(S >= 2 && S <= kAngS && S % 2 == 0, MSQ_EUNSUPPORTED,
                "msq_track_angles: the angle filter tracks (sin, cos) at order <= %d (got n_state=%d)", kAngS / 2, S);
    if (n == 0) return MSQ_OK;
    MSQ_REQUIRE(A && H && Q && R && mean && cov && angles && flips && scores, MSQ_EINVAL, "msq_track_angles: null pointer");
    AngleArgs aa{A, H, Q, R, mean, cov, angles, flips, scores, n, S, S / 2};
    TimedLaunch timed(K_KALMAN, (cudaStream_t)stream);
    switch (S) {
        case 2: track_angles_kernel<2><<<1, 32, 0, (cudaStream_t)stream>>>(aa); break;
        case 4: track_angles_kernel<4><<<1, 32, 0, (cudaStream_t)stream>>>(aa); break;
        case 6: track_angles_kernel<6><<<1, 32, 0, (cudaStream_t)stream>>>(aa); break;
        default: track_angles_kernel<8><<<1, 32, 0, (cudaStream_t)stream>>>(aa); break;
    }
    MSQ_LAUNCH_OK("track_angles");
    return MSQ_OK;
}
