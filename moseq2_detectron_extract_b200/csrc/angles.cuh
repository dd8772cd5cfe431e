// Angle / keypoint-vote device helpers shared by epilogue.cu (a8-a10) and kalman.cu (a14).
#pragma once
#include "common.cuh"
#include <math.h>

namespace msq {
namespace {

constexpr double kPiOver180 = 0.017453292519943295;     // np.pi / 180
constexpr double k180OverPi = 57.29577951308232;        // 180 / np.pi

__device__ __forceinline__ double nan_f64() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ double np_max2(double a, double b) { return (a != a || b != b) ? nan_f64() : (a > b ? a : b); }
__device__ __forceinline__ double np_min2(double a, double b) { return (a != a || b != b) ? nan_f64() : (a < b ? a : b); }

// rotate_points (ref proc/keypoints.py:11-39): R(-angle) @ (p - o) + o
__device__ __forceinline__ void rotate_about(double px, double py, double ox, double oy, double c, double s,
                                             double &rx, double &ry) {
    const double dx = px - ox, dy = py - oy;
    rx = (c * dx + (-s) * dy) + ox;
    ry = (s * dx + c * dy) + oy;
}

// ---------------------------------------------------------------------------------------------
// flips_from_keypoints for one frame (ref proc/proc.py:851-889): front (0..3) and rear (4..6)
// keypoints, rotated by -angle about the centroid, vote for the nearer end of the body box
// ---------------------------------------------------------------------------------------------
template <class KP>
__device__ __forceinline__ bool keypoint_flip_vote(const KP *__restrict__ kp, double cx, double cy, double angle,
                                                   double length, double *conf) {
    const double t = (-angle) * kPiOver180;
    const double c = cos(t), s = sin(t);
    const double lo = cx - length / 2, hi = cx + length / 2;
    int votes[MSQ_NUM_KEYPOINTS - 1];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        double rx, ry;
        rotate_about((double)kp[k * 3], (double)kp[k * 3 + 1], cx, cy, c, s, rx, ry);
        votes[k] = (fabs(lo - rx) < fabs(hi - rx)) ? -1 : 1;
    }
    const int front = votes[0] + votes[1] + votes[2] + votes[3];
    const int rear = votes[4] + votes[5] + votes[6];
    const bool flip = 3 * front < 4 * rear;              // mean(front) < mean(rear)
    const int want_front = flip ? -1 : 1;
    int agree = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) agree += votes[k] == want_front;
#pragma unroll
    for (int k = 4; k < 7; ++k) agree += votes[k] == -want_front;
    if (conf) *conf = (double)agree / 7.0;
    return flip;
}

}  // namespace
}  // namespace msq
