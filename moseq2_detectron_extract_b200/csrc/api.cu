// Error plumbing, version and device queries of the C ABI.
#include "common.cuh"
#include <atomic>
#include <mutex>
#include <algorithm>
#include <string.h>

namespace msq {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached = v;
        cached_dev = dev;
    }
    return cached;
}

// ------------------------------------------------------------------------------------------------
// kernel timing: a fixed pool of event pairs, filled round-robin while enabled
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int kTimerSlots = 4096;
struct TimerState {
    std::atomic<bool> enabled{false};
    std::atomic<int> used{0};            // steps of the pipeline are threads of one process: slots are claimed atomically
    std::atomic<int> dropped{0};
    std::atomic<long long> launches[K_COUNT];
    std::mutex collect_lock;             // enable / collect against each other
    cudaEvent_t start[kTimerSlots], stop[kTimerSlots];
    int kernel[kTimerSlots];
    bool created = false;
} g_timer;
const char *const kKernelNames[K_COUNT] = {"prep_frames", "scale_frames", "clean_frames", "frame_features",
                                           "angles_flips_filter", "masked_sums", "scalars_keypoints", "crop_rotate",
                                           "paste_masks", "inpaint", "kalman_tracking", "bground_median", "session_roi", "detector_glue"};
}  // namespace

TimedLaunch::TimedLaunch(int kernel_id, cudaStream_t stream, int kernels) : slot(-1), st(stream) {
    g_timer.launches[kernel_id].fetch_add(kernels, std::memory_order_relaxed);
    if (!g_timer.enabled.load(std::memory_order_acquire)) return;
    const int claimed = g_timer.used.fetch_add(1, std::memory_order_acq_rel);
    if (claimed >= kTimerSlots) { g_timer.dropped.fetch_add(1, std::memory_order_relaxed); return; }
    slot = claimed;
    g_timer.kernel[slot] = kernel_id;
    cudaEventRecord(g_timer.start[slot], st);
}
TimedLaunch::~TimedLaunch() {
    if (slot >= 0) cudaEventRecord(g_timer.stop[slot], st);
}

}  // namespace msq

extern "C" int msq_kernel_timing_enable(int enable) {
    using namespace msq;
    std::lock_guard<std::mutex> guard(g_timer.collect_lock);
    if (enable && !g_timer.created) {
        for (int i = 0; i < kTimerSlots; ++i) {
            MSQ_CUDA_OK(cudaEventCreate(&g_timer.start[i]));
            MSQ_CUDA_OK(cudaEventCreate(&g_timer.stop[i]));
        }
        g_timer.created = true;
    }
    g_timer.enabled.store(enable != 0, std::memory_order_release);
    return MSQ_OK;
}

// Synchronises on the recorded events, ADDS their durations into total_ms[K]/timed[K], clears the pool.
extern "C" int msq_kernel_timing_collect(double *total_ms, long long *timed, int capacity) {
    using namespace msq;
    MSQ_REQUIRE(total_ms && timed && capacity >= K_COUNT, MSQ_EINVAL, "msq_kernel_timing_collect: need %d slots", (int)K_COUNT);
    // call with timing disabled (or from the only launching thread): a launch in flight on another thread may still be recording
    std::lock_guard<std::mutex> guard(g_timer.collect_lock);
    const int used = std::min(g_timer.used.load(std::memory_order_acquire), kTimerSlots);
    for (int i = 0; i < used; ++i) {
        MSQ_CUDA_OK(cudaEventSynchronize(g_timer.stop[i]));
        float ms = 0.f;
        MSQ_CUDA_OK(cudaEventElapsedTime(&ms, g_timer.start[i], g_timer.stop[i]));
        total_ms[g_timer.kernel[i]] += ms;
        timed[g_timer.kernel[i]] += 1;
    }
    g_timer.used.store(0, std::memory_order_release);
    return MSQ_OK;
}

extern "C" int msq_kernel_count(void) { return msq::K_COUNT; }
extern "C" const char *msq_kernel_name(int id) { return (id >= 0 && id < msq::K_COUNT) ? msq::kKernelNames[id] : nullptr; }
// number of kernel launches issued by this library since load (per kernel id), for `gpu_launches`
extern "C" long long msq_kernel_launches(int id) { return (id >= 0 && id < msq::K_COUNT) ? msq::g_timer.launches[id].load() : -1; }

extern "C" int msq_version(void) { return MSQ_VERSION; }

extern "C" const char *msq_last_error(void) { return msq::g_error; }

extern "C" int msq_device_info(int *sms, int *cc_major, int *cc_minor, char *name, int name_len) {
    int dev = 0;
    MSQ_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MSQ_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    if (sms) *sms = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (name && name_len > 0) {
        strncpy(name, prop.name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    return MSQ_OK;
}
