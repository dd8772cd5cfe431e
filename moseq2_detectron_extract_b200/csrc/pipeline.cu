// Whole-chunk pipeline: what ProcessFeaturesStep.process does per chunk with use_tracking=False and
// <=1 instance per frame (ref pipeline/process_features_step.py:56-60, 163-199; proc/proc.py:700-848):
//   clean -> moment features -> degrees/flips/angle filter -> scalars + keypoint table -> crops.
// Pure sequencing: the caller's stream plus one internal side stream forked and joined with events (no host
// synchronisation, no device allocation).
#include "common.cuh"

using namespace msq;

extern "C" size_t msq_extract_scratch_bytes(int n, int /*h*/, int /*w*/) {
    const size_t nn = (size_t)(n > 0 ? n : 0);
    return align_up(nn * sizeof(double), 256) + align_up(nn * sizeof(int2), 256) + align_up(msq_crop_scratch_bytes((int)nn), 256) +
           align_up((nn + 1) * sizeof(int), 256) + align_up(clean_scratch_bytes((int)nn, 4096), 256);
}

// The one piece of state the whole-chunk entry point needs: a side stream + two events on one device, so that the masked
// sums run beside the short latency-bound launches (left-over feature frames, angle filter) of the main stream.  Explicit object (create /
// destroy) for callers that manage their own resources; msq_extract_chunk() keeps one per (host thread, device) for callers
// that do not.
struct msq_engine {
    int device;
    cudaStream_t side;
    cudaEvent_t fork, join, angles_done;
};

extern "C" int msq_engine_create(msq_engine **engine) {
    MSQ_REQUIRE(engine, MSQ_EINVAL, "msq_engine_create: null pointer");
    msq_engine *e = new msq_engine();
    if (cudaGetDevice(&e->device) != cudaSuccess || cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->angles_done, cudaEventDisableTiming) != cudaSuccess) {
        set_error("msq_engine_create: %s", cudaGetErrorString(cudaGetLastError()));
        delete e;
        return MSQ_ECUDA;
    }
    *engine = e;
    return MSQ_OK;
}

extern "C" int msq_engine_destroy(msq_engine *e) {
    if (!e) return MSQ_OK;
    cudaEventDestroy(e->fork);
    cudaEventDestroy(e->join);
    cudaEventDestroy(e->angles_done);
    cudaStreamDestroy(e->side);
    delete e;
    return MSQ_OK;
}

extern "C" int msq_extract_chunk_engine(msq_engine *engine, const uint8_t *chunk_dev, const uint32_t *positive_bits, const uint8_t *mask_dev,
                                        const float *kpts_dev, int n, int h, int w, int chunk, double min_height, double max_height, double true_depth,
                                        int crop_w, int crop_h, const msq_chunk_outputs *out, void *scratch, size_t scratch_bytes,
                                        void *stream) {
    MSQ_REQUIRE(engine, MSQ_EINVAL, "msq_extract_chunk_engine: null engine");
    MSQ_REQUIRE(chunk_dev && mask_dev && kpts_dev && out, MSQ_EINVAL, "msq_extract_chunk: null input pointer");
    MSQ_REQUIRE((uintptr_t)positive_bits % 4 == 0, MSQ_EINVAL, "msq_extract_chunk: positive_bits must be 4-byte aligned");
    MSQ_REQUIRE(out->cleaned && out->centroid && out->angle_deg && out->axis_length && out->flips && out->scalars &&
                    out->kpt_cols && out->depth_crops && out->mask_crops,
                MSQ_EINVAL, "msq_extract_chunk: null output pointer");
    MSQ_REQUIRE(n >= 0 && h > 0 && w > 0 && chunk > 0 && crop_w > 0 && crop_h > 0, MSQ_EINVAL,
                "msq_extract_chunk: bad sizes n=%d h=%d w=%d chunk=%d crop=%dx%d", n, h, w, chunk, crop_w, crop_h);
    if (n == 0) return MSQ_OK;
    int dev = -1;
    MSQ_CUDA_OK(cudaGetDevice(&dev));
    MSQ_REQUIRE(dev == engine->device, MSQ_EINVAL, "msq_extract_chunk: engine belongs to device %d, current device is %d", engine->device, dev);
    MSQ_REQUIRE(scratch && (uintptr_t)scratch % 256 == 0 && scratch_bytes >= msq_extract_scratch_bytes(n, h, w), MSQ_ENOMEM,
                "msq_extract_chunk: scratch must be 256-byte aligned and >= %zu bytes", msq_extract_scratch_bytes(n, h, w));
    cudaStream_t st = (cudaStream_t)stream;
    double *orientation = reinterpret_cast<double *>(scratch);
    char *base = reinterpret_cast<char *>(scratch);
    int2 *sums = reinterpret_cast<int2 *>(base + align_up((size_t)n * sizeof(double), 256));
    void *crop_scratch = base + align_up((size_t)n * sizeof(double), 256) + align_up((size_t)n * sizeof(int2), 256);
    int *feature_list = reinterpret_cast<int *>(reinterpret_cast<char *>(crop_scratch) + align_up(msq_crop_scratch_bytes(n), 256));
    int2 *clean_bands = reinterpret_cast<int2 *>(reinterpret_cast<char *>(feature_list) + align_up((size_t)(n + 1) * sizeof(int), 256));
    MSQ_REQUIRE(w <= 4096, MSQ_EUNSUPPORTED, "msq_extract_chunk: frames wider than 4096 pixels (got %d)", w);
    int rc;
    RowBands rows = {nullptr, 0};          // rows of every cleaned frame that can be non-zero: the feature pass reads only those
    if ((rc = launch_clean(chunk_dev, out->cleaned, n, h, w, st, clean_bands, &rows, positive_bits)) != MSQ_OK) return rc;
    // frame_threshold = 3 (ref proc/proc.py:716).  After the streaming feature pass the main stream only has short, latency-bound
    // work for a while (the general feature kernel on the few frames the fast path left over, then the per-chunk angle filter: a
    // handful of CTAs each); the bandwidth-bound masked sums, which need none of it, run beside them on the side stream.
    if ((rc = launch_frame_features(out->cleaned, mask_dev, n, h, w, 3.0, out->centroid, orientation, out->axis_length,
                                    nullptr, feature_list, st, engine->fork, rows)) != MSQ_OK) return rc;
    MSQ_CUDA_OK(cudaStreamWaitEvent(engine->side, engine->fork, 0));
    if ((rc = launch_masked_sums(chunk_dev, mask_dev, n, h, w, min_height, max_height, sums, engine->side)) != MSQ_OK) return rc;
    if ((rc = launch_angles_and_flips(orientation, out->axis_length, out->centroid, kpts_dev, n, chunk, out->angle_deg,
                                      out->flips, nullptr, out->filter_passes, st)) != MSQ_OK) return rc;
    // the scalar / keypoint table (needs the sums and the angles, feeds nothing inside this call) follows the sums on the side
    // stream and runs beside the crops
    MSQ_CUDA_OK(cudaEventRecord(engine->angles_done, st));
    MSQ_CUDA_OK(cudaStreamWaitEvent(engine->side, engine->angles_done, 0));
    if ((rc = launch_scalars_and_keypoints(chunk_dev, mask_dev, out->cleaned, out->centroid, out->angle_deg,
                                           out->axis_length, kpts_dev, false, n, h, w, chunk, min_height, max_height,
                                           true_depth, out->scalars, out->kpt_cols, sums, engine->side, true)) != MSQ_OK) return rc;
    MSQ_CUDA_OK(cudaEventRecord(engine->join, engine->side));
    rc = launch_crop_rotate(chunk_dev, mask_dev, n, h, w, out->centroid, out->angle_deg, crop_w, crop_h,
                            out->depth_crops, out->mask_crops, crop_scratch, st);
    MSQ_CUDA_OK(cudaStreamWaitEvent(st, engine->join, 0));          // everything of this call is ordered before what follows on `st`
    return rc;
}

extern "C" int msq_extract_chunk(const uint8_t *chunk_dev, const uint8_t *mask_dev, const float *kpts_dev, int n,
                                 int h, int w, int chunk, double min_height, double max_height, double true_depth,
                                 int crop_w, int crop_h, const msq_chunk_outputs *out, void *scratch,
                                 size_t scratch_bytes, void *stream) {
    // one lazily created engine per (host thread, device): a thread that moves to another GPU gets a fresh one there
    constexpr int kMaxDevices = 64;
    static thread_local msq_engine *engines[kMaxDevices] = {nullptr};
    int dev = 0;
    MSQ_CUDA_OK(cudaGetDevice(&dev));
    MSQ_REQUIRE(dev >= 0 && dev < kMaxDevices, MSQ_EUNSUPPORTED, "msq_extract_chunk: device index %d", dev);
    if (!engines[dev]) {
        const int rc = msq_engine_create(&engines[dev]);
        if (rc != MSQ_OK) return rc;
    }
    return msq_extract_chunk_engine(engines[dev], chunk_dev, nullptr, mask_dev, kpts_dev, n, h, w, chunk, min_height, max_height, true_depth,
                                    crop_w, crop_h, out, scratch, scratch_bytes, stream);
}
