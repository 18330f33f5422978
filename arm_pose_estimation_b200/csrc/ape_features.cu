// Stage 1 kernel: raw IMU rows -> calibrated, z-scored feature rows in the per-stream feature ring.
// HBM-bound byte shuffling + ~250 float64 flops per row: rows are staged through shared memory with
// coalesced 4-byte loads (a row is 112 or 220 B, so per-thread row reads would touch 2-3 lines each),
// then one thread per row does the quaternion calibration in float64.
#include "ape_features.cuh"

namespace ape {

constexpr int FEAT_ROWS_PER_CTA = 32;    // latency-bound fp64 row math: small CTAs spread a frame of ~1k rows over more SMs

struct SmemRow {                       // indexable view of one staged row (odd stride -> no bank conflicts)
    const float* p;
    __device__ float operator[](int i) const { return p[i]; }
};

__global__ void __launch_bounds__(FEAT_ROWS_PER_CTA)
features_kernel(const float* __restrict__ raw, int layout, int kind, const double* __restrict__ xx_m,
                const double* __restrict__ xx_s, int normalize, float* __restrict__ feats,
                int total_rows, int nF, int frame0, const int32_t* __restrict__ stream_frames, int feat_ring) {
    extern __shared__ float s_rows[];
    const int ncols = row_layout(layout).ncols;
    const int stride = ncols | 1;                                  // 29 / 55 words: odd
    const int row0 = blockIdx.x * FEAT_ROWS_PER_CTA;
    const int rows_here = min(FEAT_ROWS_PER_CTA, total_rows - row0);

    const float* src = raw + (size_t)row0 * ncols;
    for (int i = threadIdx.x; i < rows_here * ncols; i += FEAT_ROWS_PER_CTA)
        s_rows[(i / ncols) * stride + (i % ncols)] = __ldg(src + i);
    __syncthreads();
    if ((int)threadIdx.x >= rows_here) return;

    double xx[38];
    const int I = compute_features(kind, layout, SmemRow{s_rows + threadIdx.x * stride}, xx);
    const int row = row0 + threadIdx.x;
    const int b = row / nF;
    const int fb = stream_frames ? stream_frames[b] : frame0;       // per-stream frame counter; < 0: the stream sits this call out
    if (fb < 0) return;
    const int f = fb + row % nF;
    float* dst = feats + ((size_t)b * feat_ring + (f % feat_ring)) * I;
    for (int j = 0; j < I; ++j) {
        double v = normalize ? (xx[j] - xx_m[j]) / xx_s[j] : xx[j];   // estimator.py:103-104
        dst[j] = (float)v;                                            // torch.tensor(..., float32), watch_only.py:86
    }
}

// The second of the reference's three per-frame calls handed an already-parsed feature row (estimator.py:93-104): z-score it and
// put it into the stream's feature ring at the stream's frame, exactly where features_kernel would have put it.
__global__ void features_push_kernel(const double* __restrict__ xx, int I, const double* __restrict__ xx_m,
                                     const double* __restrict__ xx_s, int normalize, float* __restrict__ feats, int B,
                                     int frame0, const int32_t* __restrict__ stream_frames, int feat_ring) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * I) return;
    const int b = idx / I, j = idx - b * I;
    const int f = stream_frames ? stream_frames[b] : frame0;
    if (f < 0) return;
    const double v = normalize ? (xx[idx] - xx_m[j]) / xx_s[j] : xx[idx];
    feats[((size_t)b * feat_ring + (f % feat_ring)) * I + j] = (float)v;
}

}  // namespace ape

extern "C" int ape_features_push(const double* xx, int I, const double* xx_m, const double* xx_s, int normalize, float* feats,
                                 int B, int frame0, const int32_t* stream_frames, int feat_ring, void* stream) {
    using namespace ape;
    if (!xx || !feats || I < 1 || B < 0 || frame0 < 0 || feat_ring < 1) return APE_ERR_BAD_ARG;
    if (normalize && (!xx_m || !xx_s)) return APE_ERR_BAD_ARG;
    if (B == 0) return APE_OK;
    if ((long long)B * I > 0x7fffffffLL) return APE_ERR_BAD_ARG;
    const int threads = 128, grid = (B * I + threads - 1) / threads;
    features_push_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(xx, I, xx_m, xx_s, normalize, feats, B, frame0, stream_frames, feat_ring);
    return check_launch();
}

extern "C" int ape_features(const float* raw, int layout, int kind, const double* xx_m, const double* xx_s,
                            int normalize, float* feats, int B, int nF, int frame0, const int32_t* stream_frames, int feat_ring,
                            void* stream) {
    using namespace ape;
    if (!raw || !feats || B < 0 || nF < 0 || frame0 < 0 || feat_ring < 1) return APE_ERR_BAD_ARG;
    if (layout != APE_LAYOUT_WATCH_ONLY && layout != APE_LAYOUT_WATCH_PHONE) return APE_ERR_BAD_ARG;
    if (kind < APE_KIND_WATCH_ONLY || kind > APE_KIND_UARM) return APE_ERR_BAD_ARG;
    if (kind != APE_KIND_WATCH_ONLY && layout != APE_LAYOUT_WATCH_PHONE) return APE_ERR_BAD_ARG;   // phone columns needed
    if (normalize && (!xx_m || !xx_s)) return APE_ERR_BAD_ARG;
    if (nF > feat_ring) return APE_ERR_BAD_ARG;                    // a launch may not lap its own ring
    const long long total = (long long)B * nF;
    if (total == 0) return APE_OK;
    if (total > 0x7fffffffLL) return APE_ERR_BAD_ARG;
    const int stride = row_layout(layout).ncols | 1;
    const size_t smem = (size_t)FEAT_ROWS_PER_CTA * stride * sizeof(float);
    const int grid = (int)((total + FEAT_ROWS_PER_CTA - 1) / FEAT_ROWS_PER_CTA);
    features_kernel<<<grid, FEAT_ROWS_PER_CTA, smem, (cudaStream_t)stream>>>(
        raw, layout, kind, xx_m, xx_s, normalize, feats, (int)total, nF, frame0, stream_frames, feat_ring);
    return check_launch();
}
