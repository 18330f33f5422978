// MC-dropout feed-forward regressors (DropoutFF / DropoutFF2D, nn_models.py:252-370 of the reference): Linear + leaky_relu
// stack, dropout ONLY in front of the output layer.  The hidden stack therefore does not depend on the MC sample: it is
// evaluated once per input row, and only  y_s = W_o (mask_s * h) / (1 - p) + b_o  runs per sample.  fp32 FFMA throughout.
// One CTA per input row: threads over output units for the dense layers (weights pre-transposed, coalesced), then one warp
// per MC sample (lanes over hidden units, warp-shuffle reduction per output).  Not on the deployed models' path; small.
#include "ape_common.cuh"

namespace ape {

constexpr int FF_THREADS = 256;

struct FfArgs {
    const float* blob;       // layer 0: W^T [I][H], b [H]; hidden l: W^T [H][H], b [H]; output: W [O][H], b [O]
    int I, H, Lh, O;
    const float* x;          // [rows][I]
    int rows, n;
    int mask_mode;
    const uint8_t* masks;    // APE_MASK_INJECTED: [rows][n][H] of {0,1}
    uint64_t seed;
    uint32_t stream_id0, frame0, keep_thr16;
    float keep_scale;
    float* preds;            // [rows][n][O]
};

__global__ void __launch_bounds__(FF_THREADS) mc_ff_kernel(FfArgs a) {
    extern __shared__ float sh[];                      // [2][max(I, H)]
    const int K0 = a.I > a.H ? a.I : a.H;
    float* cur = sh;
    float* nxt = sh + K0;
    const int row = blockIdx.x, tid = threadIdx.x;
    for (int k = tid; k < a.I; k += FF_THREADS) cur[k] = a.x[(size_t)row * a.I + k];
    __syncthreads();
    const float* w = a.blob;
    int K = a.I;
    for (int l = 0; l <= a.Lh; ++l) {                  // input layer + Lh hidden layers, leaky_relu(0.01) after each
        const float* bias = w + (size_t)K * a.H;
        for (int j = tid; j < a.H; j += FF_THREADS) {
            float acc = bias[j];
            for (int k = 0; k < K; ++k) acc = fmaf(__ldg(w + (size_t)k * a.H + j), cur[k], acc);
            nxt[j] = acc > 0.0f ? acc : 0.01f * acc;
        }
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
        w = bias + a.H;
        K = a.H;
    }
    const float* wo = w;
    const float* bo = wo + (size_t)a.O * a.H;
    const int warp = tid >> 5, lane = tid & 31;
    for (int s = warp; s < a.n; s += FF_THREADS / 32) {
        float acc[20];
#pragma unroll
        for (int o = 0; o < 20; ++o) acc[o] = 0.0f;
        for (int k0 = 0; k0 < a.H; k0 += 32) {
            const int k = k0 + lane;
            float v = 0.0f;
            if (k < a.H) {
                bool keep = true;
                if (a.mask_mode == APE_MASK_INJECTED) {
                    keep = a.masks[((size_t)row * a.n + s) * a.H + k] != 0;
                } else if (a.mask_mode == APE_MASK_PHILOX) {       // gap id 15 is reserved for the feed-forward output dropout
                    const uint32_t bits = philox_keep8(a.seed, a.stream_id0 + (uint32_t)row, a.frame0, (uint32_t)s, 15u, 0u,
                                                       (uint32_t)(k >> 3), a.keep_thr16);
                    keep = (bits >> (k & 7)) & 1u;
                }
                v = keep ? cur[k] * a.keep_scale : 0.0f;
            }
#pragma unroll
            for (int o = 0; o < 20; ++o)
                if (o < a.O && k < a.H) acc[o] = fmaf(__ldg(wo + (size_t)o * a.H + k), v, acc[o]);
        }
#pragma unroll
        for (int o = 0; o < 20; ++o) {
            if (o < a.O) {
                float v = acc[o];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                if (lane == 0) a.preds[((size_t)row * a.n + s) * a.O + o] = v + __ldg(bo + o);
            }
        }
    }
}

// One dense layer  y = act(x W^T + b)  over many rows: the input layer of ImuPoseLSTM (nn_models.py:210-249: Linear(I, 256) + relu in
// front of a plain 2-layer LSTM).  W^T [K][H] pre-transposed like the feed-forward blob; a CTA stages DENSE_ROWS input rows in
// shared memory and its threads sweep the output units (coalesced weight reads, DENSE_ROWS accumulators per thread).
constexpr int DENSE_ROWS = 8;

__global__ void __launch_bounds__(FF_THREADS) dense_act_kernel(const float* __restrict__ wt, const float* __restrict__ bias,
                                                               const float* __restrict__ x, float* __restrict__ y,
                                                               int rows, int K, int H, int act) {
    extern __shared__ float sx[];                      // [DENSE_ROWS][K]
    const int row0 = blockIdx.x * DENSE_ROWS, tid = threadIdx.x;
    const int nr = min(DENSE_ROWS, rows - row0);
    for (int i = tid; i < DENSE_ROWS * K; i += FF_THREADS) sx[i] = i < nr * K ? x[(size_t)row0 * K + i] : 0.0f;
    __syncthreads();
    for (int j = tid; j < H; j += FF_THREADS) {
        float acc[DENSE_ROWS];
        const float b = bias[j];
#pragma unroll
        for (int r = 0; r < DENSE_ROWS; ++r) acc[r] = b;
        for (int k = 0; k < K; ++k) {
            const float w = __ldg(wt + (size_t)k * H + j);
#pragma unroll
            for (int r = 0; r < DENSE_ROWS; ++r) acc[r] = fmaf(w, sx[r * K + k], acc[r]);
        }
#pragma unroll
        for (int r = 0; r < DENSE_ROWS; ++r) {
            if (r < nr) {
                float v = acc[r];
                if (act == 1) v = v > 0.0f ? v : 0.0f;               // relu
                else if (act == 2) v = v > 0.0f ? v : 0.01f * v;     // leaky_relu(0.01)
                y[(size_t)(row0 + r) * H + j] = v;
            }
        }
    }
}

}  // namespace ape

extern "C" int ape_dense_act(const float* wt, const float* bias, const float* x, float* y, int rows, int K, int H, int act,
                             void* stream) {
    using namespace ape;
    if (!wt || !bias || !x || !y || rows < 0 || K < 1 || H < 1 || act < 0 || act > 2) return APE_ERR_BAD_ARG;
    if (rows == 0) return APE_OK;
    const size_t smem = sizeof(float) * (size_t)DENSE_ROWS * K;
    if (smem > 200 * 1024) return APE_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        APE_CUDA_TRY(cudaFuncSetAttribute(dense_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense_act_kernel<<<(rows + DENSE_ROWS - 1) / DENSE_ROWS, FF_THREADS, smem, (cudaStream_t)stream>>>(wt, bias, x, y, rows, K, H, act);
    return check_launch();
}

extern "C" int ape_ff_blob_floats(int I, int H, int Lh, int O, int64_t* floats) {
    if (!floats || I < 1 || H < 1 || Lh < 0 || O < 1) return APE_ERR_BAD_ARG;
    *floats = (int64_t)I * H + H + (int64_t)Lh * ((int64_t)H * H + H) + (int64_t)O * H + O;
    return APE_OK;
}

extern "C" int ape_mc_ff(const float* blob, int I, int H, int Lh, int O, float dropout_p, const float* x, int rows,
                         int n_samples, int mask_mode, const uint8_t* masks, uint64_t philox_seed, uint32_t stream_id0,
                         uint32_t frame0, float* preds, void* stream) {
    using namespace ape;
    if (!blob || !x || !preds || I < 1 || H < 1 || Lh < 0 || O < 1 || O > 20 || rows < 0 || n_samples < 1) return APE_ERR_BAD_ARG;
    if (mask_mode < APE_MASK_NONE || mask_mode > APE_MASK_PHILOX) return APE_ERR_BAD_ARG;
    if (mask_mode == APE_MASK_INJECTED && !masks) return APE_ERR_BAD_ARG;
    if (mask_mode != APE_MASK_NONE && !(dropout_p >= 0.0f && dropout_p < 1.0f)) return APE_ERR_BAD_ARG;
    if (rows == 0) return APE_OK;
    const size_t smem = 2 * sizeof(float) * (size_t)(I > H ? I : H);
    if (smem > 200 * 1024) return APE_ERR_UNSUPPORTED;
    FfArgs a{blob, I, H, Lh, O, x, rows, n_samples, mask_mode, masks, philox_seed, stream_id0, frame0,
             keep_threshold16(dropout_p), mask_mode == APE_MASK_NONE ? 1.0f : 1.0f / (1.0f - dropout_p), preds};
    if (smem > 48 * 1024)
        APE_CUDA_TRY(cudaFuncSetAttribute(mc_ff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mc_ff_kernel<<<rows, FF_THREADS, smem, (cudaStream_t)stream>>>(a);
    return check_launch();
}
