// Measured fp32 FMA peak of the current device (bench.py's roofline denominator of the exact fp32 LSTM kernel: MEASURED_PEAKS.json
// carries HBM and bf16 tensor figures only).  An unrolled FFMA microkernel: 2 x 1024 threads per SM, 8 independent accumulator
// chains per thread (the FMA pipe's latency is 4 cycles; 64 resident warps x 8 chains cover it many times over), register
// operands only.  Measurement hook - the product path never calls it.
#include "ape_common.cuh"
#include "ape_b200.h"

namespace ape {
__global__ void __launch_bounds__(1024, 2) ffma_peak_kernel(float* out, int iters, float x, float y) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (float)(threadIdx.x + j) * 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], x, y);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 12345.678f) out[blockIdx.x] = s;     // never true for the operands used: keeps the chains alive
}
}  // namespace ape

// *tflops = best of `reps` launches of 2 * sm_count CTAs x 1024 threads x iters x 128 FFMA (2 flops each); synchronises `stream`.
extern "C" int ape_selftest_ffma_peak(float* scratch, int iters, int reps, float* tflops, void* stream) {
    if (!scratch || !tflops || iters < 1 || reps < 1) return APE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sm_count = 0;
    APE_CUDA_TRY(cudaGetDevice(&dev));
    APE_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    cudaEvent_t e0, e1;
    APE_CUDA_TRY(cudaEventCreate(&e0));
    APE_CUDA_TRY(cudaEventCreate(&e1));
    const int grid = 2 * sm_count;
    float best = 0.0f;
    for (int r = 0; r <= reps; ++r) {                         // (launch 0 is the warm-up)
        APE_CUDA_TRY(cudaEventRecord(e0, st));
        ape::ffma_peak_kernel<<<grid, 1024, 0, st>>>(scratch, iters, 0.9999f, 1e-4f);
        APE_CUDA_TRY(cudaEventRecord(e1, st));
        APE_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.0f;
        APE_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const float tf = (float)((double)grid * 1024.0 * iters * 128.0 * 2.0 / (ms * 1e-3) / 1e12);
        if (r > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return ape::check_launch();
}
