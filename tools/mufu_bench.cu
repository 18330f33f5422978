// Micro-benchmark: sustained MUFU.EX2 / MUFU.RCP / MUFU.TANH throughput per SM on this GPU (lanes per clock), with 8 independent
// chains per thread and 4..16 warps per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanha(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE> __global__ void k(float* out, int iters, long long* cyc) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) v[i] = ex2a(v[i]);
            else if (MODE == 1) v[i] = rcpa(v[i]);
            else if (MODE == 2) v[i] = rcpa(1.0f + ex2a(v[i]));                 // 2 MUFU + 1 FADD
            else if (MODE == 3) v[i] = rcpa(fmaf(ex2a(v[i]), 1.0001f, 1.0f)) * 1.5f + 0.25f;     // 2 MUFU + 2 FMA-pipe
            else if (MODE == 4) v[i] = tanha(v[i]);                              // the cell update's only transcendental
            else v[i] = fmaf(tanha(v[i] * 0.5f), 0.5f, 0.5f);                      // sigmoid via tanh: 1 MUFU + 2 FMA-pipe
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int mode = 0; mode < 6; ++mode)
        for (int threads : {128, 256, 512, 1024}) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, threads>>>(out, iters, cyc);
                if (mode == 1) k<1><<<148, threads>>>(out, iters, cyc);
                if (mode == 2) k<2><<<148, threads>>>(out, iters, cyc);
                if (mode == 3) k<3><<<148, threads>>>(out, iters, cyc);
                if (mode == 4) k<4><<<148, threads>>>(out, iters, cyc);
                if (mode == 5) k<5><<<148, threads>>>(out, iters, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double mufu_per_thread = (double)iters * 8 * ((mode == 2 || mode == 3) ? 2 : 1);
            printf("mode %d threads/SM %4d: %lld cycles, %.2f MUFU lanes/clk/SM\n", mode, threads, h, mufu_per_thread * threads / h);
        }
    return 0;
}
