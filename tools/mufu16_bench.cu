// Micro-benchmark: is MUFU.TANH.F16 issued at a higher rate than MUFU.TANH (fp32)?  8 independent chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu16_bench mufu16_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float tanha(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned tanh2(unsigned x) { unsigned y; asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned short tanh1(unsigned short x) { unsigned short y; asm volatile("tanh.approx.f16 %0, %1;" : "=h"(y) : "h"(x)); return y; }
template <int MODE> __global__ void k(float* out, int iters, long long* cyc) {
    float v[8]; unsigned w[8]; unsigned short h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = 0.001f * (threadIdx.x + i); w[i] = 0x3c003800u + threadIdx.x + i; h[i] = 0x3800 + threadIdx.x + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) v[i] = tanha(v[i]);
            else if (MODE == 1) w[i] = tanh2(w[i]);
            else h[i] = tanh1(h[i]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += v[i] + (float)w[i] + (float)h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int mode = 0; mode < 3; ++mode)
        for (int threads : {256, 512, 1024}) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, threads>>>(out, iters, cyc);
                if (mode == 1) k<1><<<148, threads>>>(out, iters, cyc);
                if (mode == 2) k<2><<<148, threads>>>(out, iters, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double vals = (double)iters * 8 * (mode == 1 ? 2 : 1);
            printf("mode %d (%s) threads/SM %4d: %lld cycles, %.2f tanh values/clk/SM\n", mode, mode == 0 ? "f32" : mode == 1 ? "f16x2" : "f16", threads, h, vals * threads / h);
        }
    return 0;
}
