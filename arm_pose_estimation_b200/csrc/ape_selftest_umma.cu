// GPU self-test of the tcgen05 / TMEM plumbing in ape_umma.cuh: one (M = 128 * cta_group) x N x K fp16 GEMM with
// fp32 accumulation through exactly the descriptor, commit and tcgen05.ld conventions the MC-LSTM kernel uses.
// Operands arrive already packed in the canonical K-major no-swizzle layout (see ape_umma.cuh).
// tests/test_gpu_umma.py compares D with a float64 matmul.  TEST HOOK: the product path never calls it.
#include "ape_common.cuh"
#include "ape_umma.cuh"

namespace ape {

template <int CTA_GROUP, bool A_TMEM>
__global__ void __launch_bounds__(128, 1) selftest_umma_kernel(const uint4* __restrict__ a_packed, const uint4* __restrict__ b_packed,
                                                              float* __restrict__ d, int N, int K) {
    using namespace umma;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = CTA_GROUP == 2 ? cluster_ctarank() : 0;
    const int Nloc = N / CTA_GROUP, KG = K / 8;
    uint8_t* sA = smem;                              // [KG][128][16 B]
    uint8_t* sB = smem + (size_t)KG * 128 * 16;      // [KG][Nloc][16 B]

    const uint4* ga = a_packed + (size_t)rank * KG * 128;
    const uint4* gb = b_packed + (size_t)rank * KG * Nloc;
    for (int i = tid; i < KG * 128; i += 128) reinterpret_cast<uint4*>(sA)[i] = ga[i];
    for (int i = tid; i < KG * Nloc; i += 128) reinterpret_cast<uint4*>(sB)[i] = gb[i];
    fence_proxy_async_smem();

    // A_TMEM: the A operand goes through tensor memory (columns [N, N + K/2): lane = row, column = a pair of K-values)
    uint32_t ncols = 32;
    while (ncols < (uint32_t)(N + (A_TMEM ? K / 2 : 0))) ncols <<= 1;
    if (warp == 0) {
        tmem_alloc<CTA_GROUP>(&tmem_base, ncols);
        tmem_relinquish<CTA_GROUP>();
    }
    if (tid == 0) {
        mbar_init(&bar_done, 1);
        mbar_init_fence();
    }
    fence_before_sync();
    if (CTA_GROUP == 2) cluster_sync(); else __syncthreads();
    fence_after_sync();
    const uint32_t taddr = tmem_base;

    if (A_TMEM) {                                    // every thread stores ITS row (TMEM lane) of A, 4 columns per k-group
        for (int j = 0; j < KG; ++j) {
            const uint4 v = reinterpret_cast<const uint4*>(sA)[(size_t)j * 128 + tid];
            tmem_st_x4(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)(N + 4 * j), v.x, v.y, v.z, v.w);
        }
        tmem_st_wait();
        fence_before_sync();
        if (CTA_GROUP == 2) cluster_sync(); else __syncthreads();
        fence_after_sync();
    }

    if (rank == 0 && tid == 0) {
        const uint32_t idesc = make_idesc_f16(128 * CTA_GROUP, N);
        const uint32_t lboA = 128 * 16, lboB = (uint32_t)Nloc * 16;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * lboA, lboA, 128);
            const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * lboB, lboB, 128);
            if (A_TMEM) mma_f16_ts<CTA_GROUP>(taddr, taddr + (uint32_t)(N + 8 * ks), db, idesc, ks > 0 ? 1u : 0u);
            else mma_f16<CTA_GROUP>(taddr, da, db, idesc, ks > 0 ? 1u : 0u);
        }
        if (CTA_GROUP == 2) commit_pair(&bar_done, 0x3); else commit(&bar_done);
    }
    mbar_wait(&bar_done, 0);
    fence_after_sync();

    const int row = rank * 128 + warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(taddr + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) d[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
    }
    fence_before_sync();
    if (CTA_GROUP == 2) cluster_sync(); else __syncthreads();
    if (warp == 0) tmem_dealloc<CTA_GROUP>(taddr, ncols);
}

}  // namespace ape

template <int CTA_GROUP, bool A_TMEM>
static int launch_selftest(const void* a_packed, const void* b_packed, float* d, int N, int K, size_t smem, cudaStream_t st) {
    using namespace ape;
    APE_CUDA_TRY(cudaFuncSetAttribute(selftest_umma_kernel<CTA_GROUP, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CTA_GROUP); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTA_GROUP; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    APE_CUDA_TRY(cudaLaunchKernelEx(&cfg, selftest_umma_kernel<CTA_GROUP, A_TMEM>, (const uint4*)a_packed, (const uint4*)b_packed, d, N, K));
    return check_launch();
}

// cta_group: 1 | 2; add 16 to route the A operand through tensor memory (tcgen05.st + the [a_tmem] form of tcgen05.mma)
extern "C" int ape_selftest_umma(const void* a_packed, const void* b_packed, float* d, int N, int K, int cta_group, void* stream) {
    using namespace ape;
    const bool a_tmem = (cta_group & 16) != 0;
    cta_group &= ~16;
    if (!a_packed || !b_packed || !d || (cta_group != 1 && cta_group != 2)) return APE_ERR_BAD_ARG;
    if (K < 16 || K % 16 != 0 || N < 32 || N > 256 || N % 32 != 0) return APE_ERR_BAD_ARG;
    if (a_tmem && N + K / 2 > 512) return APE_ERR_BAD_ARG;
    const size_t smem = (size_t)(K / 8) * (128 + N / cta_group) * 16;
    if (smem > 200 * 1024) return APE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (cta_group == 1) return a_tmem ? launch_selftest<1, true>(a_packed, b_packed, d, N, K, smem, st) : launch_selftest<1, false>(a_packed, b_packed, d, N, K, smem, st);
    return a_tmem ? launch_selftest<2, true>(a_packed, b_packed, d, N, K, smem, st) : launch_selftest<2, false>(a_packed, b_packed, d, N, K, smem, st);
}
