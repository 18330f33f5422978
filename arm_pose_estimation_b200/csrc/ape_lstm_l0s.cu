// Stage 2, layer 0 of a SMALL call (E = B x nF <= 8 estimates): the single-stream real-time case (BASELINE configs[1]: one
// stream x 100 MC samples).  Layer 0 runs once per estimate, so for one stream it is a recurrence over ONE row: T dependent steps
// of a (I + H) x 4H matrix-vector product.  Through the tensor-core layer kernels that row costs what a 256-row tile costs - every
// step walks all gate chunks of a full MMA tile on one CTA pair: 0.07 ms of the 0.19 ms frame.  Here the 4H gate columns are split
// over a cluster of 8 CTAs whose fp32 weight slices stay resident in shared memory; every step each CTA computes its H/8 hidden
// units for all E rows with plain FFMA, writes its slice of h_t into the h buffers of all eight CTAs through distributed shared
// memory, and one cluster barrier ends the step.  Exact fp32 arithmetic (the fp32 blob, expf / tanhf), so every LSTM variant can use it.
// Output: the fp16 operand units the tensor-core layer-1 loaders read (IN_SHARED_UNITS; optionally as hi + lo pairs for the
// split-precision kernel), scaled by the consumer's 1/(1-p) before the rounding like every other producer.
#include <cstdlib>

#include "ape_common.cuh"
#include "ape_lstm_pack.h"
#include "ape_lstm_tc_args.cuh"
#include "ape_umma.cuh"

namespace ape {
namespace l0s {

constexpr int NCTA = 8, THREADS = 256, MAX_E = 8, UNIT_ROWS = 128;

struct Args {
    const float* Wp;            // layer 0 of the fp32 blob: K-major [kin_pad + H][4H], column 4u + g
    const float* bias;          // [4H]
    int I, kin_pad, T, E, nF, frame0, feat_ring, dense;
    const float* in;            // feature ring [B][feat_ring][I], or dense windows [E][T][I]
    const int32_t* stream_frames;
    uint4* out_hi;              // [T][2][H/8][128 rows] fp16 units (row = estimate; CTA 0's half of pair tile 0)
    uint4* out_lo;              // the same for fp16(v - fp16(v)), or null
    float out_scale;
};

__device__ __forceinline__ void st_cluster_f32(uint32_t local_addr, uint32_t cta, float v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

template <int H>
__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(THREADS, 1) layer0_small_kernel(const Args a) {
    constexpr int HU = H / NCTA, NCOL = 4 * HU, KSPLIT = THREADS / NCOL;   // this CTA's units / gate columns; K slices per column
    static_assert(THREADS % NCOL == 0 && HU % 8 == 0, "layout");
    extern __shared__ __align__(16) float sm[];
    const int Ktot = a.kin_pad + H;
    float* sW = sm;                                  // [Ktot][NCOL]
    float* sBias = sW + (size_t)Ktot * NCOL;         // [NCOL]
    float* sX = sBias + NCOL;                        // [MAX_E][64]
    float* sH = sX + MAX_E * 64;                     // [2][MAX_E][H]   h_{t-1} / h_t of ALL units (filled by all eight CTAs)
    float* sG = sH + 2 * MAX_E * H;                  // [KSPLIT][MAX_E][NCOL] partial gate sums
    float* sC = sG + KSPLIT * MAX_E * NCOL;          // [MAX_E][HU] cell state of this CTA's units
    float* sO = sC + MAX_E * HU;                     // [MAX_E][HU] this CTA's h_t (for the unit output)

    const int tid = threadIdx.x;
    const uint32_t rank = umma::cluster_ctarank();
    const int E = a.E, T = a.T;

    for (int i = tid; i < Ktot * NCOL; i += THREADS) {
        const int k = i / NCOL, j = i - k * NCOL;
        sW[i] = __ldg(a.Wp + (size_t)k * 4 * H + rank * NCOL + j);
    }
    for (int i = tid; i < NCOL; i += THREADS) sBias[i] = __ldg(a.bias + rank * NCOL + i);
    for (int i = tid; i < MAX_E * HU; i += THREADS) sC[i] = 0.0f;
    for (int i = tid; i < 2 * MAX_E * H; i += THREADS) sH[i] = 0.0f;
    __syncthreads();
    umma::cluster_sync();                            // every CTA's buffers are initialised before anyone writes into them remotely

    const int col = tid % NCOL, ks = tid / NCOL;
    const int k_per = (Ktot + KSPLIT - 1) / KSPLIT, k0 = ks * k_per, k1 = min(Ktot, k0 + k_per);
    for (int t = 0; t < T; ++t) {
        const float* hprev = sH + (size_t)(t & 1) * MAX_E * H;
        float* hnext = sH + (size_t)((t + 1) & 1) * MAX_E * H;
        // x_t of every estimate (zero-padded to kin_pad)
        for (int i = tid; i < E * a.kin_pad; i += THREADS) {
            const int e = i / a.kin_pad, k = i - e * a.kin_pad;
            float v = 0.0f;
            if (k < a.I) {
                if (a.dense) {
                    v = __ldg(a.in + ((size_t)e * T + t) * a.I + k);
                } else {                               // sliding window, clamped at frame 0 (estimator.py:96-97)
                    const int b = e / a.nF;
                    int fw = stream_frame0(a.stream_frames, a.frame0, b) + e % a.nF - T + 1 + t;
                    fw = fw < 0 ? 0 : fw;
                    v = __ldg(a.in + ((size_t)b * a.feat_ring + fw % a.feat_ring) * a.I + k);
                }
            }
            sX[e * 64 + k] = v;
        }
        __syncthreads();
        // partial gate sums of column `col` over this thread's K slice, for every estimate
        float acc[MAX_E];
#pragma unroll
        for (int e = 0; e < MAX_E; ++e) acc[e] = 0.0f;
        for (int k = k0; k < k1; ++k) {
            if (k >= a.kin_pad && t == 0) break;       // h_0 = 0
            const float w = sW[(size_t)k * NCOL + col];
            const float* src = k < a.kin_pad ? sX + k : hprev + (k - a.kin_pad);
            const int stride = k < a.kin_pad ? 64 : H;
#pragma unroll
            for (int e = 0; e < MAX_E; ++e)
                if (e < E) acc[e] = fmaf(w, src[e * stride], acc[e]);
        }
#pragma unroll
        for (int e = 0; e < MAX_E; ++e)
            if (e < E) sG[((size_t)ks * MAX_E + e) * NCOL + col] = acc[e];
        __syncthreads();
        // cell update of (estimate, unit); h_t goes to all eight CTAs
        for (int i = tid; i < E * HU; i += THREADS) {
            const int e = i / HU, u = i - e * HU;
            float g[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float s = sBias[4 * u + q];
                for (int z = 0; z < KSPLIT; ++z) s += sG[((size_t)z * MAX_E + e) * NCOL + 4 * u + q];
                g[q] = s;
            }
            const float gi = 1.0f / (1.0f + expf(-g[0])), gf = 1.0f / (1.0f + expf(-g[1]));
            const float gg = tanhf(g[2]), go = 1.0f / (1.0f + expf(-g[3]));
            const float c = fmaf(gf, sC[i], gi * gg), h = go * tanhf(c);
            sC[i] = c;
            sO[i] = h;
            const uint32_t dst = umma::smem_u32(hnext + (size_t)e * H + rank * HU + u);
#pragma unroll
            for (uint32_t r = 0; r < NCTA; ++r) st_cluster_f32(dst, r, h);
        }
        __syncthreads();
        // this CTA's operand units of step t: k-groups rank * HU / 8 .. + HU / 8 - 1
        for (int i = tid; i < E * (HU / 8); i += THREADS) {
            const int e = i / (HU / 8), jl = i - e * (HU / 8);
            const float* h8 = sO + (size_t)e * HU + 8 * jl;
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float v0 = h8[2 * q] * a.out_scale, v1 = h8[2 * q + 1] * a.out_scale;
                const __half2 hh = __floats2half2_rn(v0, v1);
                const float2 hf = __half22float2(hh);
                const __half2 ll = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                hi[q] = *reinterpret_cast<const uint32_t*>(&hh);
                lo[q] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            const size_t at = ((size_t)(t * 2) * (H / 8) + rank * (HU / 8) + jl) * UNIT_ROWS + e;
            a.out_hi[at] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (a.out_lo) a.out_lo[at] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        umma::cluster_sync();                        // h_t is complete in every CTA (release / acquire at cluster scope)
    }
}

template <int H> static size_t smem_bytes(int kin_pad) {
    constexpr int HU = H / NCTA, NCOL = 4 * HU, KSPLIT = THREADS / NCOL;
    return ((size_t)(kin_pad + H) * NCOL + NCOL + MAX_E * 64 + 2 * MAX_E * H + KSPLIT * MAX_E * NCOL + 2 * MAX_E * HU) * sizeof(float);
}

bool enabled() {
    static const bool on = std::getenv("APE_NO_L0S") == nullptr;
    return on;
}

bool supported(int H, int I, long long E) { return (H == 128 || H == 256) && E >= 1 && E <= MAX_E && ape_pack_kin_pad(0, I, H) <= 64; }

// layer 0 of `g` into out_hi (/ out_lo): [T][2][H/8][128] units, row e of CTA 0's half
int launch(const ape_lstm_args* g, uint4* out_hi, uint4* out_lo, float out_scale, cudaStream_t st) {
    const long long E = (long long)g->B * g->nF;
    if (!supported(g->H, g->I, E)) return APE_ERR_UNSUPPORTED;
    Args a{};
    a.Wp = g->weights + ape_pack_layer_offset(0, g->I, g->H);
    a.bias = g->weights + ape_pack_bias_offset(0, g->I, g->H);
    a.I = g->I; a.kin_pad = ape_pack_kin_pad(0, g->I, g->H); a.T = g->T; a.E = (int)E; a.nF = g->nF; a.frame0 = g->frame0;
    a.feat_ring = g->feat_ring; a.dense = g->x_dense != nullptr;
    a.in = g->x_dense ? g->x_dense : g->feat_ring_buf;
    a.stream_frames = g->stream_frames;
    a.out_hi = out_hi; a.out_lo = out_lo; a.out_scale = out_scale;
    if (g->H == 256) {
        const size_t smem = smem_bytes<256>(a.kin_pad);
        APE_CUDA_TRY(cudaFuncSetAttribute(layer0_small_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        layer0_small_kernel<256><<<NCTA, THREADS, smem, st>>>(a);
    } else {
        const size_t smem = smem_bytes<128>(a.kin_pad);
        APE_CUDA_TRY(cudaFuncSetAttribute(layer0_small_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        layer0_small_kernel<128><<<NCTA, THREADS, smem, st>>>(a);
    }
    return check_launch();
}

}  // namespace l0s
}  // namespace ape
