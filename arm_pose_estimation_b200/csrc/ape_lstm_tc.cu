// Stage 2, tensor-core variant: layers >= 1 of the MC-dropout LSTM on tcgen05 with TMEM accumulators.
//
// One CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns a tile of 256 (estimate, MC-sample) rows - 128 per CTA,
// one row per TMEM lane - for all T steps of one layer.  Everything the recurrence needs stays on chip:
//   * the layer's gate weights as fp16 in shared memory, split by gate column across the pair (each CTA holds
//     64 of every 128-column chunk; the hardware shares the halves), loaded once per CTA;
//   * x_t and h_{t-1} as fp16 A-operand tiles in shared memory (K-major no-swizzle, see ape_umma.cuh), written by
//     the epilogue warps themselves - h_t never goes through global memory inside a layer;
//   * the 128 x 4H fp32 gate pre-activations in TMEM (H/32 chunks of 32 hidden units x 4 gates = 128 columns);
//   * the fp32 cell state c in registers (8 units x H/32 chunks per thread).
// Warp roles per CTA: 16 epilogue warps (TMEM -> registers, bias, sigmoid/tanh cell update, dropout-masked
// operand packing, output layer) and one MMA-issue warp; only the leader CTA's issuer runs.  Hand-offs go
// through mbarriers: tcgen05.commit (multicast to both CTAs) publishes accumulators, the epilogue warps of both
// CTAs arrive on the leader's barriers to publish operands / free accumulator chunks.  The x-part of step t+1
// is issued into a chunk as soon as the epilogue has drained it, so the tensor pipe runs under the epilogue of
// step t; only the recurrent half waits for h_t.
// Precision: fp16 operands, fp32 accumulate ("x1").  The fp32 FFMA kernel remains the exact path; the host
// API picks per model (see estimate/batched.py).
#include "ape_common.cuh"
#include "ape_lstm_pack.h"
#include "ape_lstm_plan.cuh"
#include "ape_umma.cuh"

namespace ape {
namespace tc {

constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = EPI_THREADS + 32;          // + the MMA-issue warp
constexpr int ROWS = 128;                          // rows per CTA = TMEM lanes
constexpr float LOG2E = 1.4426950408889634f;
constexpr float EX2_CLAMP = 40.0f;                 // (1 + 2^40)^3 still fits fp32

enum { IN_SHARED_F32 = 0, IN_UNITS_F16 = 1 };
enum { BAR_X_READY = 0, BAR_X_DONE = 1, BAR_H_READY = 2, BAR_ACC_READY = 3, BAR_SLOT_FREE = 7, BAR_COUNT = 11 };

struct TcLayerArgs {
    const uint8_t* W;          // this layer: [cta 2][part x|h][chunk][k-group][64 gate columns][8 halfs]
    const float* bias_s;       // [4H] column c = 4u+g, pre-multiplied by -log2(e) (i, f, o) / -2 log2(e) (g)
    int T;
    int in_mode;
    const void* in;
    int in_R;
    int nF, frame0, rows, n;
    int mask_mode;
    const uint8_t* masks;
    int gap, n_gaps;
    uint64_t seed;
    uint32_t stream_id0;
    float keep_scale;
    uint32_t keep_thr16;
    uint4* out_units;          // [pair tile][T][cta][k-group][128 rows] 16-byte units of fp16 (h_t * out_scale), or null
    float out_scale;           // the consumer's 1/(1-p): scaling BEFORE the fp16 rounding keeps it a single rounding
    const float* Wo;
    const float* bo;
    int O;
    float* preds;
    int pred_ring, all_steps, n_out;
    int n_pair_tiles;
};

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// One LSTM cell from the four ex2 ARGUMENTS  p = -(gate + bias) * log2e  (g gate: * 2 log2e):
//   i*g~ = (1 - e_g) / ((1 + e_i)(1 + e_g)),  f = 1 / (1 + e_f)  share one reciprocal; h = o * tanh(c) another.
__device__ __forceinline__ void lstm_cell(float pi, float pf, float pg, float po, float& c, float& h) {
    const float ei = ex2_approx(fminf(pi, EX2_CLAMP)), ef = ex2_approx(fminf(pf, EX2_CLAMP));
    const float eg = ex2_approx(fminf(pg, EX2_CLAMP)), eo = ex2_approx(fminf(po, EX2_CLAMP));
    const float ab = (1.0f + ei) * (1.0f + eg), cf = 1.0f + ef;
    const float num = fmaf(c, ab, (1.0f - eg) * cf);
    c = num * rcp_approx(ab * cf);
    const float ec = ex2_approx(fminf(c * (-2.0f * LOG2E), EX2_CLAMP));
    h = (1.0f - ec) * rcp_approx((1.0f + eo) * (1.0f + ec));
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

template <int H>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) lstm_layer_tc_kernel(TcLayerArgs a) {
    using namespace umma;
    constexpr int NCH = H / 32, KG = H / 8;
    constexpr uint32_t B_TILE = KG * 64 * 16;                  // one (part, chunk) weight tile of this CTA
    constexpr uint32_t W_BYTES = 2 * NCH * B_TILE;
    constexpr uint32_t A_BYTES = KG * ROWS * 16;               // one A-operand tile
    constexpr uint32_t LBO_A = ROWS * 16, LBO_B = 64 * 16, SBO = 128;
    constexpr uint32_t TMEM_COLS = NCH * 128 <= 32 ? 32 : NCH * 128 <= 64 ? 64 : NCH * 128 <= 128 ? 128 : NCH * 128 <= 256 ? 256 : 512;

    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sW = smem;
    uint8_t* sAh = sW + W_BYTES;                               // [2][A_BYTES]
    uint8_t* sAx = sAh + 2 * A_BYTES;
    float* sBias = reinterpret_cast<float*>(sAx + A_BYTES);    // [4H]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 4 * H);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int T = a.T;

    // ---- one-time set-up: weights + bias -> smem, TMEM, barriers ------------------------------------------------
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.W + (size_t)rank * W_BYTES);
        uint4* dst = reinterpret_cast<uint4*>(sW);
        for (uint32_t i = tid; i < W_BYTES / 16; i += THREADS) dst[i] = __ldg(src + i);
        for (int i = tid; i < 4 * H; i += THREADS) sBias[i] = a.bias_s[i];
    }
    if (warp == EPI_WARPS) {
        tmem_alloc<2>(tmem_slot, TMEM_COLS);
        tmem_relinquish<2>();
    }
    if (tid == 0) {
        mbar_init(&bars[BAR_X_READY], 2 * EPI_WARPS);
        mbar_init(&bars[BAR_X_DONE], 1);
        mbar_init(&bars[BAR_H_READY], 2 * EPI_WARPS);
        for (int c = 0; c < 4; ++c) {
            mbar_init(&bars[BAR_ACC_READY + c], 1);
            mbar_init(&bars[BAR_SLOT_FREE + c], 2 * EPI_WARPS);
        }
        mbar_init_fence();
    }
    fence_proxy_async_smem();
    fence_before_sync();
    cluster_sync();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp < EPI_WARPS) {
        // =================================== epilogue / operand warps ==============================================
        const int q = warp & 3, s = warp >> 2;                 // TMEM lane quarter; 8-unit slice of every 32-unit chunk
        const int row_l = 32 * q + lane;                       // local row == TMEM lane == tid & 127
        const uint32_t t_lane = (uint32_t)(32 * q) << 16;
        uint32_t ph_acc = 0, ph_xdone = 0;
        float cst[NCH][8];

        for (int tile = cluster_id; tile < a.n_pair_tiles; tile += n_clusters) {
            const int row = (tile * 2 + (int)rank) * ROWS + row_l;
            const bool valid = row < a.rows;
            const int e = valid ? row / a.n : 0, smp = valid ? row - e * a.n : 0;
            const int b = e / a.nF, f = a.frame0 + e % a.nF;
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int u = 0; u < 8; ++u) cst[c][u] = 0.0f;

            // masked, scaled fp16 x_t of this thread's row for k-groups s, s+4, ... -> sAx
            auto load_x = [&](int t) {
#pragma unroll
                for (int i = 0; i < KG / 4; ++i) {
                    const int j = s + 4 * i;
                    float v[8];
                    uint32_t keep = 0xFFu;
                    if (valid) {
                        if (a.in_mode == IN_SHARED_F32) {
                            const float* src = reinterpret_cast<const float*>(a.in) +
                                               (((size_t)(e / a.in_R) * T + t) * H + j * 8) * a.in_R + e % a.in_R;
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] = __ldg(src + (size_t)k * a.in_R);
                        } else {
                            const uint4 u4 = __ldg(reinterpret_cast<const uint4*>(a.in) +
                                                   ((((size_t)tile * T + t) * 2 + rank) * KG + j) * ROWS + row_l);
                            const __half2* h2 = reinterpret_cast<const __half2*>(&u4);
#pragma unroll
                            for (int k = 0; k < 4; ++k) { const float2 f2 = __half22float2(h2[k]); v[2 * k] = f2.x; v[2 * k + 1] = f2.y; }
                        }
                        if (a.mask_mode == APE_MASK_INJECTED) {
                            const uint2 m = __ldg(reinterpret_cast<const uint2*>(
                                a.masks + ((((size_t)e * a.n_gaps + a.gap) * T + t) * a.n + smp) * H + j * 8));
                            keep = 0;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                keep |= ((m.x >> (8 * k)) & 0xFFu ? 1u : 0u) << k;
                                keep |= ((m.y >> (8 * k)) & 0xFFu ? 1u : 0u) << (4 + k);
                            }
                        } else if (a.mask_mode == APE_MASK_PHILOX) {
                            keep = philox_keep8(a.seed, a.stream_id0 + (uint32_t)b, (uint32_t)f, (uint32_t)smp, (uint32_t)a.gap,
                                                (uint32_t)t, (uint32_t)j, a.keep_thr16);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = 0.0f;
                    }
                    const float scale = a.in_mode == IN_SHARED_F32 ? a.keep_scale : 1.0f;   // fp16 units arrive pre-scaled
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = ((keep >> k) & 1u) ? v[k] * scale : 0.0f;
                    const uint4 packed = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]),
                                                    pack_half2(v[6], v[7]));
                    *reinterpret_cast<uint4*>(sAx + unit_offset(ROWS, row_l, j)) = packed;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(&bars[BAR_X_READY], 0);
            };

            load_x(0);
            for (int t = 0; t < T; ++t) {
                if (t + 1 < T) {
                    mbar_wait(&bars[BAR_X_DONE], ph_xdone);    // the x-part MMAs of step t have read sAx
                    ph_xdone ^= 1;
                    load_x(t + 1);
                }
                uint8_t* sAh_next = sAh + ((t + 1) & 1) * A_BYTES;
                // last step of the last layer: h_T goes to the output layer only - keep it in fp32, [unit][row], in the
                // two operand tiles nobody reads any more (sAx: units < H/2, the idle h tile: the rest)
                const bool final_f32 = a.preds != nullptr && t == T - 1;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    mbar_wait(&bars[BAR_ACC_READY + c], ph_acc);
                    fence_after_sync();
                    uint32_t r[32];
                    tmem_ld_x32(tmem + t_lane + (uint32_t)(c * 128 + 32 * s), r);
                    tmem_ld_wait();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(&bars[BAR_SLOT_FREE + c], 0);

                    const float4* bias4 = reinterpret_cast<const float4*>(sBias + (c * 32 + 8 * s) * 4);
                    float hv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float4 bs = bias4[u];
                        const float pi = fmaf(__uint_as_float(r[4 * u + 0]), -LOG2E, bs.x);
                        const float pf = fmaf(__uint_as_float(r[4 * u + 1]), -LOG2E, bs.y);
                        const float pg = fmaf(__uint_as_float(r[4 * u + 2]), -2.0f * LOG2E, bs.z);
                        const float po = fmaf(__uint_as_float(r[4 * u + 3]), -LOG2E, bs.w);
                        lstm_cell(pi, pf, pg, po, cst[c][u], hv[u]);
                    }
                    const int j = 4 * c + s;                   // k-group of units 32c + 8s .. + 7
                    if (final_f32) {
                        float* dst = reinterpret_cast<float*>(c < NCH / 2 ? sAx : sAh_next) + ((8 * j) % (H / 2)) * ROWS + row_l;
#pragma unroll
                        for (int u = 0; u < 8; ++u) dst[u * ROWS] = hv[u];
                    } else {
                        *reinterpret_cast<uint4*>(sAh_next + unit_offset(ROWS, row_l, j)) =
                            make_uint4(pack_half2(hv[0], hv[1]), pack_half2(hv[2], hv[3]), pack_half2(hv[4], hv[5]), pack_half2(hv[6], hv[7]));
                    }
                    if (a.out_units) {
                        const float os = a.out_scale;
                        a.out_units[((((size_t)tile * T + t) * 2 + rank) * KG + j) * ROWS + row_l] =
                            make_uint4(pack_half2(hv[0] * os, hv[1] * os), pack_half2(hv[2] * os, hv[3] * os),
                                       pack_half2(hv[4] * os, hv[5] * os), pack_half2(hv[6] * os, hv[7] * os));
                    }
                }
                ph_acc ^= 1;
                if (t + 1 < T) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(&bars[BAR_H_READY], 0);
                }
                if (a.preds && (a.all_steps || t == T - 1)) {  // output_layer (nn_models.py:189)
                    epi_bar_sync();
                    if (valid) {
                        for (int o = s; o < a.O; o += 4) {
                            const float* w = a.Wo + (size_t)o * H;
                            float sum = __ldg(a.bo + o);
                            if (final_f32) {
                                const float* h0 = reinterpret_cast<const float*>(sAx) + row_l;
                                const float* h1 = reinterpret_cast<const float*>(sAh_next) + row_l;
#pragma unroll 8
                                for (int k = 0; k < H / 2; ++k) sum = fmaf(__ldg(w + k), h0[k * ROWS], sum);
#pragma unroll 8
                                for (int k = 0; k < H / 2; ++k) sum = fmaf(__ldg(w + H / 2 + k), h1[k * ROWS], sum);
                            } else {
#pragma unroll 4
                                for (int j = 0; j < KG; ++j) {
                                    const uint4 u4 = *reinterpret_cast<const uint4*>(sAh_next + unit_offset(ROWS, row_l, j));
                                    const __half2* h2 = reinterpret_cast<const __half2*>(&u4);
                                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + 8 * j));
                                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + 8 * j + 4));
                                    const float2 p0 = __half22float2(h2[0]), p1 = __half22float2(h2[1]);
                                    const float2 p2 = __half22float2(h2[2]), p3 = __half22float2(h2[3]);
                                    sum = fmaf(w0.x, p0.x, sum); sum = fmaf(w0.y, p0.y, sum); sum = fmaf(w0.z, p1.x, sum); sum = fmaf(w0.w, p1.y, sum);
                                    sum = fmaf(w1.x, p2.x, sum); sum = fmaf(w1.y, p2.y, sum); sum = fmaf(w1.z, p3.x, sum); sum = fmaf(w1.w, p3.y, sum);
                                }
                            }
                            if (a.all_steps) a.preds[((size_t)row * T + t) * a.O + o] = sum;
                            else a.preds[((((size_t)b * a.pred_ring + f % a.pred_ring) * a.n_out) + smp) * a.O + o] = sum;
                        }
                    }
                    epi_bar_sync();                            // next tile's load_x reuses sAx: readers first
                }
            }
        }
    } else if (rank == 0 && lane == 0) {
        // =================================== MMA issuer (leader CTA, one thread) =====================================
        const uint32_t idesc = make_idesc_f16(256, 128);
        const uint32_t aX = smem_u32(sAx), aH = smem_u32(sAh), wB = smem_u32(sW);
        uint32_t ph_xready = 0, ph_hready = 0, ph_slot = 0;
        bool first = true;
        for (int tile = cluster_id; tile < a.n_pair_tiles; tile += n_clusters) {
            for (int t = 0; t < T; ++t) {
                mbar_wait_cluster(&bars[BAR_X_READY], ph_xready);
                ph_xready ^= 1;
                fence_after_sync();
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (!first) { mbar_wait_cluster(&bars[BAR_SLOT_FREE + c], ph_slot); fence_after_sync(); }
                    const uint32_t wb = wB + (0 * NCH + c) * B_TILE;
#pragma unroll
                    for (int k2 = 0; k2 < KG / 2; ++k2)
                        mma_f16<2>(tmem + c * 128, make_desc(aX + k2 * 2 * LBO_A, LBO_A, SBO), make_desc(wb + k2 * 2 * LBO_B, LBO_B, SBO),
                                   idesc, k2 > 0 ? 1u : 0u);
                }
                if (!first) ph_slot ^= 1;
                first = false;
                if (t + 1 < T) commit_pair(&bars[BAR_X_DONE], 0x3);
                if (t > 0) {
                    mbar_wait_cluster(&bars[BAR_H_READY], ph_hready);
                    ph_hready ^= 1;
                    fence_after_sync();
                }
                const uint32_t ah = aH + (t & 1) * A_BYTES;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (t > 0) {
                        const uint32_t wb = wB + (1 * NCH + c) * B_TILE;
#pragma unroll
                        for (int k2 = 0; k2 < KG / 2; ++k2)
                            mma_f16<2>(tmem + c * 128, make_desc(ah + k2 * 2 * LBO_A, LBO_A, SBO), make_desc(wb + k2 * 2 * LBO_B, LBO_B, SBO),
                                       idesc, 1u);
                    }
                    commit_pair(&bars[BAR_ACC_READY + c], 0x3);
                }
            }
        }
    }
    __syncwarp();
    fence_before_sync();
    cluster_sync();
    if (warp == EPI_WARPS) tmem_dealloc<2>(tmem, TMEM_COLS);
}

template <int H> static size_t smem_bytes() {
    return (size_t)2 * (H / 32) * (H / 8) * 64 * 16 + 3 * (size_t)(H / 8) * ROWS * 16 + 4 * H * sizeof(float) + BAR_COUNT * 8 + 16;
}

template <int H> static int launch(const TcLayerArgs& a, int sm_count, cudaStream_t st) {
    const size_t smem = smem_bytes<H>();
    APE_CUDA_TRY(cudaFuncSetAttribute(lstm_layer_tc_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int clusters = sm_count / 2;
    if (clusters > a.n_pair_tiles) clusters = a.n_pair_tiles;
    lstm_layer_tc_kernel<H><<<2 * clusters, THREADS, smem, st>>>(a);
    return check_launch();
}

}  // namespace tc
}  // namespace ape

// bytes of the fp16 weight blob of the tensor-core path: per layer >= 1 the two CTAs' tiles, then the scaled bias
extern "C" int ape_lstm_tc_blob_bytes(int H, int L, int64_t* bytes) {
    if (!bytes || H < 32 || H % 32 != 0 || L < 1) return APE_ERR_BAD_ARG;
    *bytes = (int64_t)(L - 1) * ((int64_t)2 * 2 * (H / 32) * (H / 8) * 64 * 16 + (int64_t)4 * H * 4);
    return APE_OK;
}

extern "C" int ape_mc_lstm_tc_supported(int H) { return (H == 64 || H == 128) ? 1 : 0; }

extern "C" int ape_mc_lstm_tc_workspace_bytes(int I, int H, int L, int T, int O, int E, int n_samples, uint64_t* bytes) {
    using namespace ape;
    if (!bytes || I < 1 || H < 1 || L < 1 || T < 1 || O < 1 || E < 0 || n_samples < 1) return APE_ERR_BAD_ARG;
    if (!ape_mc_lstm_tc_supported(H)) return APE_ERR_UNSUPPORTED;
    FmaPlan p;
    const int rc = make_plan(I, H, L, T, E, n_samples, &p);
    if (rc != APE_OK) return rc;
    const uint64_t pair_tiles = ((uint64_t)E * n_samples + 255) / 256;
    const uint64_t units = L > 2 ? pair_tiles * 256 * T * H * 2 : 0;           // fp16 h_t of one layer
    *bytes = p.seq0_bytes + (L > 3 ? 2 : 1) * ((units + 255) & ~(uint64_t)255) + 512;
    return APE_OK;
}

extern "C" int ape_mc_lstm_tc(const ape_lstm_args* g, void* stream) {
    using namespace ape;
    int rc = check_lstm_args(g);
    if (rc != APE_OK) return rc;
    if (!ape_mc_lstm_tc_supported(g->H) || g->L < 2) return APE_ERR_UNSUPPORTED;
    if (!g->weights_tc) return APE_ERR_BAD_ARG;
    const long long E = (long long)g->B * g->nF;
    if (E == 0) return APE_OK;
    FmaPlan p;
    rc = make_plan(g->I, g->H, g->L, g->T, E, g->n_samples, &p);
    if (rc != APE_OK) return rc;
    if (!g->workspace) return APE_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sm_count = 0;
    APE_CUDA_TRY(cudaGetDevice(&dev));
    APE_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));

    const long long rows = E * g->n_samples;
    const int pair_tiles = (int)((rows + 255) / 256);
    const size_t units_bytes = ((size_t)pair_tiles * 256 * g->T * g->H * 2 + 255) & ~(size_t)255;
    char* wsp = (char*)(((uintptr_t)g->workspace + 255) & ~(uintptr_t)255);
    float* seq0 = (float*)wsp;
    uint4* units[2] = {(uint4*)(wsp + p.seq0_bytes), (uint4*)(wsp + p.seq0_bytes + units_bytes)};

    cudaEvent_t ev[17] = {};
    const bool prof = g->layer_ms != nullptr && g->L <= 16;
    if (prof) for (int l = 0; l <= g->L; ++l) APE_CUDA_TRY(cudaEventCreate(&ev[l]));
    if (prof) APE_CUDA_TRY(cudaEventRecord(ev[0], st));

    rc = fma_launch_layer(g, 0, p, nullptr, seq0, st);          // layer 0: once per estimate, fp32
    if (rc != APE_OK) return rc;
    if (prof) APE_CUDA_TRY(cudaEventRecord(ev[1], st));

    const int H = g->H, NCH = H / 32, KG = H / 8;
    const size_t w_layer = (size_t)2 * 2 * NCH * KG * 64 * 16, layer_bytes = w_layer + (size_t)4 * H * 4;
    for (int l = 1; l < g->L; ++l) {
        const bool last = l == g->L - 1;
        tc::TcLayerArgs a{};
        a.W = (const uint8_t*)g->weights_tc + (size_t)(l - 1) * layer_bytes;
        a.bias_s = (const float*)(a.W + w_layer);
        a.T = g->T;
        a.in_mode = l == 1 ? tc::IN_SHARED_F32 : tc::IN_UNITS_F16;
        a.in = l == 1 ? (const void*)seq0 : (const void*)units[(l - 2) & 1];
        a.in_R = 16 * p.rt0;
        a.nF = g->nF; a.frame0 = g->frame0; a.rows = (int)rows; a.n = g->n_samples;
        a.mask_mode = g->mask_mode; a.masks = g->masks; a.gap = l - 1; a.n_gaps = g->L - 1;
        a.seed = g->philox_seed; a.stream_id0 = g->stream_id0;
        a.keep_scale = g->mask_mode == APE_MASK_NONE ? 1.0f : 1.0f / (1.0f - g->dropout_p);
        a.keep_thr16 = keep_threshold16(g->dropout_p);
        a.out_units = last ? nullptr : units[(l - 1) & 1];
        a.out_scale = a.keep_scale;
        a.Wo = g->weights + ape_pack_out_offset(g->I, g->H, g->L);
        a.bo = a.Wo + (size_t)g->O * g->H;
        a.O = g->O;
        a.preds = last ? g->preds : nullptr;
        a.pred_ring = g->pred_ring; a.all_steps = g->all_steps; a.n_out = g->n_samples;
        a.n_pair_tiles = pair_tiles;
        rc = H == 128 ? tc::launch<128>(a, sm_count, st) : tc::launch<64>(a, sm_count, st);
        if (rc != APE_OK) return rc;
        if (prof) APE_CUDA_TRY(cudaEventRecord(ev[l + 1], st));
    }
    if (prof) {
        APE_CUDA_TRY(cudaStreamSynchronize(st));
        for (int l = 0; l < g->L; ++l) APE_CUDA_TRY(cudaEventElapsedTime(&g->layer_ms[l], ev[l], ev[l + 1]));
        for (int l = 0; l <= g->L; ++l) cudaEventDestroy(ev[l]);
    }
    return APE_OK;
}
