// Stage 2, tensor-core variant for H = 256 (the watch-only and pocket models): the same CTA-pair tcgen05 design as
// ape_lstm_tc.cu, re-balanced for a layer whose fp16 gate weights (1 MB) no longer fit the pair's shared memory.
//
//   * A CTA PAIR (cluster of 2, cta_group::2, M = 256) owns 256 (estimate, MC-sample) rows for all T steps of a layer.
//     x_t, h_{t-1} and h_t live in shared memory as fp16 A-operand tiles (3 x 64 KB per CTA), so the recurrence never
//     leaves the SM.
//   * Gate WEIGHTS STREAM from L2 through a ring of 4 KB pieces (32 K-values x this CTA's 64 gate columns of a 32-unit
//     chunk), filled by one thread per CTA with bulk asynchronous copies (cp.async.bulk, mbarrier complete_tx).  The MMA
//     issuer consumes pieces in a fixed per-step order (walk_step below, evaluated on the host into a table), the
//     producer walks the same table, and every piece is released by a tcgen05.commit once its two K=16 MMAs have retired.
//     The peer CTA forwards its ring's "full" events to the leader, because the leader's MMAs read both CTAs' halves of B.
//   * TMEM: two 128-column accumulator slots (one 32-unit chunk x 4 gates each, ping-pong) + 256 columns holding the
//     fp32 CELL STATE c (one column per hidden unit, read and rewritten by the epilogue with tcgen05.ld / .st) - at
//     H = 256 the state does not fit the register file next to the epilogue's working set.  On the last step of the
//     last layer the same columns receive h_T in fp32 and the output layer reads them back.
//   * Schedule: chunk c of step t+1 is issued into its slot as soon as the epilogue of step t has drained chunk
//     c + NCH - 2 from it - first the x-part, then the recurrent K-slices already published - so when the last slice of
//     h_t lands only 2 pieces (4 MMAs) stand between it and the first accumulator of step t+1.
// Per step and CTA the tensor pipe needs 128 pieces x 128 cycles = 16.4 k cycles; the cell update of 128 rows x 256
// units needs 14.3 k cycles of the 16-lane MUFU pipe: the two are balanced, unlike H = 128 where the cell update binds.
// L2 -> SM weight traffic is 512 KB per CTA and step (~20 B/clk/SM, half of the measured L2 cap).
// Precision and dropout keying are those of ape_lstm_tc.cu (fp16 operands rounded once, fp32 accumulate and state).
#include "ape_common.cuh"
#include "ape_lstm_pack.h"
#include "ape_lstm_tc_args.cuh"
#include "ape_umma.cuh"

namespace ape {
namespace tcs {

using tc::TcLayerArgs;
using tc::ex2_approx;
using tc::rcp_approx;
using tc::LOG2E;
using tc::EX2_CLAMP;

constexpr int EPI_WARPS = 16, LOAD_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int MMA_WARP = EPI_WARPS + LOAD_WARPS;   // leader CTA: MMA issuer; peer CTA: forwards "piece landed" to the leader
constexpr int TMA_WARP = MMA_WARP + 1;             // one lane per CTA fills the weight ring
constexpr int THREADS = (TMA_WARP + 1) * 32;       // 832
constexpr int ROWS = 128;                          // rows per CTA = TMEM lanes
constexpr int NSLOT = 2;                           // accumulator slots of 128 TMEM columns
constexpr int PIECE_KG = 4;                        // k-groups (of 8 K-values) per streamed weight piece
constexpr uint32_t KG_BYTES_B = 64 * 16;           // one k-group of a 64-column weight tile
constexpr uint32_t PIECE_BYTES = PIECE_KG * KG_BYTES_B;
constexpr uint32_t BAR_BLOCK_BYTES = 512;
constexpr uint32_t SMEM_LIMIT = 227 * 1024;
constexpr int TRACE_T = 3;                         // tracing (args.trace): the step of the first tile of CTA 0 that is stamped

template <int H> struct Cfg {
    static constexpr int NCH = H / 32, KG = H / 8;
    static constexpr uint32_t A_BYTES = KG * ROWS * 16;
    static constexpr uint32_t FIXED = 3 * A_BYTES + 4 * H * 4 + BAR_BLOCK_BYTES;
    static constexpr int NP_FIT = (SMEM_LIMIT - FIXED) / PIECE_BYTES;
    static constexpr int NP = NP_FIT > 12 ? 12 : NP_FIT;            // ring depth (pieces)
    static constexpr uint32_t SMEM = FIXED + NP * PIECE_BYTES;
    static constexpr uint32_t C_COL = NSLOT * 128;                   // first TMEM column of the cell state
    static constexpr uint32_t TMEM_COLS = 512;
    enum {
        BAR_X_READY = 0, BAR_X_DONE = 1, BAR_ACC_READY = 2, BAR_SLOT_FREE = BAR_ACC_READY + NSLOT,
        BAR_H_READY = BAR_SLOT_FREE + NSLOT, BAR_W_FULL = BAR_H_READY + NCH, BAR_W_EMPTY = BAR_W_FULL + NP,
        BAR_COUNT = BAR_W_EMPTY + NP
    };
    static_assert(NCH % 2 == 0 && NCH > NSLOT, "chunks alternate between the two accumulator slots");
    static_assert(C_COL + H <= TMEM_COLS, "accumulator slots + cell state must fit the 512 TMEM columns");
    static_assert(NP >= 4, "weight ring too shallow");
    static_assert(BAR_COUNT * 8 + 16 <= BAR_BLOCK_BYTES, "barrier block too small");
};

// The order in which ONE step consumes weight pieces and signals its hand-offs.  The MMA issuer, the ring producer and
// the peer's forwarder all walk this function with their own visitor, which is what keeps the ring in lock-step.
//   chunk_begin(c)                      slot c & 1 is about to be refilled
//   piece(c, is_h, kg0, nkg, first)     chunk c (+)= A[:, k-groups kg0 .. kg0+nkg) x W_c[k-groups]^T  (x-part or recurrent part)
//   h_wait(ks)                          units 32 ks .. 32 ks + 31 of h_{t-1} must be in the operand tile
//   acc_done(c) / x_done()              chunk c complete / the x tile has been consumed
template <int NCH, class V>
static void walk_step(bool first_step, int kgx, V& v) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        v.chunk_begin(c);
        for (int kg0 = 0; kg0 < kgx; kg0 += PIECE_KG) v.piece(c, false, kg0, kgx - kg0 < PIECE_KG ? kgx - kg0 : PIECE_KG, kg0 == 0);
        if (c == NCH - 1) v.x_done();
        if (first_step) {                                      // h_{-1} = 0: no recurrent half
            v.acc_done(c);
        } else if (c >= NSLOT) {                               // refilled inside the step: all of h_{t-1} is there
#pragma unroll
            for (int ks = 0; ks < NCH; ++ks) v.piece(c, true, ks * PIECE_KG, PIECE_KG, false);
            v.acc_done(c);
        } else {                                               // refilled under the tail of the previous step's epilogue
            const int avail = NCH - NSLOT + c;                 // K-slices published before this slot was drained
#pragma unroll
            for (int ks = 0; ks < NCH; ++ks) {
                if (ks < avail) {
                    if (c == 0) v.h_wait(ks);
                    v.piece(c, true, ks * PIECE_KG, PIECE_KG, false);
                }
            }
            v.h_wait(avail);
#pragma unroll
            for (int cc = 0; cc < NSLOT; ++cc) {
                if (cc <= c) {
                    v.piece(cc, true, avail * PIECE_KG, PIECE_KG, false);
                    if (c == NSLOT - 1) v.acc_done(cc);
                }
            }
        }
    }
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

// ---- the step schedule as a table ----------------------------------------------------------------------------------
// walk_step is evaluated on the HOST into a table that travels as a kernel parameter (constant bank): the issuer, the
// producer and the forwarder run short table-driven loops whose operands stay in uniform registers.  (Inlining the walk -
// ~130 pieces of straight-line code per step - or reading the table from shared memory made the single issuing thread
// spend ~450 cycles per piece against the 128 cycles its two MMAs occupy the tensor pipe.)
enum : uint32_t {
    E_SLOT1 = 1u << 0,        // accumulator slot (chunk & 1)
    E_IS_H = 1u << 1,         // recurrent piece: A = h_{t-1} (else x_t)
    E_TWO = 1u << 2,          // 4 k-groups = two K=16 MMAs (else one)
    E_FIRST = 1u << 3,        // first MMA of the chunk: overwrite the accumulator
    E_PRE_SLOT = 1u << 4,     // before: wait until the epilogue has drained this slot
    E_PRE_H = 1u << 5,        // before: wait for this piece's K-slice of h_{t-1}
    E_POST_ACC = 1u << 6,     // after: the chunk is complete
    E_POST_X = 1u << 7,       // after: the x tile has been consumed
    E_KS_SHIFT = 8,           // [8, 12)  K-slice (kg0 / PIECE_KG) of a recurrent piece
    E_SRC_SHIFT = 12          // [12, 32) byte offset of the piece inside this CTA's weight tiles / 1024
};
constexpr int MAX_ENTRIES = 128;                               // NCH * (KG / PIECE_KG + NCH) for H = 256

struct Schedule {                                              // [0]: first step of a tile (no recurrent half), [1]: later steps
    uint2 e[2][MAX_ENTRIES];                                   // .x flags / K-slice / weight offset, .y A-operand offset >> 4
    uint32_t n[2];
};

struct Recorder {
    uint2* tab;
    int n, kgx, KG;
    uint32_t pre;
    void chunk_begin(int) { pre |= E_PRE_SLOT; }
    void piece(int c, bool is_h, int kg0, int nkg, bool first) {
        const uint32_t src_kg = (uint32_t)c * (uint32_t)(kgx + KG) + (is_h ? (uint32_t)kgx : 0u) + (uint32_t)kg0;   // 1 KB per k-group
        tab[n].x = (uint32_t)(c & 1) | (is_h ? E_IS_H : 0u) | (nkg > 2 ? E_TWO : 0u) | (first ? E_FIRST : 0u) | pre |
                   ((uint32_t)(kg0 / PIECE_KG) << E_KS_SHIFT) | (src_kg << E_SRC_SHIFT);
        tab[n].y = (uint32_t)kg0 * (ROWS * 16) >> 4;           // k-group offset inside the x / h operand tile, descriptor units
        ++n;
        pre = 0;
    }
    void h_wait(int) { pre |= E_PRE_H; }                       // always the K-slice of the piece that follows
    void acc_done(int) { tab[n - 1].x |= E_POST_ACC; }
    void x_done() { tab[n - 1].x |= E_POST_X; }
};

template <int H>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) lstm_layer_tcs_kernel(const __grid_constant__ TcLayerArgs a, const __grid_constant__ Schedule sched) {
    using namespace umma;
    using C = Cfg<H>;
    constexpr int NCH = C::NCH, KG = C::KG, NP = C::NP;
    constexpr uint32_t A_BYTES = C::A_BYTES;
    constexpr uint32_t LBO_A = ROWS * 16, LBO_B = KG_BYTES_B, SBO = 128;
    constexpr uint32_t C_COL = C::C_COL;

    const int kgx = a.kgx, T = a.T;
    const uint32_t chunk_bytes = (uint32_t)(kgx + KG) * KG_BYTES_B;   // x k-groups then h k-groups of one chunk
    const uint32_t w_bytes = NCH * chunk_bytes;                       // one CTA's half of the layer

    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sAh = smem;                                       // [2][A_BYTES]
    uint8_t* sAx = sAh + 2 * A_BYTES;                          // [A_BYTES] (kgx k-groups used)
    uint8_t* sW = sAx + A_BYTES;                               // [NP][PIECE_BYTES] weight ring
    float* sBias = reinterpret_cast<float*>(sW + NP * PIECE_BYTES);   // [4H]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 4 * H);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::BAR_COUNT);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    // ---- one-time set-up: bias -> smem, TMEM, barriers ---------------------------------------------------------------
    for (int i = tid; i < 4 * H; i += THREADS) sBias[i] = a.bias_s[i];
    if (warp == MMA_WARP) {
        tmem_alloc<2>(tmem_slot, C::TMEM_COLS);
        tmem_relinquish<2>();
    }
    if (tid == 0) {
        mbar_init(&bars[C::BAR_X_READY], 2 * LOAD_WARPS);
        mbar_init(&bars[C::BAR_X_DONE], 1);
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(&bars[C::BAR_ACC_READY + s], 1);
            mbar_init(&bars[C::BAR_SLOT_FREE + s], 2 * EPI_WARPS);
        }
        for (int c = 0; c < NCH; ++c) mbar_init(&bars[C::BAR_H_READY + c], 2 * EPI_WARPS);
        for (int p = 0; p < NP; ++p) {
            // leader: its own copy (arrive.expect_tx + bytes) and the peer's "my copy landed" arrival; peer: its own copy
            mbar_init(&bars[C::BAR_W_FULL + p], rank == 0 ? 2 : 1);
            mbar_init(&bars[C::BAR_W_EMPTY + p], 1);
        }
        mbar_init_fence();
    }
    fence_proxy_async_smem();
    fence_before_sync();
    cluster_sync();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp < EPI_WARPS) {
        // =================================== epilogue warps ===========================================================
        // warp (q, s): rows 32q..32q+31 (its TMEM lane quarter) x the 8 hidden units 8s..8s+7 of EVERY 32-unit chunk.
        const int q = warp & 3, s = warp >> 2;
        const int row_l = 32 * q + lane;                       // local row == TMEM lane
        const uint32_t t_lane = (uint32_t)(32 * q) << 16;
        const bool warp_live = 32 * q < a.rpc;                 // a quarter without rows only keeps the barrier protocol going
        uint32_t gstep = 0;

        for (int tile = cluster_id; tile < a.n_pair_tiles; tile += n_clusters) {
            const int row = (tile * 2 + (int)rank) * a.rpc + row_l;
            const bool valid = row_l < a.rpc && row < a.rows;

            for (int t = 0; t < T; ++t, ++gstep) {
                uint8_t* sAh_next = sAh + ((t + 1) & 1) * A_BYTES;
                const bool final_out = a.preds != nullptr && t == T - 1;
                // chunk c uses slot c & 1 for the (gstep * NCH/2 + c/2)-th time
                const uint32_t par0 = (gstep * (uint32_t)(NCH / 2)) & 1u;
                if (!warp_live) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        mbar_wait_wd(&bars[C::BAR_ACC_READY + (c & 1)], par0 ^ ((c >> 1) & 1));
                        if (lane == 0) {
                            mbar_arrive_leader(&bars[C::BAR_SLOT_FREE + (c & 1)], rank);
                            if (t + 1 < T) mbar_arrive_leader(&bars[C::BAR_H_READY + c], rank);
                        }
                    }
                } else {
                    // Half-passes of 4 hidden units (16 accumulator columns + 4 state columns), double-buffered: the TMEM
                    // loads of half-pass hp+1 are in flight while the cells of half-pass hp are computed.
                    uint32_t rbuf[2][16], cbuf[2][4];
                    float hlo[4];
                    mbar_wait_wd(&bars[C::BAR_ACC_READY + 0], par0);
                    fence_after_sync();
                    tmem_ld_x16(tmem + t_lane + (uint32_t)(32 * s), rbuf[0]);
                    if (t > 0) tmem_ld_x4(tmem + t_lane + C_COL + (uint32_t)(8 * s), cbuf[0]);
#pragma unroll
                    for (int hp = 0; hp < 2 * NCH; ++hp) {
                        const int c = hp >> 1, half = hp & 1;
                        const uint32_t* r = rbuf[hp & 1];
                        tmem_ld_wait();                        // this half-pass's columns have landed
                        if (a.trace && blockIdx.x == 0 && tile == cluster_id && warp == 0 && lane == 0 && (t == TRACE_T || t == TRACE_T - 1))
                            a.trace[512 + (t - TRACE_T + 1) * 16 + hp] = clock64();
                        if (half == 1) {                       // chunk c fully drained: the issuer may refill its slot
                            fence_before_sync();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_leader(&bars[C::BAR_SLOT_FREE + (c & 1)], rank);
                        }
                        if (hp + 1 < 2 * NCH) {
                            const int c1 = (hp + 1) >> 1, h1 = (hp + 1) & 1;
                            if (half == 1) {
                                mbar_wait_wd(&bars[C::BAR_ACC_READY + (c1 & 1)], par0 ^ ((c1 >> 1) & 1));
                                fence_after_sync();
                            }
                            tmem_ld_x16(tmem + t_lane + (uint32_t)((c1 & 1) * 128 + 32 * s + 16 * h1), rbuf[(hp + 1) & 1]);
                            if (t > 0) tmem_ld_x4(tmem + t_lane + C_COL + (uint32_t)(32 * c1 + 8 * s + 4 * h1), cbuf[(hp + 1) & 1]);
                        }
                        const float4* bias4 = reinterpret_cast<const float4*>(sBias + (c * 32 + 8 * s + 4 * half) * 4);
                        // the 4 cells advance in lock-step through the transcendental stages (independent MUFU ops back to back):
                        //   e = 2^-(gate+bias)  ->  i*g~ and f share one reciprocal  ->  2^(-2c)  ->  h = o * tanh(c)
                        float ev[16], hv[4], num[4], den[4], cn[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float4 bs = bias4[u];
                            ev[4 * u + 0] = ex2_approx(fminf(fmaf(__uint_as_float(r[4 * u + 0]), -LOG2E, bs.x), EX2_CLAMP));
                            ev[4 * u + 1] = ex2_approx(fminf(fmaf(__uint_as_float(r[4 * u + 1]), -LOG2E, bs.y), EX2_CLAMP));
                            ev[4 * u + 2] = ex2_approx(fminf(fmaf(__uint_as_float(r[4 * u + 2]), -2.0f * LOG2E, bs.z), EX2_CLAMP));
                            // (o gate unclamped: an infinite e_o only makes the reciprocal below 0; e_c is finite since |c| <= T)
                            ev[4 * u + 3] = ex2_approx(fmaf(__uint_as_float(r[4 * u + 3]), -LOG2E, bs.w));
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float cprev = t > 0 ? __uint_as_float(cbuf[hp & 1][u]) : 0.0f;
                            const float ab = (1.0f + ev[4 * u + 0]) * (1.0f + ev[4 * u + 2]), cf = 1.0f + ev[4 * u + 1];
                            num[u] = fmaf(cprev, ab, (1.0f - ev[4 * u + 2]) * cf);
                            den[u] = rcp_approx(ab * cf);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            cn[u] = num[u] * den[u];
                            num[u] = ex2_approx(cn[u] * (-2.0f * LOG2E));   // e_c = 2^(-2c log2e), |c| <= T: no overflow
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) hv[u] = (1.0f - num[u]) * rcp_approx((1.0f + ev[4 * u + 3]) * (1.0f + num[u]));

                        // cell state back to its TMEM columns; on the last step of the last layer they receive h_T (fp32)
                        // instead, which is what the output layer reads
                        tmem_st_x4(tmem + t_lane + C_COL + (uint32_t)(32 * c + 8 * s + 4 * half),
                                   __float_as_uint(final_out ? hv[0] : cn[0]), __float_as_uint(final_out ? hv[1] : cn[1]),
                                   __float_as_uint(final_out ? hv[2] : cn[2]), __float_as_uint(final_out ? hv[3] : cn[3]));

                        const int j = 4 * c + s;               // k-group of units 32c + 8s .. + 7
                        if (half == 1) {
                            if (!final_out)
                                *reinterpret_cast<uint4*>(sAh_next + unit_offset(ROWS, row_l, j)) =
                                    make_uint4(pack_half2(hlo[0], hlo[1]), pack_half2(hlo[2], hlo[3]), pack_half2(hv[0], hv[1]), pack_half2(hv[2], hv[3]));
                            if (t + 1 < T) {                   // publish this chunk's slice of h_t: its K-slice of the next recurrent
                                fence_proxy_async_smem();      // product can be issued while later chunks still run
                                __syncwarp();
                                if (lane == 0) mbar_arrive_leader(&bars[C::BAR_H_READY + c], rank);
                            }
                            if (a.out_units) {
                                const float os = a.out_scale;
                                a.out_units[((((size_t)tile * T + t) * 2 + rank) * KG + j) * ROWS + row_l] =
                                    make_uint4(pack_half2(hlo[0] * os, hlo[1] * os), pack_half2(hlo[2] * os, hlo[3] * os),
                                               pack_half2(hv[0] * os, hv[1] * os), pack_half2(hv[2] * os, hv[3] * os));
                            }
                        } else {
#pragma unroll
                            for (int u = 0; u < 4; ++u) hlo[u] = hv[u];
                        }
                    }
                    tmem_st_wait();                            // the state columns are re-read one step later (and by other warps below)
                }
                if (final_out) {                               // output_layer (nn_models.py:189), last step only
                    fence_before_sync();
                    epi_bar_sync();                            // h_T of all 4 unit groups of this lane quarter is in TMEM
                    fence_after_sync();
                    if (warp_live) {
                        // this thread: outputs o = s, s+4, ... of its row (<= 5 for O <= 20), one sweep over the H state columns
                        float acc[5];
#pragma unroll
                        for (int i = 0; i < 5; ++i) acc[i] = (s + 4 * i < a.O) ? __ldg(a.bo + s + 4 * i) : 0.0f;
#pragma unroll 1
                        for (int blk = 0; blk < NCH; ++blk) {
                            uint32_t hr[32];
                            tmem_ld_x32(tmem + t_lane + C_COL + (uint32_t)(32 * blk), hr);
                            tmem_ld_wait();
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
#pragma unroll
                                for (int i = 0; i < 5; ++i) {
                                    if (s + 4 * i < a.O) {
                                        const float4 w = __ldg(reinterpret_cast<const float4*>(a.Wo + (size_t)(s + 4 * i) * H + 32 * blk + 4 * k));
                                        acc[i] = fmaf(w.x, __uint_as_float(hr[4 * k + 0]), acc[i]);
                                        acc[i] = fmaf(w.y, __uint_as_float(hr[4 * k + 1]), acc[i]);
                                        acc[i] = fmaf(w.z, __uint_as_float(hr[4 * k + 2]), acc[i]);
                                        acc[i] = fmaf(w.w, __uint_as_float(hr[4 * k + 3]), acc[i]);
                                    }
                                }
                            }
                        }
                        if (valid) {
                            const int e = row / a.n, smp = row - e * a.n;
                            const int b = e / a.nF, f = a.frame0 + e % a.nF;
#pragma unroll
                            for (int i = 0; i < 5; ++i) {
                                const int o = s + 4 * i;
                                if (o < a.O) {
                                    float* dst = a.preds + (((size_t)b * a.pred_ring + f % a.pred_ring) * a.n_out) * a.O + o;
                                    if (a.n == 1 && a.n_out > 1) for (int s2 = 0; s2 < a.n_out; ++s2) dst[(size_t)s2 * a.O] = acc[i];
                                    else dst[(size_t)smp * a.O] = acc[i];
                                }
                            }
                        }
                    }
                    fence_before_sync();
                    epi_bar_sync();                            // the next tile's first cell update rewrites the state columns
                    fence_after_sync();
                }
            }
        }
    } else if (warp < MMA_WARP) {
        // =================================== operand-loader warps: x_t -> sAx ===========================================
        const int row_l = (tid - EPI_THREADS) & (ROWS - 1);    // two threads per row, each half of the x k-groups
        const int half = (tid - EPI_THREADS) >> 7;
        uint32_t ph_xdone = 0;
        bool first = true;
        for (int tile = cluster_id; tile < a.n_pair_tiles; tile += n_clusters) {
            const int row = (tile * 2 + (int)rank) * a.rpc + row_l;
            const bool valid = row_l < a.rpc && row < a.rows;
            const int e = valid ? row / a.n : 0, smp = valid ? row - e * a.n : 0;
            const int b = e / a.nF, f = a.frame0 + e % a.nF;
            for (int t = 0; t < T; ++t) {
                if (a.in_mode == tc::IN_UNITS || a.in_mode == tc::IN_SHARED_UNITS) {
                    // The first batch of 8 k-groups is fetched and masked before anything waits; only the shared-memory
                    // stores (and the remaining batches, which would not fit the register file) sit behind "the previous
                    // x tile has been consumed".
                    const uint4* src = reinterpret_cast<const uint4*>(a.in);
                    if (a.in_mode == tc::IN_UNITS) src += ((((size_t)tile * T + t) * 2 + rank) * KG) * ROWS + row_l;
                    else src += ((((size_t)(e >> (a.in_rpc_shift + 1)) * T + t) * 2 + ((e >> a.in_rpc_shift) & 1)) * KG) * ROWS +
                                (e & ((1 << a.in_rpc_shift) - 1));
                    constexpr int KH = KG / 2, BK = 8;
                    const int j0 = half * KH;
                    src += (size_t)j0 * ROWS;
#pragma unroll 1
                    for (int b0 = 0; b0 < KH; b0 += BK) {
                        uint4 pre[BK];
#pragma unroll
                        for (int jj = 0; jj < BK; ++jj) pre[jj] = valid ? __ldg(src + (size_t)(b0 + jj) * ROWS) : make_uint4(0, 0, 0, 0);
#pragma unroll
                        for (int jj = 0; jj < BK; ++jj) {
                            const int j = j0 + b0 + jj;
                            if (a.mask_mode == APE_MASK_PHILOX) {
                                const uint4 m = philox_keep_halfmask(a.seed, a.stream_id0 + (uint32_t)b, (uint32_t)f, (uint32_t)smp,
                                                                     (uint32_t)a.gap, (uint32_t)t, (uint32_t)j, a.keep_thr16);
                                pre[jj].x &= m.x; pre[jj].y &= m.y; pre[jj].z &= m.z; pre[jj].w &= m.w;
                            } else if (a.mask_mode == APE_MASK_INJECTED && valid) {
                                const uint2 m = __ldg(reinterpret_cast<const uint2*>(
                                    a.masks + ((((size_t)e * a.n_gaps + a.gap) * T + t) * a.n + smp) * H + j * 8));
                                pre[jj].x &= ((m.x & 0xFFu) ? 0xFFFFu : 0u) | ((m.x & 0xFF00u) ? 0xFFFF0000u : 0u);
                                pre[jj].y &= ((m.x & 0xFF0000u) ? 0xFFFFu : 0u) | ((m.x & 0xFF000000u) ? 0xFFFF0000u : 0u);
                                pre[jj].z &= ((m.y & 0xFFu) ? 0xFFFFu : 0u) | ((m.y & 0xFF00u) ? 0xFFFF0000u : 0u);
                                pre[jj].w &= ((m.y & 0xFF0000u) ? 0xFFFFu : 0u) | ((m.y & 0xFF000000u) ? 0xFFFF0000u : 0u);
                            }
                        }
                        if (b0 == 0) {
                            const bool trl = a.trace && blockIdx.x == 0 && tile == cluster_id && t == TRACE_T && tid == EPI_THREADS;
                            if (trl) a.trace[580] = clock64();
                            if (!first) { mbar_wait_wd(&bars[C::BAR_X_DONE], ph_xdone); ph_xdone ^= 1; }
                            first = false;
                            if (trl) a.trace[581] = clock64();
                        }
#pragma unroll
                        for (int jj = 0; jj < BK; ++jj) *reinterpret_cast<uint4*>(sAx + unit_offset(ROWS, row_l, j0 + b0 + jj)) = pre[jj];
                    }
                } else {                                       // layer 0: fp32 features (window of the ring, or dense rows)
                    if (!first) { mbar_wait_wd(&bars[C::BAR_X_DONE], ph_xdone); ph_xdone ^= 1; }
                    first = false;
                    const float* src = nullptr;
                    if (valid) {
                        if (a.in_mode == tc::IN_DENSE_F32) {
                            src = reinterpret_cast<const float*>(a.in) + ((size_t)row * T + t) * a.Kin;
                        } else {                               // sliding window, clamped at frame 0 (estimator.py:96-97)
                            int fw = f - T + 1 + t;
                            fw = fw < 0 ? 0 : fw;
                            src = reinterpret_cast<const float*>(a.in) + ((size_t)b * a.feat_ring + fw % a.feat_ring) * a.Kin;
                        }
                    }
                    for (int j = half; j < kgx; j += 2) {
                        float v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[k] = (valid && 8 * j + k < a.Kin) ? __ldg(src + 8 * j + k) : 0.0f;
                        *reinterpret_cast<uint4*>(sAx + unit_offset(ROWS, row_l, j)) =
                            make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&bars[C::BAR_X_READY], rank);
                if (a.trace && blockIdx.x == 0 && tile == cluster_id && t == TRACE_T && tid == EPI_THREADS) a.trace[582] = clock64();
            }
        }
    } else if (warp == MMA_WARP) {
        if (rank == 0) {
            // =============================== MMA issuer (leader CTA, one lane) ============================================
            // Software-pipelined by hand: the barrier probe of piece i+1 (and its table entry) is issued before the MMAs of
            // piece i, so its latency hides under them - a single thread is all that feeds the tensor pipe here.
            if (lane == 0) {
                const uint32_t idesc = make_idesc_f16(256, 128);
                const uint64_t dX = make_desc(smem_u32(sAx), LBO_A, SBO), dH0 = make_desc(smem_u32(sAh), LBO_A, SBO);
                const uint64_t dW = make_desc(smem_u32(sW), LBO_B, SBO);
                const uint32_t bar_full = smem_u32(&bars[C::BAR_W_FULL]), bar_empty = smem_u32(&bars[C::BAR_W_EMPTY]);
                uint32_t ph_xready = 0, wslot = 0, wphase = 0, gchunk = 0, hphase = 0;
                bool ok = false;                               // result of the early probe of the current piece's barrier
                for (int tile = cluster_id; tile < a.n_pair_tiles; tile += n_clusters) {
                    for (int t = 0; t < T; ++t) {
                        long long* tr = (a.trace && blockIdx.x == 0 && tile == cluster_id && t == TRACE_T) ? a.trace : nullptr;
                        mbar_wait_wd(&bars[C::BAR_X_READY], ph_xready);
                        ph_xready ^= 1;
                        if (tr) tr[576] = clock64();
                        const uint64_t dH = desc_advance(dH0, (t & 1) * A_BYTES);
                        const int which = t == 0 ? 0 : 1;
                        const uint32_t n_ent = sched.n[which];
                        uint2 e = sched.e[which][0];
#pragma unroll 1
                        for (uint32_t i = 0; i < n_ent; ++i) {
                            const uint2 e_next = sched.e[which][i + 1 < n_ent ? i + 1 : i];
                            if (e.x & (E_PRE_SLOT | E_PRE_H)) {
                                if (e.x & E_PRE_SLOT) {        // the epilogue has drained the chunk that used this slot before
                                    if (gchunk >= NSLOT) mbar_wait_wd(&bars[C::BAR_SLOT_FREE + (e.x & E_SLOT1)], ((gchunk >> 1) & 1) ^ 1);
                                    ++gchunk;
                                    if (tr) tr[560 + ((gchunk - 1) & 7)] = clock64();
                                }
                                if (e.x & E_PRE_H) {           // this piece's K-slice of h_{t-1} has been published
                                    mbar_wait_wd(&bars[C::BAR_H_READY + ((e.x >> E_KS_SHIFT) & 0xF)], hphase);
                                    if (tr) tr[568 + ((e.x >> E_KS_SHIFT) & 0xF)] = clock64();
                                }
                            }
                            uint32_t spins = 0;                // the piece is in BOTH CTAs' rings
                            while (!ok) { ok = mbar_try_wait_addr(bar_full + wslot * 8, wphase); if (++spins > (1u << 24)) __trap(); }
                            fence_after_sync();
                            if (tr && i < 128) { tr[2 * i] = clock64(); tr[600 + i] = spins; }
                            const uint32_t nslot = wslot + 1 == NP ? 0 : wslot + 1, nphase = wslot + 1 == NP ? wphase ^ 1 : wphase;
                            ok = mbar_test_wait_addr(bar_full + nslot * 8, nphase);      // early, non-blocking probe of the NEXT piece
                            const uint64_t ad = ((e.x & E_IS_H) ? dH : dX) + e.y;
                            const uint64_t bd = dW + wslot * (PIECE_BYTES >> 4);
                            const uint32_t d_tmem = tmem + (e.x & E_SLOT1) * 128;
                            mma_f16<2>(d_tmem, ad, bd, idesc, (e.x & E_FIRST) ? 0u : 1u);
                            if (e.x & E_TWO) mma_f16<2>(d_tmem, ad + (2 * LBO_A >> 4), bd + (2 * LBO_B >> 4), idesc, 1u);
                            commit_pair_addr(bar_empty + wslot * 8, 0x3);      // both CTAs' producers may refill this slot
                            if (e.x & E_POST_ACC) commit_pair(&bars[C::BAR_ACC_READY + (e.x & E_SLOT1)], 0x3);
                            if (e.x & E_POST_X) commit_pair(&bars[C::BAR_X_DONE], 0x3);
                            if (tr && i < 128) tr[2 * i + 1] = clock64();
                            wslot = nslot; wphase = nphase;
                            e = e_next;
                        }
                        if (t > 0) hphase ^= 1;
                    }
                }
            }
        } else if (lane == 0) {
            // =============================== peer CTA: forward "piece landed in my ring" to the leader ==================
            uint32_t wslot = 0, wphase = 0;
            for (int tile = cluster_id; tile < a.n_pair_tiles; tile += n_clusters)
                for (int t = 0; t < T; ++t) {
                    const uint32_t n_ent = sched.n[t == 0 ? 0 : 1];
#pragma unroll 1
                    for (uint32_t i = 0; i < n_ent; ++i) {
                        mbar_wait_wd(&bars[C::BAR_W_FULL + wslot], wphase);
                        mbar_arrive_remote(&bars[C::BAR_W_FULL + wslot], 0);
                        if (++wslot == NP) { wslot = 0; wphase ^= 1; }
                    }
                }
        }
    } else if (lane == 0) {
        // =================================== weight-ring producer (one lane per CTA) =====================================
        const uint8_t* Wc = a.W + (size_t)rank * w_bytes;      // this CTA's half of the layer's weight tiles
        uint32_t wslot = 0, wphase = 0;
        bool wrapped = false;
        for (int tile = cluster_id; tile < a.n_pair_tiles; tile += n_clusters)
            for (int t = 0; t < T; ++t) {
                long long* tr = (a.trace && blockIdx.x == 0 && tile == cluster_id && t == TRACE_T) ? a.trace : nullptr;
                const int which = t == 0 ? 0 : 1;
                const uint32_t n_ent = sched.n[which];
#pragma unroll 1
                for (uint32_t i = 0; i < n_ent; ++i) {
                    const uint32_t ex = sched.e[which][i].x;
                    if (wrapped) mbar_wait_wd(&bars[C::BAR_W_EMPTY + wslot], wphase ^ 1);   // previous occupant consumed
                    if (tr && i < 128) tr[256 + 2 * i] = clock64();
                    const uint32_t bytes = (ex & E_TWO) ? PIECE_BYTES : PIECE_BYTES / 2;
                    mbar_arrive_expect_tx(&bars[C::BAR_W_FULL + wslot], bytes);
                    bulk_g2s(sW + wslot * PIECE_BYTES, Wc + (size_t)(ex >> E_SRC_SHIFT) * KG_BYTES_B, bytes, &bars[C::BAR_W_FULL + wslot]);
                    if (tr && i < 128) tr[257 + 2 * i] = clock64();
                    if (++wslot == NP) { wslot = 0; wphase ^= 1; wrapped = true; }
                }
            }
    }
    __syncwarp();
    fence_before_sync();
    cluster_sync();
    if (warp == MMA_WARP) tmem_dealloc<2>(tmem, C::TMEM_COLS);
}

bool supported(int H) { return H == 256; }

int launch_layer(int H, const TcLayerArgs& a, int sm_count, cudaStream_t st) {
    if (H != 256) return APE_ERR_UNSUPPORTED;
    using C = Cfg<256>;
    if (a.kgx < 2 || a.kgx > C::KG || (a.kgx & 1)) return APE_ERR_UNSUPPORTED;
    Schedule sched{};
    for (int which = 0; which < 2; ++which) {
        Recorder r{sched.e[which], 0, a.kgx, C::KG, 0u};
        walk_step<C::NCH>(which == 0, a.kgx, r);
        if (r.n > MAX_ENTRIES) return APE_ERR_UNSUPPORTED;
        sched.n[which] = (uint32_t)r.n;
    }
    const size_t smem = C::SMEM;
    APE_CUDA_TRY(cudaFuncSetAttribute(lstm_layer_tcs_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int clusters = sm_count / 2;
    if (clusters > a.n_pair_tiles) clusters = a.n_pair_tiles;
    lstm_layer_tcs_kernel<256><<<2 * clusters, THREADS, smem, st>>>(a, sched);
    return check_launch();
}

}  // namespace tcs
}  // namespace ape
