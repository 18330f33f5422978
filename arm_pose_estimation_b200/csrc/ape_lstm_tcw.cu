// Stage 2, tensor-core variant for H = 128, TWO consecutive layers >= 1 of the MC-dropout LSTM in ONE launch: a wavefront.
//
// The one-layer kernel (ape_lstm_tc.cu) keeps a layer's gate weights resident in shared memory and runs ONE recurrence per
// CTA pair: every gate pre-activation of step t+1 needs all of h_t, so each step ends in a barrier over the pair's 32
// epilogue warps (their skew + the publish -> MMA -> commit -> load latency: 16 % of the epilogue warps' time), and the
// sequence between two layers round-trips HBM (157 MB per layer at 1024 streams x 100 samples).  Here a CTA pair carries
// layers A = l and B = l+1 of the same 256-row tile side by side, B one step behind A:
//
//     wave w:   A computes step (tile, t) = item w        B computes item w-1
//
// A's step only needs what A produced one wave earlier, B's step needs A's h of the previous wave (its input x2 = dropout(h_A))
// and its own h of the previous wave - so while the epilogue warps run the cell update of one layer, the tensor pipe
// already works on the other layer's gates, and nothing ever waits on a result that was produced less than half a wave ago.
// Items run on across tile boundaries (A starts the next tile while B finishes the last step of the previous one), so a
// launch pays one half-wave of fill and drain in total.
//
// What lives where (per CTA = 128 rows; one row per TMEM lane):
//   * gate weights of BOTH layers (2 x 136 KB per CTA) stream from L2 through a shared-memory ring of 18 KB slots filled by one
//     thread with bulk asynchronous copies (cp.async.bulk, mbarrier complete_tx; as in ape_lstm_tcs.cu).  A piece is one
//     operand half (x- or recurrent part, K = 128) of this CTA's 64 gate columns of a 32-unit chunk = 8 MMAs.  The weights of
//     the i, f, o gates are stored halved (sigmoid(x) = 0.5 + 0.5 tanh(x / 2): the accumulator IS the tanh argument) and every
//     x-piece carries one more K = 16 step whose first two rows are the bias as an fp16 pair (rounding + remainder); it is
//     multiplied by a constant operand tile of ones, so the bias is added by the tensor pipe and the epilogue warps - the
//     binding resource: XU pipe 60 %, issue slots 61 %, tensor pipe 44 % busy before this change - neither load nor add it;
//   * accumulators: two 128-column TMEM slots, used alternately by the chunk sequence A0..A3, B0..B3, A0.. (one issuer warp
//     per slot); h_A and h_B as packed fp16 pairs in TMEM (2 x 64 columns each, written with tcgen05.st), A operand of the
//     recurrent MMAs ([a_tmem] form);
//   * x1 (layer A's input: the previous layer's fp16 units, dropout mask ANDed in by the loader warps) and x2 = dropout(h_A)
//     (written by the epilogue warps as they produce h_A; double-buffered by item parity) are shared-memory operand tiles;
//   * the dropout keep-bits of x2 are drawn by the LOADER warps a full wave ahead (Philox, or the injected bytes; three buffers)
//     and handed to the epilogue as one 32-bit word per (row, 16 units) in shared memory; the epilogue expands it with 4 PRMT
//     (sign-replicate);
//   * the fp32 cell states (2 x 32 values per thread) live in a per-CTA 128 KB scratch that stays in L2, prefetched one
//     half-pass ahead (as in ape_lstm_tcs.cu);
//   * the inter-layer sequence never touches HBM.
// Arithmetic, operand rounding points and Philox keys are those of ape_lstm_tc.cu: results are identical.
#include "ape_common.cuh"
#include "ape_lstm_pack.h"
#include "ape_lstm_tc_args.cuh"
#include "ape_umma.cuh"
#include "ape_f32x2.cuh"

namespace ape {
namespace tcw {

using tc::TcLayerArgs;
using tc::tanh_approx;

constexpr int H = 128, NCHL = H / 32, KG = H / 8;  // chunks per layer, k-groups per operand
constexpr int EPI_WARPS = 16, LOAD_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int N_ISSUERS = 2;                       // leader CTA: two MMA issuers (issuer k owns accumulator slot k); peer CTA: the first forwards "piece landed"
constexpr int THREADS = (EPI_WARPS + LOAD_WARPS + 4) * 32;   // 896 = 7 warpgroups (one idle warp pads the issuers' one: setmaxnreg works per warpgroup)
// Which warps play which role.  The warp scheduler prefers the higher warp id among eligible warps, and the epilogue warps are
// the ones the wave waits for: APE_TCW_EPI_HIGH = 1 gives them the highest ids (issuers + ring producer 0..3, loaders 4..11,
// epilogue 12..27); 0 is the first layout (epilogue 0..15, loaders 16..23, issuers 24..25, producer 26).
#ifndef APE_TCW_F32X2
#define APE_TCW_F32X2 1
#endif
#ifndef APE_TCW_EPI_HIGH
#define APE_TCW_EPI_HIGH 1
#endif
#if APE_TCW_EPI_HIGH
constexpr int MMA_WARP = 0, TMA_WARP = 2, LOAD_WARP0 = 4, EPI_WARP0 = 12;
#else
constexpr int EPI_WARP0 = 0, LOAD_WARP0 = EPI_WARPS, MMA_WARP = EPI_WARPS + LOAD_WARPS, TMA_WARP = MMA_WARP + N_ISSUERS;
#endif
static_assert(EPI_WARP0 % 4 == 0 && LOAD_WARP0 % 4 == 0 && MMA_WARP % 4 == 0, "roles start on warpgroup boundaries (TMEM lane quarter = warp % 4)");
// failed barrier probes of the epilogue warps back off for this long (0: re-probe at once; the probe itself suspends for a few dozen cycles)
#ifndef APE_TCW_BACKOFF_NS
#define APE_TCW_BACKOFF_NS 0
#endif
// Register file re-balanced per role (launch allocation 72 x 896): the epilogue's working set (accumulator columns, gate values,
// prefetched cell state) spilled ~20 values per half-pass at 80 registers
constexpr int REGS_EPI = 96, REGS_LOAD = 40, REGS_MMA = 40;
static_assert(EPI_THREADS * REGS_EPI + LOAD_WARPS * 32 * REGS_LOAD + 128 * REGS_MMA <= THREADS * 72, "the re-balanced register file must fit the launch allocation");
#define TCW_REG_INC(n) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(n))
#define TCW_REG_DEC(n) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(n))
constexpr int ROWS = 128;
constexpr uint32_t KG_BYTES_B = 64 * 16;           // one k-group of a 64-column weight tile
constexpr uint32_t PIECE_H_BYTES = KG * KG_BYTES_B;        // 16 KB: the recurrent half of one chunk
constexpr uint32_t PIECE_X_BYTES = (KG + 2) * KG_BYTES_B;  // 18 KB: the x half + the bias k-step (k-group KG: rows 0, 1 = fp16(b), b - fp16(b))
constexpr uint32_t SLOT_BYTES = PIECE_X_BYTES;     // one ring slot
constexpr uint32_t CHUNK_BYTES = PIECE_X_BYTES + PIECE_H_BYTES;
constexpr uint32_t LAYER_BYTES = NCHL * CHUNK_BYTES;   // one CTA's half of a layer (ape_lstm_tcw_layer_bytes = 2 x this)
constexpr uint32_t A_BYTES = KG * ROWS * 16;       // one operand tile (32 KB)
constexpr uint32_t ONES_BYTES = 2 * ROWS * 16;     // the constant A tile of the bias k-step: k-group 0 = [1, 1, 0 x 6] per row, k-group 1 = 0
constexpr int NMASK = 3;                           // keep-bit buffers (item i in buffer i % 3: drawn a full wave ahead)
constexpr uint32_t MASK_BYTES = (KG / 2) * ROWS * 4;   // keep-bit words of one item: one word per (row, 2 k-groups)
#ifndef APE_TCW_NP
#define APE_TCW_NP 6
#endif
constexpr int NP = APE_TCW_NP;                              // ring depth (slots)
constexpr int NFULL = 8;                           // "piece landed" barriers, indexed by piece number (> NP: never alias)
constexpr uint32_t OUT_N = 32;                     // output product: 16 outputs x {fp16(W_o), W_o - fp16(W_o)}
constexpr uint32_t OUT_BYTES = KG * (OUT_N / 2) * 16;   // this CTA's tile of it (4 KB)
constexpr uint32_t BAR_BLOCK_BYTES = 512;
constexpr uint32_t SMEM = A_BYTES + 2 * A_BYTES + NMASK * MASK_BYTES + NP * SLOT_BYTES + ONES_BYTES + BAR_BLOCK_BYTES;
constexpr uint32_t HA_COL = 256 /* behind the two 128-column accumulator slots */, HB_COL = 384, H_COLS = H / 2, TMEM_COLS = 512;
constexpr size_t CSTATE_FLOATS = (size_t)2 * H * ROWS;     // per CTA: [layer][k-group * 2 + half][row] float4
static_assert(SMEM <= 227 * 1024, "shared memory budget");

enum {
    BAR_X1_READY = 0, BAR_X1_DONE = 1, BAR_MASK_READY = 2, BAR_ACC_READY = 5, BAR_SLOT_FREE = 7, BAR_HA_READY = 9,
    BAR_HB_READY = BAR_HA_READY + NCHL, BAR_W_FULL = BAR_HB_READY + NCHL, BAR_W_EMPTY = BAR_W_FULL + NFULL,
    BAR_OUT_READY = BAR_W_EMPTY + NP, BAR_COUNT = BAR_OUT_READY + 1
};
static_assert(BAR_COUNT * 8 + 16 <= BAR_BLOCK_BYTES, "barrier block too small");

// Barrier waits of the warps that are NOT the bottleneck (loaders, issuers, ring producer / forwarder): a failed probe - it already
// suspends the warp for a few dozen cycles - is followed by a nanosleep, because a polling warp re-issues ~7 instructions per
// probe on the issue port it shares with the epilogue warps of its SM quarter.
#ifndef APE_TCW_EARLY_FREE
#define APE_TCW_EARLY_FREE 0
#endif
#ifndef APE_TCW_BO_LOAD
#define APE_TCW_BO_LOAD 0
#endif
#ifndef APE_TCW_BO_MMA
#define APE_TCW_BO_MMA 0
#endif
#ifndef APE_TCW_BO_RING
#define APE_TCW_BO_RING 0
#endif
template <int NS> __device__ __forceinline__ void mbar_wait_bo(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!umma::mbar_try_wait(bar, parity)) {
        if (++spins > umma::MBAR_WD_SPINS) __trap();
        if (NS > 0) __nanosleep(NS);
    }
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// keep-bit word of 8 units: byte k (k = 0..3) carries unit k in bit 7 and unit 4+k in bit 6.  From the four flag words of a
// draw (unit 2i: bit 15 of word i, unit 2i+1: bit 31):
__device__ __forceinline__ uint32_t keep_word(const uint4 f) {
    return (prmt(f.x, f.y, 0x7531u) & 0x80808080u) | ((prmt(f.z, f.w, 0x7531u) >> 1) & 0x40404040u);
}
// a mask word serves two k-groups: the even one in bits 7 / 6 of every byte, the odd one in bits 5 / 4

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
lstm_pair_tcw_kernel(const __grid_constant__ TcLayerArgs a, const __grid_constant__ TcLayerArgs b) {
    using namespace umma;
    constexpr uint32_t LBO_A = ROWS * 16, LBO_B = KG_BYTES_B, SBO = 128;
    const int T = a.T;

    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sX1 = smem;                                       // [A_BYTES] layer A's input tile of the current item
    uint8_t* sX2 = sX1 + A_BYTES;                              // [2][A_BYTES] dropout(h_A) of item i in buffer i & 1
    uint32_t* sMask = reinterpret_cast<uint32_t*>(sX2 + 2 * A_BYTES);   // [NMASK][KG / 2][ROWS] keep-bit words of item i in buffer i % 3
    uint8_t* sW = reinterpret_cast<uint8_t*>(sMask) + NMASK * MASK_BYTES;   // [NP][SLOT_BYTES] weight ring
    uint8_t* sOnes = sW + NP * SLOT_BYTES;                     // [2][ROWS] units: the A operand of the bias k-step
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + ONES_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int n_my_tiles = (a.n_pair_tiles - cluster_id + n_clusters - 1) / n_clusters;
    const int NI = n_my_tiles * T;                             // items (tile, step) of this pair; waves 0 .. NI
    const bool has_out = b.preds != nullptr;

    tc::timeline_stamp(a.timeline, 0);
    for (int i = tid; i < 2 * ROWS; i += THREADS)              // halfs 0, 1 of every row = 1.0: they meet the bias rows of the x-pieces
        reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(i < ROWS ? 0x3C003C00u : 0u, 0u, 0u, 0u);
    if (warp == MMA_WARP) {
        tmem_alloc<2>(tmem_slot, TMEM_COLS);
        tmem_relinquish<2>();
    }
    if (tid == 0) {
        mbar_init(&bars[BAR_X1_READY], 2 * LOAD_WARPS);
        mbar_init(&bars[BAR_X1_DONE], N_ISSUERS);              // each issuer commits after ITS last x-part of layer A
        for (int i = 0; i < NMASK; ++i) mbar_init(&bars[BAR_MASK_READY + i], LOAD_WARPS);   // CTA-local
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[BAR_ACC_READY + i], 1);
            mbar_init(&bars[BAR_SLOT_FREE + i], 2 * EPI_WARPS);
        }
        for (int c = 0; c < NCHL; ++c) {
            mbar_init(&bars[BAR_HA_READY + c], 2 * EPI_WARPS);
            mbar_init(&bars[BAR_HB_READY + c], 2 * EPI_WARPS);
        }
        for (int p = 0; p < NFULL; ++p) mbar_init(&bars[BAR_W_FULL + p], rank == 0 ? 2 : 1);   // leader: own copy + the peer's forward
        for (int p = 0; p < NP; ++p) mbar_init(&bars[BAR_W_EMPTY + p], 1);
        mbar_init(&bars[BAR_OUT_READY], 1);
        mbar_init_fence();
    }
    fence_proxy_async_smem();
    fence_before_sync();
    cluster_sync();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    tc::timeline_stamp(a.timeline, 1);

    if (warp >= EPI_WARP0 && warp < EPI_WARP0 + EPI_WARPS) {
        TCW_REG_INC(REGS_EPI);
        // =================================== epilogue warps ===========================================================
        // warp (q, s): rows 32q..32q+31 (its TMEM lane quarter) x the 8 hidden units 8s..8s+7 of EVERY 32-unit chunk of both layers.
        const int q = warp & 3, s = (warp - EPI_WARP0) >> 2;
        const int row_l = 32 * q + lane;
        const uint32_t t_lane = (uint32_t)(32 * q) << 16;
        // everything a half-pass addresses is (one per-thread base) + (a compile-time offset):
        // (the empty asm statements make the bases opaque: the compiler then keeps them in registers instead of re-deriving
        // them from the thread / CTA index in every half-pass, which had been a fifth of the epilogue's instructions)
        unsigned long long cst_base = reinterpret_cast<unsigned long long>(
            reinterpret_cast<float4*>(a.cstate + (size_t)blockIdx.x * CSTATE_FLOATS) + (size_t)(2 * s) * ROWS + row_l);
        uint32_t acc0 = tmem + t_lane + (uint32_t)(32 * s);                          // this thread's accumulator columns of a slot
        uint32_t hst0 = tmem + t_lane + (uint32_t)(4 * s);                           // ... its 4 columns of an h buffer (+ 16 per chunk)
        uint32_t x2_base = smem_u32(sX2) + unit_offset(ROWS, row_l, s);              // ... its unit of an x2 tile (+ 4 k-groups per chunk)
        uint32_t mask_base = smem_u32(sMask + (size_t)(s >> 1) * ROWS + row_l);     // word (row, k-groups 4 cl + {0, 1} | {2, 3}) + 2 cl rows of words
        const uint32_t mask_shift = 2u * (uint32_t)(s & 1);
        uint32_t bars_local = smem_u32(bars), bars_leader;                           // barrier blocks: own CTA's, the leader's (cluster address)
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bars_leader) : "r"(bars_local), "r"(0));
        asm volatile("" : "+l"(cst_base), "+r"(acc0), "+r"(hst0), "+r"(x2_base), "+r"(mask_base), "+r"(bars_local), "+r"(bars_leader));
        float4* const cst0 = reinterpret_cast<float4*>(cst_base);
        auto arrive_leader = [&](int bar) {
            asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bars_leader + (uint32_t)bar * 8u) : "memory");
        };
        auto wait_bar = [&](int bar, uint32_t parity) {        // spin with the watchdog (a protocol bug traps instead of hanging)
            uint32_t spins = 0;
            while (!mbar_try_wait_addr(bars_local + (uint32_t)bar * 8u, parity)) {
                if (++spins > MBAR_WD_SPINS) __trap();
#if APE_TCW_BACKOFF_NS > 0
                __nanosleep(APE_TCW_BACKOFF_NS);
#endif
            }
        };
        uint32_t ph_out = 0, mbuf = NMASK - 1, mph = 1;        // mask buffer / barrier phase of A's item: w % 3, (w / 3) & 1
        // (the loop bound is re-derived here from an opaque copy of the parameter: carried over from the common prologue - compiled
        // for the 72-register launch allocation - it lived in a local-memory slot that every wave re-read through a 28 KB L1)
        int npt = a.n_pair_tiles;
        asm volatile("" : "+r"(npt));
        const int NI = ((npt - cluster_id + n_clusters - 1) / n_clusters) * T;
        int tA = 0, tB = -1, tileB = cluster_id - n_clusters;  // advanced at the top of every wave

        for (int w = 0; w <= NI; ++w) {
            const bool A_on = w < NI, B_on = w >= 1;
            if (w > 0) { if (++tA == T) tA = 0; }
            if (B_on) { if (++tB == T) tB = 0; if (tB == 0) tileB += n_clusters; }
            const bool final_out = B_on && has_out && tB == T - 1;
            const uint32_t bufA = (uint32_t)w & 1u;            // x2 buffer of A's item
            if (++mbuf == NMASK) { mbuf = 0; mph ^= 1u; }
            if (A_on) wait_bar(BAR_MASK_READY + (int)mbuf, mph);

            const int hp_begin = A_on ? 0 : 2 * NCHL, hp_end = B_on ? 4 * NCHL : 2 * NCHL;
            // One set of 16 accumulator registers and ONE load site per half-pass: the second half of a chunk is requested as soon as
            // the gate pre-activations of the first half have been formed (the load runs under the transcendentals and the cell
            // update); the first half of the NEXT chunk at the end of this chunk's second half-pass, behind a blocking wait for that
            // chunk (it was issued long ago - only its slot's refill, one chunk of MMAs, can still be outstanding).
            uint32_t r[16];
            float4 cbuf[2];
            float hlo[4];
            auto t_of = [&](int layer) { return layer ? tB : tA; };
            auto cst_at = [&](int hp1) {                       // this thread's float4 of cell state of half-pass hp1: [layer][k-group * 2 + half][row]
                return cst0 + (size_t)((hp1 >> 3) * 2 * KG + 8 * ((hp1 >> 1) & 3) + (hp1 & 1)) * ROWS;
            };
            cbuf[0] = cbuf[1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (t_of(hp_begin >> 3) > 0) cbuf[0] = __ldcg(cst_at(hp_begin));
            wait_bar(BAR_ACC_READY + 0, 0);                     // every group of 4 chunks uses each slot twice: parity = (chunk >> 1) & 1
            fence_after_sync();
            tmem_ld_x16(acc0, r);
#pragma unroll
            for (int hp = 0; hp < 4 * NCHL; ++hp) {
                if (hp < hp_begin || hp >= hp_end) continue;
                const int layer = hp >> 3, cl = (hp >> 1) & 3, half = hp & 1, slot = (hp >> 1) & 1;
                const int t_cur = t_of(layer);
                tmem_ld_wait();
                // the accumulators ARE the tanh arguments: sigmoid(x) = 0.5 + 0.5 tanh(x / 2) with the 0.5 folded into the weights
                // of the i, f, o gates, and the bias added by the tensor pipe (the bias k-step of every x-piece)
                float tg[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) tg[i] = __uint_as_float(r[i]);
#if !APE_TCW_EARLY_FREE
                if (half == 1) {                               // chunk fully drained: its issuer may refill the slot
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) arrive_leader(BAR_SLOT_FREE + slot);
                }
#endif
                const float cp[4] = {cbuf[hp & 1].x, cbuf[hp & 1].y, cbuf[hp & 1].z, cbuf[hp & 1].w};
                if (hp + 1 < hp_end) cbuf[(hp + 1) & 1] = t_of((hp + 1) >> 3) > 0 ? __ldcg(cst_at(hp + 1)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (half == 0) tmem_ld_x16(acc0 + (uint32_t)(slot * 128 + 16), r);
                uint32_t mw = 0;
                if (layer == 0 && half == 1) {
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(mw) : "r"(mask_base + (mbuf * (KG / 2) + 2u * (uint32_t)cl) * (ROWS * 4)));
                    mw <<= mask_shift;
                }

                // 5 MUFU per cell
#pragma unroll
                for (int i = 0; i < 16; ++i) tg[i] = tanh_approx(tg[i]);
#if APE_TCW_EARLY_FREE
                if (half == 0) {
                    // The chunk's second half was requested above and has had the 16 tanh issues to arrive: with it in registers the
                    // chunk is drained and its issuer may refill the slot - most of a half-pass earlier than when the second
                    // half-pass starts (the refill's issue -> commit -> wake-up chain is what the step waits for).
                    tmem_ld_wait();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) arrive_leader(BAR_SLOT_FREE + slot);
                }
#endif
                float hv[4], cn[4];
#if APE_TCW_F32X2
                // the cell update of two units per instruction (FFMA2 / FMUL2: the same IEEE operations as the scalar form, half the
                // issue slots - the epilogue warps share their issue port with the XU queue they are waiting on)
#pragma unroll
                for (int u = 0; u < 4; u += 2) {
                    const F2 half2 = splat(0.5f);
                    const F2 gi = fma2(pk(tg[4 * u + 0], tg[4 * u + 4]), half2, half2), gf = fma2(pk(tg[4 * u + 1], tg[4 * u + 5]), half2, half2);
                    const F2 c2 = fma2(gf, pk(cp[u], cp[u + 1]), gi * pk(tg[4 * u + 2], tg[4 * u + 6]));
                    cn[u] = lo(c2); cn[u + 1] = hi(c2);
                    const F2 h2 = fma2(pk(tg[4 * u + 3], tg[4 * u + 7]), half2, half2) * pk(tanh_approx(cn[u]), tanh_approx(cn[u + 1]));
                    hv[u] = lo(h2); hv[u + 1] = hi(h2);
                }
#else
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float gi = fmaf(tg[4 * u + 0], 0.5f, 0.5f), gf = fmaf(tg[4 * u + 1], 0.5f, 0.5f);
                    cn[u] = fmaf(gf, cp[u], gi * tg[4 * u + 2]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) hv[u] = fmaf(tg[4 * u + 3], 0.5f, 0.5f) * tanh_approx(cn[u]);
#endif
                if (t_cur + 1 < T) __stcg(cst_at(hp), make_float4(cn[0], cn[1], cn[2], cn[3]));

                if (half == 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) hlo[u] = hv[u];
                } else if (layer == 0) {
                    // h_A(t): fp16 pairs into the TMEM operand buffer of A's next step; dropout(h_A(t)) * 1/(1-p) as fp16 units into
                    // B's input tile (the scale is applied BEFORE the rounding, as the one-layer kernel does for its out_units)
                    if (t_cur + 1 < T)
                        tmem_st_x4(hst0 + HA_COL + (uint32_t)((t_cur + 1) & 1) * H_COLS + (uint32_t)(16 * cl),
                                   pack_half2(hlo[0], hlo[1]), pack_half2(hlo[2], hlo[3]), pack_half2(hv[0], hv[1]), pack_half2(hv[2], hv[3]));
                    const float os = a.out_scale;
                    const uint32_t mw2 = mw << 1;
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(x2_base + bufA * A_BYTES + (uint32_t)(4 * cl) * (ROWS * 16)),
                                 "r"(pack_half2(hlo[0] * os, hlo[1] * os) & prmt(mw, 0u, 0x9988u)), "r"(pack_half2(hlo[2] * os, hlo[3] * os) & prmt(mw, 0u, 0xBBAAu)),
                                 "r"(pack_half2(hv[0] * os, hv[1] * os) & prmt(mw2, 0u, 0x9988u)), "r"(pack_half2(hv[2] * os, hv[3] * os) & prmt(mw2, 0u, 0xBBAAu)) : "memory");
                    if (t_cur + 1 < T) tmem_st_wait();
                    fence_proxy_async_smem();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) arrive_leader(BAR_HA_READY + cl);
                } else {
                    if (t_cur + 1 < T || final_out) {
                        tmem_st_x4(hst0 + HB_COL + (uint32_t)((t_cur + 1) & 1) * H_COLS + (uint32_t)(16 * cl),
                                   pack_half2(hlo[0], hlo[1]), pack_half2(hlo[2], hlo[3]), pack_half2(hv[0], hv[1]), pack_half2(hv[2], hv[3]));
                        tmem_st_wait();
                    }
                    if (b.out_units) {
                        const float os = b.out_scale;
                        b.out_units[((((size_t)tileB * T + t_cur) * 2 + rank) * KG + 4 * cl + s) * ROWS + row_l] =
                            make_uint4(pack_half2(hlo[0] * os, hlo[1] * os), pack_half2(hlo[2] * os, hlo[3] * os),
                                       pack_half2(hv[0] * os, hv[1] * os), pack_half2(hv[2] * os, hv[3] * os));
                    }
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) arrive_leader(BAR_HB_READY + cl);
                }
                if (half == 1 && hp + 1 < hp_end) {            // the next chunk's first half
                    wait_bar(BAR_ACC_READY + ((slot + 1) & 1), (uint32_t)((((hp + 1) >> 1) & 3) >> 1));
                    fence_after_sync();
                    tmem_ld_x16(acc0 + (uint32_t)(((slot + 1) & 1) * 128), r);
                }
            }
            if (final_out) {                                   // output_layer (nn_models.py:189), last step of the last layer only
                // Issuer 0 multiplies the published h_T (TMEM) by [fp16(W_o) | W_o - fp16(W_o)]^T (one extra ring piece, N = 32) into
                // the first 32 columns of B's OTHER h buffer - it held h_{T-1}, which nothing reads any more.
                wait_bar(BAR_OUT_READY, ph_out);
                ph_out ^= 1;
                fence_after_sync();
                uint32_t o32[32];                              // columns 0..15: h_T x fp16(W_o)^T, 16..31: h_T x (W_o - fp16(W_o))^T
                tmem_ld_x32(tmem + t_lane + HB_COL + (uint32_t)((T & 1) ^ 1) * H_COLS, o32);
                tmem_ld_wait();
                const int row = (tileB * 2 + (int)rank) * ROWS + row_l;
                if (row < b.rows) {
                    const int e = row / b.n, smp = row - e * b.n;
                    const int bb = e / b.nF, fb = stream_frame0(b.stream_frames, b.frame0, bb), f = fb + e % b.nF;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {              // this thread: outputs o = s, s+4, ... of its row (O <= 16)
                        const int o = s + 4 * i;
                        if (o < b.O && fb >= 0) {              // (an inactive stream keeps its prediction ring untouched)
                            const float y = __uint_as_float(tc::pick4(o32, i, s)) + __uint_as_float(tc::pick4(o32, 4 + i, s)) + __ldg(b.bo + o);
                            b.preds[((((size_t)bb * b.pred_ring + f % b.pred_ring) * b.n_out) + smp) * b.O + o] = y;
                        }
                    }
                }
                fence_before_sync();
            }
        }
    } else if (warp >= LOAD_WARP0 && warp < LOAD_WARP0 + LOAD_WARPS) {
        TCW_REG_DEC(REGS_LOAD);
        // =================================== loader warps: keep-bit words of item i, then x1 of item i =======================
        // Two threads per row, 8 k-groups each.  Per item: first the keep-bit words of gap B (what the epilogue needs at the top of
        // wave i - drawn here a full wave ahead, into buffer i % 3), then the loads and Philox draws of x1; item i's x1 tile may be
        // written once the x-parts of item i-1 have retired (X1_DONE).
        const int row_l = (tid - LOAD_WARP0 * 32) & (ROWS - 1);
        const int j0 = ((tid - LOAD_WARP0 * 32) >> 7) * (KG / 2);  // this thread's k-groups: j0 .. j0 + 7
        constexpr int BK = 2;                                  // k-groups per batch of loads (register budget of the loader warps)
        int t = -1, tile = cluster_id - n_clusters;
        int e = 0, smp = 0, bidx = 0, f = 0;
        bool valid = false;
        uint32_t mbuf = NMASK - 1;
        for (int i = 0; i < NI; ++i) {
            if (++t == T) t = 0;
            if (t == 0) {
                tile += n_clusters;
                const int row = (tile * 2 + (int)rank) * ROWS + row_l;
                valid = row < a.rows;
                e = valid ? row / a.n : 0; smp = valid ? row - e * a.n : 0;
                bidx = e / a.nF; f = stream_frame0(a.stream_frames, a.frame0, bidx) + e % a.nF;
            }
            const uint32_t stream = a.stream_id0 + (uint32_t)bidx;
            // keep-bit words of gap B (between layers A and B) for item i.  Buffer i % 3 was last read by the epilogue during A's
            // passes of item i-3, in wave i-3; this thread has seen X1_DONE(i-2), which fires in wave i-2.
            if (++mbuf == NMASK) mbuf = 0;
            uint32_t* mdst = sMask + (size_t)mbuf * (KG / 2) * ROWS + row_l;
#pragma unroll 2
            for (int m = j0 / 2; m < j0 / 2 + KG / 4; ++m) {
                uint32_t wbits = 0xF0F0F0F0u;
                if (b.mask_mode == APE_MASK_PHILOX) {
                    const uint32_t w0 = keep_word(philox_keep_flags_rk(b.rk, stream, (uint32_t)f, (uint32_t)smp, (uint32_t)b.gap, (uint32_t)t, (uint32_t)(2 * m), b.keep_thr16));
                    const uint32_t w1 = keep_word(philox_keep_flags_rk(b.rk, stream, (uint32_t)f, (uint32_t)smp, (uint32_t)b.gap, (uint32_t)t, (uint32_t)(2 * m + 1), b.keep_thr16));
                    wbits = w0 | (w1 >> 2);
                } else if (b.mask_mode == APE_MASK_INJECTED) {
                    wbits = 0;
                    if (valid) {
                        const uint4 mm = __ldg(reinterpret_cast<const uint4*>(
                            b.masks + ((((size_t)e * b.n_gaps + b.gap) * T + t) * b.n + smp) * H + m * 16));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            wbits |= ((mm.x >> (8 * k)) & 0xFFu) ? (0x80u << (8 * k)) : 0u;
                            wbits |= ((mm.y >> (8 * k)) & 0xFFu) ? (0x40u << (8 * k)) : 0u;
                            wbits |= ((mm.z >> (8 * k)) & 0xFFu) ? (0x20u << (8 * k)) : 0u;
                            wbits |= ((mm.w >> (8 * k)) & 0xFFu) ? (0x10u << (8 * k)) : 0u;
                        }
                    }
                }
                mdst[(size_t)m * ROWS] = wbits;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[BAR_MASK_READY + mbuf]);

            const uint4* src = reinterpret_cast<const uint4*>(a.in);
            if (a.in_mode == tc::IN_UNITS) src += ((((size_t)tile * T + t) * 2 + rank) * KG) * ROWS + row_l;
            else src += ((((size_t)(e >> (a.in_rpc_shift + 1)) * T + t) * 2 + ((e >> a.in_rpc_shift) & 1)) * KG) * ROWS +
                        (e & ((1 << a.in_rpc_shift) - 1));
#pragma unroll 1
            for (int b0 = j0; b0 < j0 + KG / 2; b0 += BK) {
                uint4 pre[BK];
#pragma unroll
                for (int jj = 0; jj < BK; ++jj) pre[jj] = valid ? __ldg(src + (size_t)(b0 + jj) * ROWS) : make_uint4(0, 0, 0, 0);
                if (a.mask_mode == APE_MASK_PHILOX) {
#pragma unroll
                    for (int jj = 0; jj < BK; ++jj) {
                        const uint4 m = APE_PHILOX_DRAW(a, stream, (uint32_t)f, (uint32_t)smp, (uint32_t)a.gap, (uint32_t)t,
                                                             (uint32_t)(b0 + jj), a.keep_thr16);
                        pre[jj].x &= m.x; pre[jj].y &= m.y; pre[jj].z &= m.z; pre[jj].w &= m.w;
                    }
                } else if (a.mask_mode == APE_MASK_INJECTED && valid) {
#pragma unroll
                    for (int jj = 0; jj < BK; ++jj) {
                        const uint2 m = __ldg(reinterpret_cast<const uint2*>(
                            a.masks + ((((size_t)e * a.n_gaps + a.gap) * T + t) * a.n + smp) * H + (b0 + jj) * 8));
                        pre[jj].x &= ((m.x & 0xFFu) ? 0xFFFFu : 0u) | ((m.x & 0xFF00u) ? 0xFFFF0000u : 0u);
                        pre[jj].y &= ((m.x & 0xFF0000u) ? 0xFFFFu : 0u) | ((m.x & 0xFF000000u) ? 0xFFFF0000u : 0u);
                        pre[jj].z &= ((m.y & 0xFFu) ? 0xFFFFu : 0u) | ((m.y & 0xFF00u) ? 0xFFFF0000u : 0u);
                        pre[jj].w &= ((m.y & 0xFF0000u) ? 0xFFFFu : 0u) | ((m.y & 0xFF000000u) ? 0xFFFF0000u : 0u);
                    }
                }
                if (b0 == j0 && i >= 1) mbar_wait_bo<APE_TCW_BO_LOAD>(&bars[BAR_X1_DONE], ((uint32_t)(i - 1)) & 1u);
#pragma unroll
                for (int jj = 0; jj < BK; ++jj) *reinterpret_cast<uint4*>(sX1 + unit_offset(ROWS, row_l, b0 + jj)) = pre[jj];
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&bars[BAR_X1_READY], rank);
        }
    } else {
      TCW_REG_DEC(REGS_MMA);            // (issuers, ring producer and the idle warp that pads their warpgroup)
      if (warp >= MMA_WARP && warp < MMA_WARP + N_ISSUERS) {
        if (rank == 0) {
            // =============================== MMA issuers (leader CTA; the whole warp runs, one elected lane issues) =====
            // Issuer k owns accumulator slot k, i.e. chunks 1, 3 (k = 1) or 0, 2 (k = 0) of every layer group.  Both walk the whole
            // piece sequence (ring position and piece number advance with every piece) and act on their own chunks only.
            const uint32_t my_slot = (uint32_t)(warp - MMA_WARP);
            const uint32_t idesc = make_idesc_f16(256, 128), idesc_out = make_idesc_f16(256, OUT_N);
            const uint64_t dX1 = make_desc(smem_u32(sX1), LBO_A, SBO), dX2 = make_desc(smem_u32(sX2), LBO_A, SBO);
            const uint64_t dW = make_desc(smem_u32(sW), LBO_B, SBO), dOnes = make_desc(smem_u32(sOnes), LBO_A, SBO);
            const uint32_t bar_full = smem_u32(&bars[BAR_W_FULL]), bar_empty = smem_u32(&bars[BAR_W_EMPTY]);
            const uint32_t d_tmem = tmem + my_slot * 128u;
            uint32_t wslot = 0, gpiece = 0, uses = 0;
            auto ring_next = [&]() { wslot = wslot + 1 == NP ? 0 : wslot + 1; ++gpiece; };
            auto wait_full = [&]() {
                uint32_t spins = 0;
                while (!mbar_try_wait_addr(bar_full + (gpiece & (NFULL - 1)) * 8, (gpiece / NFULL) & 1)) {
                    if (++spins > MBAR_WD_SPINS) __trap();
                    if (APE_TCW_BO_MMA > 0) __nanosleep(APE_TCW_BO_MMA);
                }
            };
            // one chunk of one layer: x-part from the shared-memory tile `dx`, recurrent part (t > 0) from the TMEM buffer `h_tmem`
            auto chunk = [&](int cl, bool mine, int t, uint64_t dx, uint32_t h_tmem, bool x1_layer) {
                if (!mine) { ring_next(); if (t > 0) ring_next(); return; }
                if (uses >= 1) mbar_wait_bo<APE_TCW_BO_MMA>(&bars[BAR_SLOT_FREE + my_slot], (uses & 1) ^ 1);   // the epilogue drained the previous occupant
                ++uses;
                wait_full();
                fence_after_sync();
                if (elect_one()) {
                    const uint64_t bd = dW + wslot * (SLOT_BYTES >> 4);
                    mma_f16<2>(d_tmem, dx, bd, idesc, 0u);
#pragma unroll
                    for (uint32_t m = 1; m < KG / 2; ++m) mma_f16<2>(d_tmem, dx + m * (2 * LBO_A >> 4), bd + m * (2 * LBO_B >> 4), idesc, 1u);
                    mma_f16<2>(d_tmem, dOnes, bd + (KG / 2) * (2 * LBO_B >> 4), idesc, 1u);      // + bias (ones x the piece's bias rows)
                    commit_pair_addr(bar_empty + wslot * 8, 0x3);          // both CTAs' producers may refill this slot
                    if (t == 0) commit_pair(&bars[BAR_ACC_READY + my_slot], 0x3);
                    if (x1_layer && cl >= NCHL - 2) commit_pair(&bars[BAR_X1_DONE], 0x3);
                }
                __syncwarp();
                ring_next();
                if (t > 0) {
                    wait_full();
                    fence_after_sync();
                    if (elect_one()) {
                        const uint64_t bd = dW + wslot * (SLOT_BYTES >> 4);
#pragma unroll
                        for (uint32_t m = 0; m < KG / 2; ++m) mma_f16_ts<2>(d_tmem, h_tmem + 8 * m, bd + m * (2 * LBO_B >> 4), idesc, 1u);
                        commit_pair_addr(bar_empty + wslot * 8, 0x3);
                        commit_pair(&bars[BAR_ACC_READY + my_slot], 0x3);
                    }
                    __syncwarp();
                    ring_next();
                }
            };
            // output layer of a finished tile: h_T (TMEM, B's buffer T & 1) x [fp16(W_o) | W_o - fp16(W_o)]^T into B's other h buffer
            auto out_piece = [&](uint32_t item_parity) {
                if (my_slot == 0) {
                    mbar_wait_bo<APE_TCW_BO_MMA>(&bars[BAR_HB_READY + NCHL - 1], item_parity);     // all of h_T (slices are published in order)
                    wait_full();
                    fence_after_sync();
                    if (elect_one()) {
                        const uint64_t bd = make_desc(smem_u32(sW) + wslot * SLOT_BYTES, (OUT_N / 2) * 16, SBO);
                        const uint32_t at = tmem + HB_COL + (uint32_t)(T & 1) * H_COLS, dt = tmem + HB_COL + (uint32_t)((T & 1) ^ 1) * H_COLS;
#pragma unroll
                        for (uint32_t m = 0; m < KG / 2; ++m) mma_f16_ts<2>(dt, at + 8 * m, bd + m * (2 * (OUT_N / 2) * 16 >> 4), idesc_out, m > 0 ? 1u : 0u);
                        commit_pair_addr(bar_empty + wslot * 8, 0x3);
                        commit_pair(&bars[BAR_OUT_READY], 0x3);
                    }
                    __syncwarp();
                }
                ring_next();
            };
            int tA = 0, tB = -1;
            for (int w = 0; w <= NI; ++w) {
                const bool A_on = w < NI, B_on = w >= 1;
                if (w > 0) { if (++tA == T) tA = 0; }
                if (B_on) { if (++tB == T) tB = 0; }
                // h_A of item w-1 (A's recurrent operand, and B's input x2 of item w-1) is complete.  Waited for here, at the top of the
                // wave, by both issuers: the barrier cannot run a second phase ahead, because item w's chunks are issued below.
                if (w >= 1) mbar_wait_bo<APE_TCW_BO_MMA>(&bars[BAR_HA_READY + NCHL - 1], ((uint32_t)(w - 1)) & 1u);
                if (A_on) {
                    mbar_wait_bo<APE_TCW_BO_MMA>(&bars[BAR_X1_READY], (uint32_t)w & 1u);
                    fence_after_sync();
                    const uint32_t hA = tmem + HA_COL + (uint32_t)(tA & 1) * H_COLS;
#pragma unroll
                    for (int cl = 0; cl < NCHL; ++cl) {
                        if (cl == 2 && has_out && w >= 2 && tB == 0) out_piece(((uint32_t)(w - 2)) & 1u);   // the previous tile's output layer
                        chunk(cl, (uint32_t)(cl & 1) == my_slot, tA, dX1, hA, true);
                    }
                }
                if (B_on) {
                    if (w >= 2) mbar_wait_bo<APE_TCW_BO_MMA>(&bars[BAR_HB_READY + NCHL - 1], ((uint32_t)(w - 2)) & 1u);   // h_B of item w-2
                    fence_after_sync();
                    const uint64_t dx = dX2 + (((uint32_t)(w - 1)) & 1u) * (A_BYTES >> 4);
                    const uint32_t hB = tmem + HB_COL + (uint32_t)(tB & 1) * H_COLS;
#pragma unroll
                    for (int cl = 0; cl < NCHL; ++cl) chunk(cl, (uint32_t)(cl & 1) == my_slot, tB, dx, hB, false);
                }
            }
            if (has_out) out_piece(((uint32_t)(NI - 1)) & 1u);   // the last tile's output layer
        } else if (warp == MMA_WARP && lane == 0) {
            // =============================== peer CTA: forward "piece landed in my ring" to the leader ==================
            uint32_t gpiece = 0;
            auto forward = [&]() {
                mbar_wait_bo<APE_TCW_BO_RING>(&bars[BAR_W_FULL + (gpiece & (NFULL - 1))], (gpiece / NFULL) & 1);
                mbar_arrive_remote(&bars[BAR_W_FULL + (gpiece & (NFULL - 1))], 0);
                ++gpiece;
            };
            int tA = 0, tB = -1;
            for (int w = 0; w <= NI; ++w) {
                const bool A_on = w < NI, B_on = w >= 1;
                if (w > 0) { if (++tA == T) tA = 0; }
                if (B_on) { if (++tB == T) tB = 0; }
                if (A_on) for (int cl = 0; cl < NCHL; ++cl) {
                    if (cl == 2 && has_out && w >= 2 && tB == 0) forward();
                    forward();
                    if (tA > 0) forward();
                }
                if (B_on) for (int cl = 0; cl < NCHL; ++cl) { forward(); if (tB > 0) forward(); }
            }
            if (has_out) forward();
        }
      } else if (warp == TMA_WARP && lane == 0) {
        // =================================== weight-ring producer (one lane per CTA) =====================================
        const uint8_t* WA = a.Ww + (size_t)rank * LAYER_BYTES;  // this CTA's half of each layer's pieces
        const uint8_t* WB = b.Ww + (size_t)rank * LAYER_BYTES;
        uint32_t wslot = 0, wphase = 0, gpiece = 0;
        bool wrapped = false;
        auto put = [&](const uint8_t* src, uint32_t bytes) {
            uint64_t* full = &bars[BAR_W_FULL + (gpiece & (NFULL - 1))];
            if (wrapped) mbar_wait_bo<APE_TCW_BO_RING>(&bars[BAR_W_EMPTY + wslot], wphase ^ 1);   // previous occupant consumed
            mbar_arrive_expect_tx(full, bytes);
            bulk_g2s(sW + wslot * SLOT_BYTES, src, bytes, full);
            if (++wslot == NP) { wslot = 0; wphase ^= 1; wrapped = true; }
            ++gpiece;
        };
        int tA = 0, tB = -1;
        for (int w = 0; w <= NI; ++w) {
            const bool A_on = w < NI, B_on = w >= 1;
            if (w > 0) { if (++tA == T) tA = 0; }
            if (B_on) { if (++tB == T) tB = 0; }
            if (A_on) for (int cl = 0; cl < NCHL; ++cl) {
                if (cl == 2 && has_out && w >= 2 && tB == 0) put(b.Wo16 + (size_t)rank * OUT_BYTES, OUT_BYTES);
                put(WA + (size_t)cl * CHUNK_BYTES, PIECE_X_BYTES);
                if (tA > 0) put(WA + (size_t)cl * CHUNK_BYTES + PIECE_X_BYTES, PIECE_H_BYTES);
            }
            if (B_on) for (int cl = 0; cl < NCHL; ++cl) {
                put(WB + (size_t)cl * CHUNK_BYTES, PIECE_X_BYTES);
                if (tB > 0) put(WB + (size_t)cl * CHUNK_BYTES + PIECE_X_BYTES, PIECE_H_BYTES);
            }
        }
        if (has_out) put(b.Wo16 + (size_t)rank * OUT_BYTES, OUT_BYTES);
      }
    }
    __syncwarp();
    fence_before_sync();
    cluster_sync();
    tc::timeline_stamp(a.timeline, 2);
    if (warp == MMA_WARP) tmem_dealloc<2>(tmem, TMEM_COLS);
}

bool supported(int Hh, int T, int O) { return Hh == H && T >= 2 && O <= (int)OUT_N / 2; }

size_t layer_bytes(int Hh) { return Hh == H ? (size_t)2 * LAYER_BYTES : 0; }

size_t scratch_bytes(int Hh, int sm_count) { return Hh == H ? (size_t)sm_count * CSTATE_FLOATS * sizeof(float) : 0; }

// layers A = a, B = b of one launch; a.cstate: scratch_bytes() bytes
int launch_pair(const TcLayerArgs& a, const TcLayerArgs& b, int sm_count, cudaStream_t st) {
    if (a.kgx != KG || b.kgx != KG || a.rpc != ROWS || !a.cstate || a.T != b.T || a.T < 2 || !a.Ww || !b.Ww) return APE_ERR_UNSUPPORTED;
    if (a.in_mode != tc::IN_UNITS && a.in_mode != tc::IN_SHARED_UNITS) return APE_ERR_UNSUPPORTED;
    if (b.preds && (!b.Wo16 || b.O > (int)OUT_N / 2)) return APE_ERR_UNSUPPORTED;
    APE_CUDA_TRY(cudaFuncSetAttribute(lstm_pair_tcw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    int clusters = sm_count / 2;
    if (clusters > a.n_pair_tiles) clusters = a.n_pair_tiles;
    lstm_pair_tcw_kernel<<<2 * clusters, THREADS, SMEM, st>>>(a, b);
    return check_launch();
}

}  // namespace tcw
}  // namespace ape
