// Stage 3 kernel: de-normalise the network targets, 6D rotation -> quaternion, forward kinematics through
// shoulder -> elbow -> hand for every MC / smoothing row, then the per-estimate reduction: sign-aligned
// quaternion average (transformations.py:32-51), FK again from the averaged quaternions (compose_msg.py:59-61,
// :92-93), population std of the per-row hand / elbow positions.  One HALF-WARP per estimate (two estimates per warp: 100 rows are
// 6.25 passes of 16 lanes instead of 3.1 passes of 32, and every reduction instruction serves two estimates), lanes stride over the
// S = smooth * n_samples rows (consecutive lanes read consecutive rows), reductions by shuffles inside the half-warp, results staged
// through shared memory so the 25-float message leaves coalesced.
#include "ape_fk.cuh"

namespace ape {

// Resident CTAs per SM -> registers per thread.  With the next row prefetched the kernel wants 80 registers (O = 12) / 96 (O = 14, 20):
// at 64 it spills 45 values per pass.  Measured on B200, 32 768 estimates x 100 rows: O = 12: 8 / 6 / 5 CTAs = 0.080 / 0.068 / 0.070 ms,
// O = 14: 0.109 / 0.089 / 0.076 ms.
#ifndef APE_FK_MIN_BLOCKS
#define APE_FK_MIN_BLOCKS(TARGET) ((TARGET) == APE_TARGET_ORI_CAL_LARM_UARM ? 6 : 5)
#endif
#ifndef APE_FK_WARPS
#define APE_FK_WARPS 4
#endif
constexpr int FK_WARPS_PER_CTA = APE_FK_WARPS;

struct FkArgs {
    const float* preds;
    int pred_ring;
    const float* yy_m;
    const float* yy_s;
    const float* body9;
    int O, B, nF, frame0, n, smooth;
    const int32_t* stream_frames;   // per-stream frame counters (null: frame0 for all); < 0: skip the stream
    float* msg;
    float* samples;
    float* stdev;
    float* est_rows;
    int32_t* status;
    const float* est_in;      // FROM_EST: [E][S][W] rows of arm_pose_from_nn_targets instead of network targets
};

#ifndef APE_FK_LANES
#define APE_FK_LANES 16
#endif
constexpr int FK_LANES = APE_FK_LANES;             // lanes per estimate (fixes the summation order: never chosen by batch size)
constexpr int FK_SUB = 32 / FK_LANES;              // estimates per warp
static_assert(FK_LANES == 8 || FK_LANES == 16, "lanes per estimate");

__device__ __forceinline__ float warp_sum(float v, unsigned m) {          // over the 16 lanes of this half-warp (mask m)
#pragma unroll
    for (int o = FK_LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
}

template <typename T> __device__ __forceinline__ Quat<T> bcast0(const Quat<T>& q, unsigned m, int l0) {   // from the half-warp's lane 0
    return {__shfl_sync(m, q.w, l0), __shfl_sync(m, q.x, l0), __shfl_sync(m, q.y, l0), __shfl_sync(m, q.z, l0)};
}

// accumulate q flipped onto the hemisphere of q0 (transformations.py:44-49)
__device__ __forceinline__ void acc_aligned(Quat<float>& s, const Quat<float>& q, const Quat<float>& q0) {
    const float d = q.w * q0.w + q.x * q0.x + q.y * q0.y + q.z * q0.z;
    const float sg = d < 0.0f ? -1.0f : 1.0f;
    s.w += sg * q.w; s.x += sg * q.x; s.y += sg * q.y; s.z += sg * q.z;
}

__device__ __forceinline__ Quat<float> warp_sum_normalised(Quat<float> s, unsigned m) {
    s.w = warp_sum(s.w, m); s.x = warp_sum(s.x, m); s.y = warp_sum(s.y, m); s.z = warp_sum(s.z, m);
    const float inv = inv_sqrt(s.w * s.w + s.x * s.x + s.y * s.y + s.z * s.z);
    return {s.w * inv, s.x * inv, s.y * inv, s.z * inv};
}

template <int TARGET, bool FROM_EST>
__global__ void __launch_bounds__(FK_WARPS_PER_CTA * 32, APE_FK_MIN_BLOCKS(TARGET)) fk_reduce_kernel(FkArgs a) {
    constexpr int O = TARGET == APE_TARGET_ORI_CAL_LARM_UARM ? 12 : (TARGET == APE_TARGET_ORI_CAL_LARM_UARM_HIPS ? 14 : 20);
    constexpr int W = TARGET == APE_TARGET_ORI_CAL_LARM_UARM ? 14 : 21;
    __shared__ float s_msg[FK_WARPS_PER_CTA * FK_SUB][32];
    __shared__ float s_m[O], s_s[O];

    if (threadIdx.x < O) {
        s_m[threadIdx.x] = a.yy_m ? a.yy_m[threadIdx.x] : 0.0f;
        s_s[threadIdx.x] = a.yy_s ? a.yy_s[threadIdx.x] : 1.0f;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, half = (threadIdx.x & 31) / FK_LANES, lane = threadIdx.x & (FK_LANES - 1);   // lane within the sub-warp
    const unsigned hm = ((1u << FK_LANES) - 1u) << (FK_LANES * half);   // this sub-warp's lanes: every shuffle / ballot below stays inside it
    const int l0 = FK_LANES * half;                    // its first lane
    const int e = (blockIdx.x * FK_WARPS_PER_CTA + warp) * FK_SUB + half;
    const int E = a.B * a.nF;
    if (e >= E) return;
    const int b = e / a.nF;
    const int fb = a.stream_frames ? a.stream_frames[b] : a.frame0;
    if (fb < 0) return;                            // the stream has no new frame in this call: its outputs stay untouched
    const int f = fb + e % a.nF;
    const int S = a.smooth * a.n;

    Body<float> body;
    body.larm_vec = {a.body9[0], a.body9[1], a.body9[2]};
    body.uarm_vec = {a.body9[3], a.body9[4], a.body9[5]};
    body.uarm_orig = {a.body9[6], a.body9[7], a.body9[8]};
    body.bones_along_x = bones_are_along_x(body);      // (the default skeleton: a uniform branch)

    Quat<float> q0l{}, q0u{}, q0h{}, sl{0, 0, 0, 0}, su{0, 0, 0, 0}, sh{0, 0, 0, 0};
    float piv[6] = {0, 0, 0, 0, 0, 0}, d1[6] = {0, 0, 0, 0, 0, 0}, d2[6] = {0, 0, 0, 0, 0, 0};
    float psum[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // ORI_POS target: plain means of hand / elbow / shoulder
    float sh0[3] = {0, 0, 0};                      // shoulder of row 0 (S == 1: the message copies row 0)
    bool bad = false;

    int rw = lane / a.n, rs = lane - rw * a.n;        // this lane's row i = rw * n + rs (window frame, MC sample), advanced by 16 per pass
    // first prediction row of window frame rw (frames clamp to frame 0, estimator.py:114-115; the ring slot costs an integer
    // division, so it is re-derived only when the lane moves on to the next window frame - not per row)
    auto frame_rows = [&](int w) {
        int fw = f - a.smooth + 1 + w;
        fw = fw < 0 ? 0 : fw;
        return a.preds + ((size_t)b * a.pred_ring + (size_t)(fw % a.pred_ring)) * a.n * O;
    };
    const float* wrows = FROM_EST ? nullptr : frame_rows(rw);
    // The prediction row of the NEXT pass is requested before this pass's arithmetic (a lane reads 48 - 80 bytes per pass: with the
    // load issued where it is consumed, a third of the kernel's warp-cycles were spent waiting for it)
    float pre[O];
    auto request = [&](const float* row) {
        if (O % 4 == 0) {                                        // 48 / 80-byte rows: 16-byte loads
#pragma unroll
            for (int j = 0; j < O / 4; ++j) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(row) + j);
                pre[4 * j] = v.x; pre[4 * j + 1] = v.y; pre[4 * j + 2] = v.z; pre[4 * j + 3] = v.w;
            }
        } else {                                                 // 56-byte rows: 8-byte loads
#pragma unroll
            for (int j = 0; j < O / 2; ++j) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(row) + j);
                pre[2 * j] = v.x; pre[2 * j + 1] = v.y;
            }
        }
    };
    if (!FROM_EST && lane < S) request(wrows + rs * O);
    float2* smp = a.samples ? reinterpret_cast<float2*>(a.samples + ((size_t)e * S + lane) * 6) : nullptr;   // this lane's row of the pass
    for (int i0 = 0; i0 < S; i0 += FK_LANES) {
        const int i = i0 + lane;
        const bool live = i < S;
        RowPose<float> r;
        r.larm = r.uarm = r.hips = {1.0f, 0.0f, 0.0f, 0.0f};
        r.hand = r.elbow = r.shoulder = {0.0f, 0.0f, 0.0f};
        if (live && FROM_EST) {                                      // compose_msg.py entry: rows already hold quats + origins
            const float* src = a.est_in + ((size_t)e * S + i) * W;
            int k = 0;
            r.hand = {src[0], src[1], src[2]};
            r.elbow = {src[3], src[4], src[5]};
            k = 6;
            if (W == 21) { r.shoulder = {src[6], src[7], src[8]}; k = 9; } else { r.shoulder = body.uarm_orig; }
            r.larm = {src[k], src[k + 1], src[k + 2], src[k + 3]};
            r.uarm = {src[k + 4], src[k + 5], src[k + 6], src[k + 7]};
            if (W == 21) r.hips = {src[k + 8], src[k + 9], src[k + 10], src[k + 11]};
        } else if (live) {
            float p[O];
#pragma unroll
            for (int j = 0; j < O; ++j) p[j] = pre[j];
            // advance to the next pass's row (no integer division per row) and request it
            rs += FK_LANES;
            if (rs >= a.n) {
                do { rs -= a.n; ++rw; } while (rs >= a.n);
                if (rw < a.smooth) wrows = frame_rows(rw);
            }
            if (i + FK_LANES < S) request(wrows + rs * O);
#pragma unroll
            for (int j = 0; j < O; ++j) p[j] = fmaf(p[j], s_s[j], s_m[j]);     // estimator.py:108-109
            bool rb = false;
            r = row_pose<float>(TARGET, p, body, rb);
            bad |= rb;
        }
        if (i0 == 0) {                                               // row 0 anchors the sign alignment and the std pivot
            q0l = bcast0(r.larm, hm, l0); q0u = bcast0(r.uarm, hm, l0); q0h = bcast0(r.hips, hm, l0);
            const float pv[6] = {r.hand.x, r.hand.y, r.hand.z, r.elbow.x, r.elbow.y, r.elbow.z};
#pragma unroll
            for (int j = 0; j < 6; ++j) piv[j] = __shfl_sync(hm, pv[j], l0);
            sh0[0] = __shfl_sync(hm, r.shoulder.x, l0);
            sh0[1] = __shfl_sync(hm, r.shoulder.y, l0);
            sh0[2] = __shfl_sync(hm, r.shoulder.z, l0);
        }
        if (live) {
            acc_aligned(sl, r.larm, q0l);
            acc_aligned(su, r.uarm, q0u);
            if (TARGET != APE_TARGET_ORI_CAL_LARM_UARM) acc_aligned(sh, r.hips, q0h);
            const float pv[6] = {r.hand.x, r.hand.y, r.hand.z, r.elbow.x, r.elbow.y, r.elbow.z};
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const float d = pv[j] - piv[j];
                d1[j] += d;
                d2[j] = fmaf(d, d, d2[j]);
            }
            if (TARGET == APE_TARGET_ORI_POS_CAL_LARM_UARM_HIPS) {
#pragma unroll
                for (int j = 0; j < 6; ++j) psum[j] += pv[j];
                psum[6] += r.shoulder.x; psum[7] += r.shoulder.y; psum[8] += r.shoulder.z;
            }
            if (smp) {
                smp[0] = make_float2(pv[0], pv[1]);
                smp[1] = make_float2(pv[2], pv[3]);
                smp[2] = make_float2(pv[4], pv[5]);
                smp += FK_LANES * 3;
            }
            if (a.est_rows) {
                float* dst = a.est_rows + ((size_t)e * S + i) * W;
                int k = 0;
#pragma unroll
                for (int j = 0; j < 6; ++j) dst[k++] = pv[j];
                if (W == 21) { dst[k++] = r.shoulder.x; dst[k++] = r.shoulder.y; dst[k++] = r.shoulder.z; }
                dst[k++] = r.larm.w; dst[k++] = r.larm.x; dst[k++] = r.larm.y; dst[k++] = r.larm.z;
                dst[k++] = r.uarm.w; dst[k++] = r.uarm.x; dst[k++] = r.uarm.y; dst[k++] = r.uarm.z;
                if (W == 21) { dst[k++] = r.hips.w; dst[k++] = r.hips.x; dst[k++] = r.hips.y; dst[k++] = r.hips.z; }
            }
        }
    }

    // ---- reduction over the S rows --------------------------------------------------------------------
    RowPose<float> m;
    m.larm = warp_sum_normalised(sl, hm);
    m.uarm = warp_sum_normalised(su, hm);
    m.hips = TARGET == APE_TARGET_ORI_CAL_LARM_UARM ? Quat<float>{1.0f, 0.0f, 0.0f, 0.0f} : warp_sum_normalised(sh, hm);
    if (S == 1) {                                                             // copied, not re-normalised
        m.larm = q0l; m.uarm = q0u;
        if (TARGET != APE_TARGET_ORI_CAL_LARM_UARM) m.hips = q0h;
    }
    const float invS = 1.0f / (float)S;
    if (TARGET == APE_TARGET_ORI_POS_CAL_LARM_UARM_HIPS) {
#pragma unroll
        for (int j = 0; j < 9; ++j) psum[j] = warp_sum(psum[j], hm) * invS;       // compose_msg.py:27-29
        m.hand = {psum[0], psum[1], psum[2]};
        m.elbow = {psum[3], psum[4], psum[5]};
        m.shoulder = {psum[6], psum[7], psum[8]};
    } else if (S > 1) {
        chain(TARGET, body, m);                                               // FK again from the means
    } else {                                                                  // one row: copied as is (compose_msg.py:62-66)
        m.hand = {piv[0], piv[1], piv[2]};
        m.elbow = {piv[3], piv[4], piv[5]};
        m.shoulder = {sh0[0], sh0[1], sh0[2]};
    }
    float sd[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const float m1 = warp_sum(d1[j], hm) * invS, m2 = warp_sum(d2[j], hm) * invS;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sd[j]) : "f"(fmaxf(m2 - m1 * m1, 0.0f)));
    }
    const unsigned any_bad = __ballot_sync(hm, bad) & hm;

    if (lane == 0) {
        float* o = s_msg[warp * FK_SUB + half];                                             // compose_msg.py:67-79 / :100-108
        o[0] = m.larm.w; o[1] = m.larm.x; o[2] = m.larm.y; o[3] = m.larm.z;
        o[4] = m.hand.x; o[5] = m.hand.y; o[6] = m.hand.z;
        o[7] = m.larm.w; o[8] = m.larm.x; o[9] = m.larm.y; o[10] = m.larm.z;
        o[11] = m.elbow.x; o[12] = m.elbow.y; o[13] = m.elbow.z;
        o[14] = m.uarm.w; o[15] = m.uarm.x; o[16] = m.uarm.y; o[17] = m.uarm.z;
        o[18] = m.shoulder.x; o[19] = m.shoulder.y; o[20] = m.shoulder.z;
        o[21] = m.hips.w; o[22] = m.hips.x; o[23] = m.hips.y; o[24] = m.hips.z;
        if (a.status) a.status[e] = any_bad ? 1 : 0;
    }
    __syncwarp(hm);
    for (int k = lane; k < 25; k += FK_LANES) a.msg[(size_t)e * 25 + k] = s_msg[warp * FK_SUB + half][k];
    if (a.stdev && lane < 6) {
        float v = sd[0];
#pragma unroll
        for (int j = 1; j < 6; ++j) v = lane == j ? sd[j] : v;
        a.stdev[(size_t)e * 6 + lane] = v;
    }
}

}  // namespace ape

extern "C" int ape_fk_reduce(const float* preds, int pred_ring, const float* yy_m, const float* yy_s,
                             const float* body9, int target, int O, int B, int nF, int frame0, const int32_t* stream_frames, int n_samples,
                             int smooth, float* msg, float* samples, float* stdev, float* est_rows,
                             int32_t* status, void* stream) {
    using namespace ape;
    if (!preds || !body9 || !msg || B < 0 || nF < 0 || frame0 < 0 || n_samples < 1 || smooth < 1 || pred_ring < 1)
        return APE_ERR_BAD_ARG;
    if (target < APE_TARGET_ORI_CAL_LARM_UARM || target > APE_TARGET_ORI_POS_CAL_LARM_UARM_HIPS) return APE_ERR_BAD_ARG;
    if (O != target_num_outputs(target)) return APE_ERR_BAD_ARG;
    if ((yy_m == nullptr) != (yy_s == nullptr)) return APE_ERR_BAD_ARG;
    if (nF + smooth - 1 > pred_ring) return APE_ERR_BAD_ARG;          // the smoothing window must still be in the ring
    const long long E = (long long)B * nF;
    if (E == 0) return APE_OK;
    if (E > 0x7fffffffLL || (long long)smooth * n_samples > 0x7fffffffLL) return APE_ERR_BAD_ARG;
    FkArgs a{preds, pred_ring, yy_m, yy_s, body9, O, B, nF, frame0, n_samples, smooth, stream_frames, msg, samples, stdev, est_rows, status, nullptr};
    const int grid = (int)((E + FK_SUB * FK_WARPS_PER_CTA - 1) / (FK_SUB * FK_WARPS_PER_CTA));
    cudaStream_t st = (cudaStream_t)stream;
    if (target == APE_TARGET_ORI_CAL_LARM_UARM)
        fk_reduce_kernel<APE_TARGET_ORI_CAL_LARM_UARM, false><<<grid, FK_WARPS_PER_CTA * 32, 0, st>>>(a);
    else if (target == APE_TARGET_ORI_CAL_LARM_UARM_HIPS)
        fk_reduce_kernel<APE_TARGET_ORI_CAL_LARM_UARM_HIPS, false><<<grid, FK_WARPS_PER_CTA * 32, 0, st>>>(a);
    else
        fk_reduce_kernel<APE_TARGET_ORI_POS_CAL_LARM_UARM_HIPS, false><<<grid, FK_WARPS_PER_CTA * 32, 0, st>>>(a);
    return check_launch();
}

extern "C" int ape_msg_from_est(const float* est, int W, const float* body9, int target, int E, int S, float* msg,
                                float* stdev, void* stream) {
    using namespace ape;
    if (!est || !body9 || !msg || E < 0 || S < 1) return APE_ERR_BAD_ARG;
    if (target < APE_TARGET_ORI_CAL_LARM_UARM || target > APE_TARGET_ORI_POS_CAL_LARM_UARM_HIPS) return APE_ERR_BAD_ARG;
    if (W != (target == APE_TARGET_ORI_CAL_LARM_UARM ? 14 : 21)) return APE_ERR_BAD_ARG;
    if (E == 0) return APE_OK;
    FkArgs a{nullptr, 1, nullptr, nullptr, body9, target_num_outputs(target), E, 1, 0, S, 1, nullptr, msg, nullptr, stdev, nullptr, nullptr, est};
    const int grid = (E + FK_SUB * FK_WARPS_PER_CTA - 1) / (FK_SUB * FK_WARPS_PER_CTA);
    cudaStream_t st = (cudaStream_t)stream;
    if (target == APE_TARGET_ORI_CAL_LARM_UARM)
        fk_reduce_kernel<APE_TARGET_ORI_CAL_LARM_UARM, true><<<grid, FK_WARPS_PER_CTA * 32, 0, st>>>(a);
    else if (target == APE_TARGET_ORI_CAL_LARM_UARM_HIPS)
        fk_reduce_kernel<APE_TARGET_ORI_CAL_LARM_UARM_HIPS, true><<<grid, FK_WARPS_PER_CTA * 32, 0, st>>>(a);
    else
        fk_reduce_kernel<APE_TARGET_ORI_POS_CAL_LARM_UARM_HIPS, true><<<grid, FK_WARPS_PER_CTA * 32, 0, st>>>(a);
    return check_launch();
}
