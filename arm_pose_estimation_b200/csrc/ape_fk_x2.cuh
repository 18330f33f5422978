// Stage 3 row math in PACKED fp32 pairs (device only).  Blackwell issues two fp32 operations per instruction on a 64-bit register pair
// (PTX add / sub / mul / fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2; operand negation, an immediate and a scalar broadcast are folded into
// the instruction by ptxas).  The stage-3 kernel is ISSUE-bound (ncu: issue slots 70 - 80 %, FMA pipe 41 %, DRAM 28 %), and a row holds two
// rotations of identical structure - the lower-arm and the upper-arm 6D columns - so the pair (lo = lower arm, hi = upper arm) goes
// through Gram-Schmidt, matrix -> quaternion, the bone rotation and the sign-aligned sums as ONE instruction stream: half the issue
// slots for the same arithmetic (each half is the IEEE operation the scalar code of ape_fk.cuh does; only where a product and a sum fuse
// into an FMA may the last bit differ from it).
#pragma once
#include "ape_f32x2.cuh"
#include "ape_fk.cuh"

namespace ape {

__device__ __forceinline__ F2 rsq2(F2 a) { return pk(inv_sqrt(lo(a)), inv_sqrt(hi(a))); }      // two MUFU.RSQ
__device__ __forceinline__ F2 shfl_xor2(F2 a, int o) { return pk(__shfl_xor_sync(0xffffffffu, lo(a), o), __shfl_xor_sync(0xffffffffu, hi(a), o)); }
__device__ __forceinline__ F2 shfl_idx2(F2 a, int l) { return pk(__shfl_sync(0xffffffffu, lo(a), l), __shfl_sync(0xffffffffu, hi(a), l)); }

struct Quat2 { F2 w, x, y, z; };                                         // lo: lower arm, hi: upper arm
__device__ __forceinline__ Quat<float> q_lo(const Quat2& q) { return {lo(q.w), lo(q.x), lo(q.y), lo(q.z)}; }
__device__ __forceinline__ Quat<float> q_hi(const Quat2& q) { return {hi(q.w), hi(q.x), hi(q.y), hi(q.z)}; }
__device__ __forceinline__ Quat2 q_pack(const Quat<float>& l, const Quat<float>& u) { return {pk(l.w, u.w), pk(l.x, u.x), pk(l.y, u.y), pk(l.z, u.z)}; }

// one half of the branch-free row selection of six_to_quat (ape_fk.cuh): the row of 4 q q^T with the largest diagonal entry
__device__ __forceinline__ Quat<float> pick_row(float fw, float fx, float fy, float fz, float wx, float wy, float wz, float xy, float xz, float yz) {
    const bool wx_w = fw >= fx, yz_y = fy >= fz;
    const float m01 = wx_w ? fw : fx, m23 = yz_y ? fy : fz;
    const bool first = m01 >= m23;
    const float aw = wx_w ? fw : wx, ax = wx_w ? wx : fx, ay = wx_w ? wy : xy, az = wx_w ? wz : xz;
    const float bw = yz_y ? wy : wz, bx = yz_y ? xy : xz, by = yz_y ? fy : yz, bz = yz_y ? yz : fz;
    return {first ? aw : bw, first ? ax : bx, first ? ay : by, first ? az : bz};
}

// six_to_quat (ape_fk.cuh; transformations.py:602-637 + :521-545) for the lower-arm columns cl[0..5] and the upper-arm columns cu[0..5] at once
__device__ __forceinline__ Quat2 six_to_quat_x2(const float* cl, const float* cu, bool& bad) {
    const F2 a1x = pk(cl[0], cu[0]), a1y = pk(cl[2], cu[2]), a1z = pk(cl[4], cu[4]);
    const F2 a2x = pk(cl[1], cu[1]), a2y = pk(cl[3], cu[3]), a2z = pk(cl[5], cu[5]);
    const F2 s1 = fma2(a1x, a1x, fma2(a1y, a1y, a1z * a1z)), i1 = rsq2(s1);
    const F2 b1x = a1x * i1, b1y = a1y * i1, b1z = a1z * i1;
    const F2 d = fma2(b1x, a2x, fma2(b1y, a2y, b1z * a2z));
    const F2 ux = fma2(-d, b1x, a2x), uy = fma2(-d, b1y, a2y), uz = fma2(-d, b1z, a2z);
    const F2 s2 = fma2(ux, ux, fma2(uy, uy, uz * uz)), i2 = rsq2(s2);
    const F2 b2x = ux * i2, b2y = uy * i2, b2z = uz * i2;
    const F2 b3x = fma2(b1y, b2z, -(b1z * b2y)), b3y = fma2(b1z, b2x, -(b1x * b2z)), b3z = fma2(b1x, b2y, -(b1y * b2x));
    // a zero / denormal / non-finite squared column norm (the same bounds as the scalar code: (1e-36, inf), NaN fails both)
    const float t = 1e-36f, inf = __int_as_float(0x7f800000);
    if (!(lo(s1) > t && hi(s1) > t && lo(s2) > t && hi(s2) > t && lo(s1) < inf && hi(s1) < inf && lo(s2) < inf && hi(s2) < inf)) bad = true;
    // R = [b1 b2 b3] as columns: r00 = b1x, r01 = b2x, r02 = b3x, r10 = b1y, r11 = b2y, r12 = b3y, r20 = b1z, r21 = b2z, r22 = b3z
    const F2 one = splat(1.0f);
    const F2 tp = b1x + one, tm = one - b1x, vp = b2y + b3z, vm = b2y - b3z;
    const F2 fw = tp + vp, fx = tp - vp, fy = tm + vm, fz = tm - vm;
    const F2 wx = b2z - b3y, wy = b3x - b1z, wz = b1y - b2x, xy = b2x + b1y, xz = b3x + b1z, yz = b3y + b2z;
    const Quat<float> ql = pick_row(lo(fw), lo(fx), lo(fy), lo(fz), lo(wx), lo(wy), lo(wz), lo(xy), lo(xz), lo(yz));
    const Quat<float> qu = pick_row(hi(fw), hi(fx), hi(fy), hi(fz), hi(wx), hi(wy), hi(wz), hi(xy), hi(xz), hi(yz));
    const Quat2 q = q_pack(ql, qu);
    const F2 inv = rsq2(fma2(q.w, q.w, fma2(q.x, q.x, fma2(q.y, q.y, q.z * q.z))));
    const F2 sg = pk(ql.w < 0.0f ? -lo(inv) : lo(inv), qu.w < 0.0f ? -hi(inv) : hi(inv));      // w >= 0 (:543-544)
    return {q.w * sg, q.x * sg, q.y * sg, q.z * sg};
}

// qrot_x (ape_common.cuh) of both quaternions: len = (lower-arm length, upper-arm length), len2 = 2 len
__device__ __forceinline__ void qrot_x2(const Quat2& q, F2 len, F2 len2, F2& vx, F2& vy, F2& vz) {
    vx = len * fma2(-q.z, q.z, fma2(-q.y, q.y, fma2(q.w, q.w, q.x * q.x)));
    vy = len2 * fma2(q.x, q.y, q.w * q.z);
    vz = len2 * fma2(q.x, q.z, -(q.w * q.y));
}

}  // namespace ape
