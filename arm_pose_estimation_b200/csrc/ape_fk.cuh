// Stage 3 math: 6D rotation -> quaternion, forward kinematics through shoulder -> elbow -> hand.
// Host/device, templated on the float type (float on the GPU; double in host checks).
#pragma once
#include "ape_common.cuh"

namespace ape {

// 1 / sqrt(x): every normalisation below multiplies by it (one MUFU.RSQ, <= 2 ulp, on the device instead of an IEEE square root
// plus three IEEE divisions - those were two thirds of the stage-3 kernel's instructions)
APE_HD double inv_sqrt(double x) { return 1.0 / sqrt(x); }
APE_HD float inv_sqrt(float x) {
#ifdef __CUDA_ARCH__
    // the bare MUFU.RSQ (rsqrtf() wraps it in a denormal fix-up: 5 more instructions, six times per row); arguments here are
    // squared norms of O(1) vectors - a denormal one is reported as a degenerate row (six_to_quat: `bad`)
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return 1.0f / sqrtf(x);
#endif
}

// Gram-Schmidt of a1 = (c0,c2,c4), a2 = (c1,c3,c5) (transformations.py:616-626), then rotation matrix ->
// quaternion.  The reference takes the dominant eigenvector of transforms3d's 4x4 K matrix (:521-545);
// for the orthonormal matrix produced here that equals the closed-form conversion, evaluated with the
// best-conditioned of the four branches (largest of 4w^2,4x^2,4y^2,4z^2), then w >= 0 (:543-544).
// `bad` is set when a column norm is zero / non-finite (the reference raises LinAlgError there).
template <typename F> APE_HD Quat<F> six_to_quat(const F* c, bool& bad) {
    const F a1x = c[0], a1y = c[2], a1z = c[4], a2x = c[1], a2y = c[3], a2z = c[5];
    const F s1 = a1x * a1x + a1y * a1y + a1z * a1z, i1 = inv_sqrt(s1);
    const F b1x = a1x * i1, b1y = a1y * i1, b1z = a1z * i1;
    const F d = b1x * a2x + b1y * a2y + b1z * a2z;
    const F ux = a2x - d * b1x, uy = a2y - d * b1y, uz = a2z - d * b1z;
    const F s2 = ux * ux + uy * uy + uz * uz, i2 = inv_sqrt(s2);
    const F b2x = ux * i2, b2y = uy * i2, b2z = uz * i2;
    const F b3x = b1y * b2z - b1z * b2y, b3y = b1z * b2x - b1x * b2z, b3z = b1x * b2y - b1y * b2x;
    // a zero / non-finite / absurdly long column (squared norms: the bound is 1e30 squared, infinity in float)
    const F smin = s1 < s2 ? s1 : s2, smax = s1 < s2 ? s2 : s1;
    // (a NaN norm fails the first test whichever side it took; float: a norm below 1e-18 counts as zero - its square is denormal)
    if (!(smin > (sizeof(F) == 4 ? F(1e-36) : F(0))) || !(smax < F(1e30) * F(1e30))) bad = true;
    // R = [b1 b2 b3] as columns: r_ij = row i, column j
    const F r00 = b1x, r01 = b2x, r02 = b3x, r10 = b1y, r11 = b2y, r12 = b3y, r20 = b1z, r21 = b2z, r22 = b3z;
    const F fw = F(1) + r00 + r11 + r22, fx = F(1) + r00 - r11 - r22;
    const F fy = F(1) - r00 + r11 - r22, fz = F(1) - r00 - r11 + r22;
    // The four conversion branches are the four rows of the symmetric matrix 4 q q^T (row w = (4w^2, 4wx, 4wy, 4wz), ...): each is
    // a multiple of q, and the row of the largest diagonal entry is the well-conditioned one.  Selecting that row component by
    // component and normalising ONCE is branch-free - on the GPU the four-way branch diverged inside every warp and carried its own
    // reciprocal square root - and is the same quaternion (eigh returns a unit vector, made w >= 0 by :543-544).
    const F wx = r21 - r12, wy = r02 - r20, wz = r10 - r01, xy = r01 + r10, xz = r02 + r20, yz = r12 + r21;
    const bool wx_w = fw >= fx, yz_y = fy >= fz;                 // winners of the two pairs, then of the final
    const F m01 = wx_w ? fw : fx, m23 = yz_y ? fy : fz;
    const bool first = m01 >= m23;
    // candidate rows: W = (fw, wx, wy, wz), X = (wx, fx, xy, xz), Y = (wy, xy, fy, yz), Z = (wz, xz, yz, fz)
    const F aw = wx_w ? fw : wx, ax = wx_w ? wx : fx, ay = wx_w ? wy : xy, az = wx_w ? wz : xz;
    const F bw = yz_y ? wy : wz, bx = yz_y ? xy : xz, by = yz_y ? fy : yz, bz = yz_y ? yz : fz;
    Quat<F> q = {first ? aw : bw, first ? ax : bx, first ? ay : by, first ? az : bz};
    const F inv = inv_sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    const F sg = q.w < F(0) ? -inv : inv;
    return {q.w * sg, q.x * sg, q.y * sg, q.z * sg};
}

// yaw-only quaternion from (sin, cos): euler_to_quat([0, atan2(s, c), 0]) (transformations.py:177-179)
// Evaluated without atan2 / sin / cos: with (sy, cy) = (s, c) / |(s, c)| the half angle y / 2 in [-pi/2, pi/2] has
//   cos(y/2) = sqrt((1 + cy) / 2), sin(y/2) = sy / (2 cos(y/2))              for cy >= 0,
//   sin(y/2) = sign(s) sqrt((1 - cy) / 2), cos(y/2) = sy / (2 sin(y/2))      for cy < 0
// (the well-conditioned form on either side; s = c = 0 gives the identity like atan2(0, 0) = 0).
template <typename F> APE_HD Quat<F> hips_quat(F s, F c) {
    const F n2 = s * s + c * c;
    if (!(n2 > F(0))) return {F(1), F(0), F(0), F(0)};
    const F ir = inv_sqrt(n2), cy = c * ir, sy = s * ir;
    const F a = F(0.5) * (F(1) + (cy < F(0) ? -cy : cy)), ia = inv_sqrt(a);          // a in [0.5, 1]
    const F big = a * ia, small = F(0.5) * (sy < F(0) ? -sy : sy) * ia;              // sqrt(a), |sy| / (2 sqrt(a))
    return cy >= F(0) ? Quat<F>{big, F(0), copysign(small, s), F(0)} : Quat<F>{small, F(0), copysign(big, s), F(0)};
}

template <typename F> struct Body {
    Vec3<F> larm_vec, uarm_vec, uarm_orig;
    bool bones_along_x = false;      // both bone vectors are (len, 0, 0): chain() may take the short rotation (set by the caller)
};
template <typename F> APE_HD bool bones_are_along_x(const Body<F>& b) {
    return b.larm_vec.y == F(0) && b.larm_vec.z == F(0) && b.uarm_vec.y == F(0) && b.uarm_vec.z == F(0);
}

template <typename F> struct RowPose {
    Quat<F> larm, uarm, hips;
    Vec3<F> hand, elbow, shoulder;
};

template <typename F> APE_HD Vec3<F> vadd(const Vec3<F>& a, const Vec3<F>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }

// shoulder -> elbow -> hand from the three quaternions (estimate_joints.py:61-63 / :84-85, compose_msg.py:59-61 / :92-93)
template <typename F> APE_HD void chain(int target, const Body<F>& body, RowPose<F>& r) {
    r.shoulder = target == APE_TARGET_ORI_CAL_LARM_UARM ? body.uarm_orig : qrot(r.hips, body.uarm_orig);
    if (body.bones_along_x) {
        r.elbow = vadd(qrot_x(r.uarm, body.uarm_vec.x), r.shoulder);
        r.hand = vadd(qrot_x(r.larm, body.larm_vec.x), r.elbow);
    } else {
        r.elbow = vadd(qrot(r.uarm, body.uarm_vec), r.shoulder);
        r.hand = vadd(qrot(r.larm, body.larm_vec), r.elbow);
    }
}

// One de-normalised target row -> quaternions + joint origins (estimate_joints.py:20-92).
template <typename F> APE_HD RowPose<F> row_pose(int target, const F* p, const Body<F>& body, bool& bad) {
    RowPose<F> r;
    r.hips = {F(1), F(0), F(0), F(0)};
    if (target == APE_TARGET_ORI_POS_CAL_LARM_UARM_HIPS) {
        r.larm = six_to_quat(p + 3, bad);
        r.uarm = six_to_quat(p + 12, bad);
        r.hips = hips_quat(p[18], p[19]);
        r.hand = {p[0], p[1], p[2]};
        r.elbow = {p[9], p[10], p[11]};
        r.shoulder = qrot(r.hips, body.uarm_orig);
        return r;
    }
    r.larm = six_to_quat(p, bad);
    r.uarm = six_to_quat(p + 6, bad);
    if (target == APE_TARGET_ORI_CAL_LARM_UARM_HIPS) r.hips = hips_quat(p[12], p[13]);
    chain(target, body, r);
    return r;
}

APE_HD int target_num_outputs(int target) {
    return target == APE_TARGET_ORI_CAL_LARM_UARM ? 12 : (target == APE_TARGET_ORI_CAL_LARM_UARM_HIPS ? 14 : 20);
}

}  // namespace ape
