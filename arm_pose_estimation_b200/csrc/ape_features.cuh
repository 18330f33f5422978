// Stage 1 math: one raw IMU row -> one feature row.  Host/device so the CPU box can test it.
// Follows parse_row_to_xx of the three NN estimators; float64 like the reference's numpy code.
#pragma once
#include "ape_common.cuh"

namespace ape {

// positions inside one device block of a raw row (data_types/messaging.py:20-58)
enum { DEV_DT = 0, DEV_ROT = 5, DEV_GYRO = 10, DEV_LVEL = 13, DEV_LACC = 16, DEV_PRES = 19, DEV_GRAV = 20, DEV_LEN = 23 };

struct RowLayout { int ncols, sw, ph, sw_fwd, ph_fwd, init_pres; };

APE_HD RowLayout row_layout(int layout) {
    // watch-only: [sw block | sw_forward | sw_init_pres]; watch+phone: [sw | ph | sw_forward | ph_forward | sw_init_pres]
    if (layout == APE_LAYOUT_WATCH_ONLY) return {28, 0, -1, 23, -1, 27};
    return {55, 0, 23, 46, 50, 54};
}

APE_HD int kind_num_features(int kind) { return kind == APE_KIND_WATCH_ONLY ? 20 : (kind == APE_KIND_POCKET ? 22 : 38); }

template <typename Row> APE_HD Quat<double> load_quat(const Row& row, int at) {
    return {(double)row[at], (double)row[at + 1], (double)row[at + 2], (double)row[at + 3]};
}

// atan2(x, z) of the rotated forward vector [0,0,1] (transformations.py:200-207)
APE_HD double y_rot_of(const Quat<double>& q) {
    Vec3<double> pp = qrot(q, Vec3<double>{0.0, 0.0, 1.0});
    return atan2(pp.x, pp.z);
}

// euler_to_quat([0, -y_rot, 0]) of the swapped calibration quaternion (watch_only.py:67-69)
APE_HD Quat<double> north_quat(const Quat<double>& sw_fwd) {
    double y = y_rot_of(android_swap(sw_fwd));
    return {cos(-y * 0.5), 0.0, sin(-y * 0.5), 0.0};
}

// first two columns of the rotation matrix of q, row-major [r11,r12,r21,r22,r31,r32]
// (transformations.py:476-518 and :587-599; s = 2/|q|^2, identity below eps)
APE_HD void quat_to_six(const Quat<double>& q, double* six) {
    double nq = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    if (nq < 2.220446049250313e-16) {
        six[0] = 1; six[1] = 0; six[2] = 0; six[3] = 1; six[4] = 0; six[5] = 0;
        return;
    }
    double s = 2.0 / nq;
    double X = q.x * s, Y = q.y * s, Z = q.z * s;
    double wX = q.w * X, wY = q.w * Y, wZ = q.w * Z;
    double xX = q.x * X, xY = q.x * Y, xZ = q.x * Z;
    double yY = q.y * Y, yZ = q.y * Z, zZ = q.z * Z;
    six[0] = 1.0 - (yY + zZ); six[1] = xY - wZ;
    six[2] = xY + wZ;         six[3] = 1.0 - (xX + zZ);
    six[4] = xZ - wY;         six[5] = yZ + wX;
}

// 13 sensor floats of a device block in NNS_INPUTS order: dt, gyro, lvel, lacc, grav (names.py:35-40)
template <typename Row> APE_HD void sw_sensor13(const Row& row, int base, double* out) {
    out[0] = row[base + DEV_DT];
    for (int i = 0; i < 3; ++i) {
        out[1 + i] = row[base + DEV_GYRO + i];
        out[4 + i] = row[base + DEV_LVEL + i];
        out[7 + i] = row[base + DEV_LACC + i];
        out[10 + i] = row[base + DEV_GRAV + i];
    }
}

// Raw feature row (before the z-score).  Returns the number of features written.
template <typename Row> APE_HD int compute_features(int kind, int layout, const Row& row, double* xx) {
    const RowLayout lay = row_layout(layout);
    const double r_pres = (double)row[lay.sw + DEV_PRES] - (double)row[lay.init_pres];
    const Quat<double> sw_rot = load_quat(row, lay.sw + DEV_ROT), sw_fwd = load_quat(row, lay.sw_fwd);
    sw_sensor13(row, lay.sw, xx);

    if (kind == APE_KIND_WATCH_ONLY || kind == APE_KIND_POCKET) {
        const Quat<double> north = north_quat(sw_fwd);
        quat_to_six(qmul(north, android_swap(sw_rot)), xx + 13);         // watch_only.py:71-72
        xx[19] = r_pres;
        int n = 20;
        if (kind == APE_KIND_POCKET) {                                    // watch_phone_pocket_nn.py:76-83
            const Quat<double> ph_rot_g = qmul(north, android_swap(load_quat(row, lay.ph + DEV_ROT)));
            const Quat<double> ph_fwd_g = qmul(north, android_swap(load_quat(row, lay.ph_fwd)));
            const double hips_y = y_rot_of(qmul(ph_rot_g, qinv(ph_fwd_g)));
            xx[20] = sin(hips_y);
            xx[21] = cos(hips_y);
            n = 22;
        }
        for (int i = 0; i < n; ++i) xx[i] = (double)(float)xx[i];         // np.hstack(..., dtype=np.float32)
        return n;
    }
    // APE_KIND_UARM: watch_phone_uarm_nn.py:82-105 (stays float64 until the z-score)
    const Quat<double> north = qmul(Quat<double>{0.7071068, 0.0, -0.7071068, 0.0}, north_quat(sw_fwd));
    const Quat<double> larm_dst = {-0.7071068, 0.0, -0.7071068, 0.0}, uarm_dst = {0.7071068, 0.0, 0.7071068, 0.0};
    const Quat<double> sw_rot_g = qmul(north, android_swap(sw_rot)), sw_fwd_g = qmul(north, android_swap(sw_fwd));
    quat_to_six(qmul(sw_rot_g, qmul(qinv(sw_fwd_g), larm_dst)), xx + 13);
    xx[19] = r_pres;
    for (int i = 0; i < 3; ++i) {                                         // phone gyro, lvel, lacc, grav
        xx[20 + i] = row[lay.ph + DEV_GYRO + i];
        xx[23 + i] = row[lay.ph + DEV_LVEL + i];
        xx[26 + i] = row[lay.ph + DEV_LACC + i];
        xx[29 + i] = row[lay.ph + DEV_GRAV + i];
    }
    const Quat<double> ph_rot_g = qmul(north, android_swap(load_quat(row, lay.ph + DEV_ROT)));
    const Quat<double> ph_fwd_g = qmul(north, android_swap(load_quat(row, lay.ph_fwd)));
    quat_to_six(qmul(ph_rot_g, qmul(qinv(ph_fwd_g), uarm_dst)), xx + 32);
    return 38;
}

}  // namespace ape
