"""Stage 1 of the path: one raw IMU row -> one feature row (numpy float64 restatement).

``kind`` selects the estimator whose ``parse_row_to_xx`` is followed:
``"watch_only"`` (estimate/watch_only.py:46-82), ``"pocket"`` (estimate/watch_phone_pocket_nn.py:41-96),
``"uarm"`` (estimate/watch_phone_uarm_nn.py:43-105).  Test infrastructure only.
"""
import numpy as np

from oracle import quat as Q

KINDS = ("watch_only", "pocket", "uarm")

_SW_SENSOR = (["sw_dt"] + [f"sw_{g}_{a}" for g in ("gyro", "lvel", "lacc", "grav") for a in "xyz"])
_PH_SENSOR = [f"ph_{g}_{a}" for g in ("gyro", "lvel", "lacc", "grav") for a in "xyz"]
_LARM_DST = np.array([-0.7071068, 0.0, -0.7071068, 0.0])   # watch_phone_uarm_nn.py:84
_UARM_DST = np.array([0.7071068, 0.0, 0.7071068, 0.0])     # watch_phone_uarm_nn.py:85


def _q(row, lk, stem):
    return np.array([row[lk[f"{stem}_{c}"]] for c in "wxyz"], dtype=np.float64)


def parse_row(kind, row, lk):
    """``row``: indexable of floats in layout ``lk`` (a messaging lookup).  Returns the feature vector with
    the dtype the reference returns: float32 for watch_only / pocket (``dtype=np.float32`` at
    watch_only.py:82, watch_phone_pocket_nn.py:96), float64 for uarm (no dtype at watch_phone_uarm_nn.py:99)."""
    r_pres = float(row[lk["sw_pres"]]) - float(row[lk["sw_init_pres"]])
    sw_rot, sw_fwd = _q(row, lk, "sw_rotvec"), _q(row, lk, "sw_forward")
    sw_sensor = np.array([row[lk[k]] for k in _SW_SENSOR], dtype=np.float64)

    if kind in ("watch_only", "pocket"):
        north = Q.north_quat(sw_fwd)                                   # watch_only.py:67-69
        sw_six = Q.quat_to_six(Q.android_to_global(sw_rot, north))     # watch_only.py:71-72
        cols = [sw_sensor, sw_six, [r_pres]]
        if kind == "pocket":
            ph_rot, ph_fwd = _q(row, lk, "ph_rotvec"), _q(row, lk, "ph_forward")
            ph_rot_g = Q.android_to_global(ph_rot, north)              # watch_phone_pocket_nn.py:76
            ph_fwd_g = Q.android_to_global(ph_fwd, north)              # :77
            ph_cal = Q.hamilton(ph_rot_g, Q.invert(ph_fwd_g))          # :78
            hips_y = Q.y_rot_of(ph_cal)                                # :81
            cols += [[np.sin(hips_y)], [np.cos(hips_y)]]               # :82-83
        return np.hstack(cols).astype(np.float32)

    if kind == "uarm":
        ph_rot, ph_fwd = _q(row, lk, "ph_rotvec"), _q(row, lk, "ph_forward")
        ph_sensor = np.array([row[lk[k]] for k in _PH_SENSOR], dtype=np.float64)
        north = Q.north_quat_left_arm(sw_fwd)                          # watch_phone_uarm_nn.py:82
        sw_rot_g, sw_fwd_g = Q.android_to_global(sw_rot, north), Q.android_to_global(sw_fwd, north)   # :88-89
        sw_cal = Q.hamilton(sw_rot_g, Q.hamilton(Q.invert(sw_fwd_g), _LARM_DST))                      # :90-91
        ph_rot_g, ph_fwd_g = Q.android_to_global(ph_rot, north), Q.android_to_global(ph_fwd, north)   # :94-95
        ph_cal = Q.hamilton(ph_rot_g, Q.hamilton(Q.invert(ph_fwd_g), _UARM_DST))                      # :96-97
        return np.hstack([sw_sensor, Q.quat_to_six(sw_cal), [r_pres], ph_sensor, Q.quat_to_six(ph_cal)])

    raise ValueError(kind)
