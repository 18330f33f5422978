"""Column-name tables for model inputs and targets.

Mirrors the members of ``NNS_TARGETS`` / ``NNS_INPUTS`` that the reference defines in
``src/wear_mocap_ape/utility/names.py:4-110`` (same member names, same column order), so that
``NNS_INPUTS[params["x_inputs_n"]]`` and ``NNS_TARGETS[params["y_targets_n"]]`` (reference
``estimate/watch_only.py:36-37``) resolve identically.  The lists are assembled from column groups
instead of being spelled out; the stdlib ``enum`` is used (the reference needs ``aenum`` only for
``NoAlias``, which matters for two unrelated label members that are omitted here).
"""
from enum import Enum


def _xyz(stem):
    return [f"{stem}_{a}" for a in "xyz"]


def _six(stem):
    return [f"{stem}_{i}" for i in range(1, 7)]


_SW_IMU = ["sw_dt"] + _xyz("sw_gyro") + _xyz("sw_lvel") + _xyz("sw_lacc") + _xyz("sw_grav")
_PH_IMU = _xyz("ph_gyro") + _xyz("ph_lvel") + _xyz("ph_lacc") + _xyz("ph_grav")
_HIPS_IN = ["ph_hips_yrot_cal_sin", "ph_hips_yrot_cal_cos"]
_HIPS_GT = ["gt_hips_yrot_cal_sin", "gt_hips_yrot_cal_cos"]
_ACC_ONLY = ["sw_dt"] + _xyz("sw_lacc") + _six("sw_6drr_cal")


class NNS_TARGETS(Enum):
    # network output columns; consumed by estimate_joints / compose_msg dispatch tables
    ORI_CAL_LARM_UARM_HIPS = _six("gt_larm_6drr_cal") + _six("gt_uarm_6drr_cal") + _HIPS_GT
    ORI_CAL_LARM_UARM = _six("gt_larm_6drr_cal") + _six("gt_uarm_6drr_cal")
    ORI_POS_CAL_LARM_UARM_HIPS = (
        _xyz("gt_hand_orig_cal") + _six("gt_larm_6drr_cal") + _xyz("gt_larm_orig_cal")
        + _six("gt_uarm_6drr_cal") + _HIPS_GT
    )


class NNS_INPUTS(Enum):
    # network input columns
    WATCH_ONLY_CAL = _SW_IMU + _six("sw_6drr_cal") + ["sw_pres_cal"]
    WATCH_ONLY_ACC_ONLY = list(_ACC_ONLY)
    WATCH_PHONE_CAL_HIP = _SW_IMU + _six("sw_6drr_cal") + ["sw_pres_cal"] + _HIPS_IN
    WATCH_HIP_ACC_ONLY = _ACC_ONLY + _HIPS_IN
    WATCH_HIP_ACC_AND_BAR = _ACC_ONLY + ["sw_pres_cal"] + _HIPS_IN
    WATCH_PHONE_CAL_ALL = _SW_IMU + _six("sw_6drr_cal") + ["sw_pres_cal"] + _PH_IMU + _six("ph_6drr_cal")
    WATCH_ONLY_RAW = _SW_IMU + _six("sw_6drr_raw") + ["sw_pres_cal"]
