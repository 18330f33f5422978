"""Direct-orientation estimator: calibrated watch / phone orientations ARE the lower / upper arm orientations - no network
(``estimate/watch_phone_uarm.py:10-108`` of the reference).  Stage 1 and stage 3 are the same CUDA kernels as the NN
estimators'; "prediction" is a column selection of the feature row (``:107-108``)."""
import numpy as np

from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200.data_types import messaging
from arm_pose_estimation_b200.data_types.bone_map import BoneMap
from arm_pose_estimation_b200.estimate import estimate_joints
from arm_pose_estimation_b200.estimate.estimator import Estimator, parsed_feature_row
from arm_pose_estimation_b200.utility.names import NNS_INPUTS, NNS_TARGETS


class WatchPhoneUarm(Estimator):
    _kind = N.KIND_UARM
    _layout = N.LAYOUT_WATCH_PHONE

    def __init__(self, smooth: int = 5, tag: str = "Forward Kinematics", bonemap: BoneMap = None):
        super().__init__(
            x_inputs=NNS_INPUTS.WATCH_PHONE_CAL_ALL,
            y_targets=NNS_TARGETS.ORI_CAL_LARM_UARM,
            smooth=smooth,
            normalize=False,
            seq_len=1,
            add_mc_samples=False,
            tag=tag,
            bonemap=bonemap,
        )

    def parse_row_to_xx(self, row):
        """The calibrated 38-float feature row (watch_phone_uarm.py:56-105): the stage-1 kernel with ``normalize=0``."""
        return parsed_feature_row(row, self._layout, self._kind, len(self._x_inputs.value), np.float64)

    def make_prediction_from_row_hist(self, row_hist):
        return np.c_[row_hist[:, 13:19], row_hist[:, -6:]]   # watch 6D -> lower arm, phone 6D -> upper arm (:107-108)

    def calibrate_orientation_quats(self, sw_quat, sw_fwd, ph_quat, ph_fwd):
        """``(sw_cal_g, ph_cal_g)``: watch / phone rotation quaternions aligned to north and to the calibration pose
        (watch_phone_uarm.py:32-54), each ``[w, x, y, z]``.  Evaluated by the same kernels as the estimation path - the four
        quaternions travel as a wire row through stage 1, and stage 3 turns the calibrated 6D rotations back into quaternions -
        so the result is float32 and in the canonical sign ``w >= 0`` (q and -q are the same rotation)."""
        row = np.zeros(len(messaging.WATCH_PHONE_IMU_LOOKUP), dtype=np.float32)
        for prefix, quat in (("sw_rotvec", sw_quat), ("sw_forward", sw_fwd), ("ph_rotvec", ph_quat), ("ph_forward", ph_fwd)):
            for axis, value in zip("wxyz", np.asarray(quat, dtype=np.float64).ravel()):
                row[messaging.WATCH_PHONE_IMU_LOOKUP[f"{prefix}_{axis}"]] = value
        pred = self.make_prediction_from_row_hist(self.parse_row_to_xx(row)[None, :])
        est = estimate_joints.fk_rows(pred, self._body_measurements, self._y_targets)[0]
        return est[6:10], est[10:14]                           # est row: hand3, elbow3, larm_q4, uarm_q4
