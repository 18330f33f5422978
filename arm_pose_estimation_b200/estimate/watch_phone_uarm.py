"""Direct-orientation estimator: calibrated watch / phone orientations ARE the lower / upper arm orientations - no network
(``estimate/watch_phone_uarm.py:10-108`` of the reference).  Stage 1 and stage 3 are the same CUDA kernels as the NN
estimators'; "prediction" is a column selection of the feature row (``:107-108``)."""
import numpy as np

from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200.data_types.bone_map import BoneMap
from arm_pose_estimation_b200.estimate.estimator import Estimator, _NNEstimator
from arm_pose_estimation_b200.utility.names import NNS_INPUTS, NNS_TARGETS


class WatchPhoneUarm(Estimator):
    _kind = N.KIND_UARM
    _layout = N.LAYOUT_WATCH_PHONE
    _xx_dtype = np.float64

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

    parse_row_to_xx = _NNEstimator.parse_row_to_xx           # the calibrated 38-float feature row (stage-1 kernel)

    def make_prediction_from_row_hist(self, row_hist):
        return np.c_[row_hist[:, 13:19], row_hist[:, -6:]]   # watch 6D -> lower arm, phone 6D -> upper arm (:107-108)

    def estimate_row(self, row, add_mc_samples=None):
        add = self._add_mc_samples if add_mc_samples is None else add_mc_samples
        return self.msg_from_pred(self.add_xx_to_row_hist_and_make_prediction(self.parse_row_to_xx(row)), add)
