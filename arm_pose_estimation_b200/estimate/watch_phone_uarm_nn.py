"""Watch + phone-on-upper-arm MC-dropout LSTM estimator (``estimate/watch_phone_uarm_nn.py:13-121`` of the reference)."""
import numpy as np

from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200.data_deploy.nn import deploy_models
from arm_pose_estimation_b200.data_types.bone_map import BoneMap
from arm_pose_estimation_b200.estimate.estimator import _NNEstimator


class WatchPhoneUarmNN(_NNEstimator):
    _kind = N.KIND_UARM
    _layout = N.LAYOUT_WATCH_PHONE
    _xx_dtype = np.float64                                    # no dtype at watch_phone_uarm_nn.py:99

    def __init__(self,
                 model_hash: str = deploy_models.LSTM.WATCH_PHONE_UARM.value,
                 smooth: int = 1,
                 add_mc_samples=True,
                 monte_carlo_samples=50,
                 bonemap: BoneMap = None,
                 tag: str = "NN UARM PHONE",
                 philox_seed: int = None):
        self._init_nn(model_hash, smooth, add_mc_samples, monte_carlo_samples, bonemap, tag, philox_seed)
