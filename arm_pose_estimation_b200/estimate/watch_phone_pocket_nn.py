"""Watch + phone-in-pocket MC-dropout LSTM estimator (``estimate/watch_phone_pocket_nn.py:12-112`` of the reference)."""
import numpy as np

from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200.data_types.bone_map import BoneMap
from arm_pose_estimation_b200.estimate.estimator import _NNEstimator


class WatchPhonePocketNN(_NNEstimator):
    _kind = N.KIND_POCKET
    _layout = N.LAYOUT_WATCH_PHONE
    _xx_dtype = np.float32                                    # watch_phone_pocket_nn.py:96

    def __init__(self,
                 model_hash: str,
                 smooth: int = 1,
                 add_mc_samples=True,
                 monte_carlo_samples=25,
                 bonemap: BoneMap = None,
                 tag: str = "NN POCKET PHONE",
                 philox_seed: int = None):
        self._init_nn(model_hash, smooth, add_mc_samples, monte_carlo_samples, bonemap, tag, philox_seed)
