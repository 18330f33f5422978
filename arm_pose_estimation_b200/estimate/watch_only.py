"""Watch-only MC-dropout LSTM estimator (``estimate/watch_only.py:13-97`` of the reference)."""
import numpy as np

from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200.data_deploy.nn import deploy_models
from arm_pose_estimation_b200.data_types.bone_map import BoneMap
from arm_pose_estimation_b200.estimate.estimator import _NNEstimator


class WatchOnlyNN(_NNEstimator):
    _kind = N.KIND_WATCH_ONLY
    _xx_dtype = np.float32                                    # watch_only.py:82

    def __init__(self,
                 model_hash: str = deploy_models.LSTM.WATCH_ONLY.value,
                 smooth: int = 10,
                 add_mc_samples=True,
                 monte_carlo_samples=25,
                 bonemap: BoneMap = None,
                 watch_phone: bool = False,
                 tag: str = "PUB WATCH",
                 philox_seed: int = None):
        # watch_phone=True reads the watch columns out of the 55-float layout (watch_only.py:28-31)
        self._layout = N.LAYOUT_WATCH_PHONE if watch_phone else N.LAYOUT_WATCH_ONLY
        self._init_nn(model_hash, smooth, add_mc_samples, monte_carlo_samples, bonemap, tag, philox_seed)
