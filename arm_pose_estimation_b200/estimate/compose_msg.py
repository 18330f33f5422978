"""Per-row estimates -> the 25-float pose message (``estimate/compose_msg.py:13-108`` of the reference).

``[larm_q4, hand3, larm_q4, elbow3, uarm_q4, shoulder3, hips_q4]``: quaternions averaged over the rows after
sign alignment to row 0, joint origins re-derived from the averaged quaternions; one row is copied as is.
Runs the reduction half of the CUDA stage-3 kernel (``ape_msg_from_est``), float32.
"""
import numpy as np
import torch

from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200.estimate.estimate_joints import EST_WIDTH, TARGET_IDS, _require_cuda
from arm_pose_estimation_b200.utility.names import NNS_TARGETS


def msg_from_nn_targets_est(est: np.array, body_measure: np.array, y_targets: NNS_TARGETS):
    _require_cuda()
    target = TARGET_IDS[y_targets]
    W = EST_WIDTH[target]
    e = torch.as_tensor(np.ascontiguousarray(np.asarray(est, dtype=np.float32))).cuda()
    if e.dim() != 2 or e.shape[1] != W:
        raise UserWarning(f"est must be (rows, {W}) for {y_targets.name}, got {tuple(e.shape)}")
    body = torch.as_tensor(np.asarray(body_measure, dtype=np.float32).ravel()).cuda()
    msg = torch.empty(25, dtype=torch.float32, device="cuda")
    N.check(N.load().ape_msg_from_est(N.ptr(e), W, N.ptr(body), target, 1, int(e.shape[0]), N.ptr(msg), None,
                                      N.current_stream_ptr()), "ape_msg_from_est")
    return msg.cpu().numpy().astype(np.float64)
