"""Network targets -> per-row quaternions + joint origins (``estimate/estimate_joints.py:16-92`` of the reference).

Same call, same row layouts (``[hand3, elbow3, larm_q4, uarm_q4]`` or ``[hand3, elbow3, shoulder3, larm_q4, uarm_q4,
hips_q4]``); the arithmetic is the ``est_rows`` output of the CUDA stage-3 kernel (``ape_fk_reduce``), float32.
"""
import numpy as np
import torch

from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200.utility.names import NNS_TARGETS

TARGET_IDS = {
    NNS_TARGETS.ORI_CAL_LARM_UARM: N.TARGET_ORI_CAL_LARM_UARM,
    NNS_TARGETS.ORI_CAL_LARM_UARM_HIPS: N.TARGET_ORI_CAL_LARM_UARM_HIPS,
    NNS_TARGETS.ORI_POS_CAL_LARM_UARM_HIPS: N.TARGET_ORI_POS_CAL_LARM_UARM_HIPS,
}
EST_WIDTH = {N.TARGET_ORI_CAL_LARM_UARM: 14, N.TARGET_ORI_CAL_LARM_UARM_HIPS: 21, N.TARGET_ORI_POS_CAL_LARM_UARM_HIPS: 21}


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("arm_pose_estimation_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def fk_rows(preds, body_measurements, y_targets, want_msg=False):
    """``preds (S, O)`` de-normalised -> ``est (S, W)`` float64 [, ``msg (25,)``, ``std (6,)``]; raises
    ``numpy.linalg.LinAlgError`` on a degenerate 6D pair, like the reference's ``eigh``."""
    _require_cuda()
    target = TARGET_IDS[y_targets]
    p = torch.as_tensor(np.ascontiguousarray(np.asarray(preds, dtype=np.float32))).cuda()
    if p.dim() != 2 or p.shape[1] != len(y_targets.value):
        raise UserWarning(f"preds must be (rows, {len(y_targets.value)}) for {y_targets.name}, got {tuple(p.shape)}")
    S, O = int(p.shape[0]), int(p.shape[1])
    body = torch.as_tensor(np.asarray(body_measurements, dtype=np.float32).ravel()).cuda()
    W = EST_WIDTH[target]
    est = torch.empty((S, W), dtype=torch.float32, device="cuda")
    msg = torch.empty(25, dtype=torch.float32, device="cuda")
    std = torch.empty(6, dtype=torch.float32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    N.check(N.load().ape_fk_reduce(N.ptr(p), 1, None, None, N.ptr(body), target, O, 1, 1, 0, None, S, 1,
                                   N.ptr(msg), None, N.ptr(std), N.ptr(est), N.ptr(status), N.current_stream_ptr()),
            "ape_fk_reduce")
    if int(status.item()) != 0:
        raise np.linalg.LinAlgError("degenerate 6D rotation (zero or collinear columns)")
    est = est.cpu().numpy().astype(np.float64)
    if want_msg:
        return est, msg.cpu().numpy().astype(np.float64), std.cpu().numpy().astype(np.float64)
    return est


def arm_pose_from_nn_targets(preds: np.array, body_measurements: np.array, y_targets: NNS_TARGETS):
    return fk_rows(preds, body_measurements, y_targets)
