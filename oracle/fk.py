"""Stage 3 of the path: network targets -> quaternions + joint origins -> 25-float message (numpy float64).

Follows ``estimate/estimate_joints.py`` and ``estimate/compose_msg.py`` of the reference; ``target`` is the
``NNS_TARGETS`` member NAME.  Test infrastructure only.
"""
import numpy as np

from oracle import quat as Q

TARGETS = ("ORI_CAL_LARM_UARM", "ORI_CAL_LARM_UARM_HIPS", "ORI_POS_CAL_LARM_UARM_HIPS")


def arm_pose_from_nn_targets(preds, body, target):
    """estimate_joints.py:16-17 dispatch.  ``preds (S, O)`` de-normalised, ``body (1, 9)``."""
    preds, body = np.asarray(preds, dtype=np.float64), np.asarray(body, dtype=np.float64)
    larm_vec, uarm_vec, uarm_orig = body[:, :3], body[:, 3:6], body[:, 6:]
    if target == "ORI_CAL_LARM_UARM":                                   # estimate_joints.py:74-92
        uarm_q, larm_q = Q.six_to_quat(preds[:, 6:]), Q.six_to_quat(preds[:, :6])
        elbow = Q.rotate(uarm_q, uarm_vec) + uarm_orig
        hand = Q.rotate(larm_q, larm_vec) + elbow
        return np.hstack([hand, elbow, larm_q, uarm_q])
    if target == "ORI_CAL_LARM_UARM_HIPS":                              # estimate_joints.py:48-71
        uarm_q, larm_q = Q.six_to_quat(preds[:, 6:12]), Q.six_to_quat(preds[:, :6])
        hips_q = Q.hips_sin_cos_to_quat(preds[:, 12], preds[:, 13])
        shoulder = Q.rotate(hips_q, uarm_orig)
        elbow = Q.rotate(uarm_q, uarm_vec) + shoulder
        hand = Q.rotate(larm_q, larm_vec) + elbow
        return np.hstack([hand, elbow, shoulder, larm_q, uarm_q, hips_q])
    if target == "ORI_POS_CAL_LARM_UARM_HIPS":                          # estimate_joints.py:20-45
        uarm_q, larm_q = Q.six_to_quat(preds[:, 12:18]), Q.six_to_quat(preds[:, 3:9])
        hips_q = Q.hips_sin_cos_to_quat(preds[:, 18], preds[:, 19])
        shoulder = Q.rotate(hips_q, uarm_orig)
        return np.hstack([preds[:, :3], preds[:, 9:12], shoulder, larm_q, uarm_q, hips_q])
    raise KeyError(target)


def msg_from_est(est, body, target):
    """compose_msg.py:13-14 dispatch -> ``[larm_q, hand, larm_q, elbow, uarm_q, shoulder, hips_q]`` (25,)."""
    est, body = np.asarray(est, dtype=np.float64), np.asarray(body, dtype=np.float64)
    larm_vec, uarm_vec, uarm_orig = body[0, :3], body[0, 3:6], body[0, 6:]
    many = est.shape[0] > 1
    if target == "ORI_CAL_LARM_UARM":                                   # compose_msg.py:82-108
        if many:
            larm_q, uarm_q = Q.average_quats(est[:, 6:10]), Q.average_quats(est[:, 10:])
            elbow = Q.rotate(uarm_q, uarm_vec) + uarm_orig              # FK again from the means (:92-93)
            hand = Q.rotate(larm_q, larm_vec) + elbow
        else:
            hand, elbow, larm_q, uarm_q = est[0, :3], est[0, 3:6], est[0, 6:10], est[0, 10:]
        return np.hstack([larm_q, hand, larm_q, elbow, uarm_q, uarm_orig, np.array([1.0, 0, 0, 0])])
    if target == "ORI_CAL_LARM_UARM_HIPS":                              # compose_msg.py:48-79
        if many:
            hips_q = Q.average_quats(est[:, 17:])
            larm_q, uarm_q = Q.average_quats(est[:, 9:13]), Q.average_quats(est[:, 13:17])
            shoulder = Q.rotate(hips_q, uarm_orig)
            elbow = Q.rotate(uarm_q, uarm_vec) + shoulder
            hand = Q.rotate(larm_q, larm_vec) + elbow
        else:
            hand, elbow, shoulder = est[0, 0:3], est[0, 3:6], est[0, 6:9]
            larm_q, uarm_q, hips_q = est[0, 9:13], est[0, 13:17], est[0, 17:]
        return np.hstack([larm_q, hand, larm_q, elbow, uarm_q, shoulder, hips_q])
    if target == "ORI_POS_CAL_LARM_UARM_HIPS":                          # compose_msg.py:17-45
        if many:
            hips_q = Q.average_quats(est[:, 17:])
            larm_q, uarm_q = Q.average_quats(est[:, 9:13]), Q.average_quats(est[:, 13:17])
            shoulder, elbow, hand = est[:, 6:9].mean(axis=0), est[:, 3:6].mean(axis=0), est[:, 0:3].mean(axis=0)
        else:
            hand, elbow, shoulder = est[0, 0:3], est[0, 3:6], est[0, 6:9]
            larm_q, uarm_q, hips_q = est[0, 9:13], est[0, 13:17], est[0, 17:]
        return np.hstack([larm_q, hand, larm_q, elbow, uarm_q, shoulder, hips_q])
    raise KeyError(target)


def sample_std(est):
    """Population std (ddof=0) of the per-sample hand/elbow positions ``est[:, :6]`` - NOT in the reference
    (it ships the samples in the message tail, estimator.py:131-136); defined in SURVEY.md §8a (new)."""
    return np.std(np.asarray(est, dtype=np.float64)[:, :6], axis=0)
