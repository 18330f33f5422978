"""CPU: the oracle against the committed outputs of the reference itself (tests/golden/make_golden.py)."""
import json

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, unpack_masks
from arm_pose_estimation_b200 import synthetic as syn
from arm_pose_estimation_b200.data_types import messaging
from arm_pose_estimation_b200.utility import names
from oracle import estimator as OE, features as OF, fk as OFK, lstm as OL, quat as OQ

KINDS = (syn.KIND_WATCH_ONLY, syn.KIND_POCKET, syn.KIND_UARM)


def test_tables_match_reference():
    t = json.loads((GOLDEN / "tables.json").read_text())
    assert t["watch_only_lookup"] == messaging.WATCH_ONLY_IMU_LOOKUP
    assert t["watch_phone_lookup"] == messaging.WATCH_PHONE_IMU_LOOKUP
    assert messaging.watch_only_imu_msg_len == 112 and messaging.watch_phone_imu_msg_len == 220
    for m in names.NNS_INPUTS:
        assert t["inputs"][m.name] == list(m.value)
    for m in names.NNS_TARGETS:
        assert t["targets"][m.name] == list(m.value)


def test_quat_leaves():
    g = load_golden("quat_leaves.npz")
    a, b, qa, v, e, six, sc, near = (g[k] for k in ("a", "b", "qa", "v", "e", "six", "s_c", "near"))
    pairs = [
        (OQ.hamilton(a, b), "hamilton"), (OQ.rotate(a, v), "rotate"), (OQ.rotate(a, v[0]), "rotate_single_vec"),
        (OQ.invert(a), "invert"), (OQ.euler_to_quat(e), "euler"), (OQ.android_to_global_no_north(a), "a2g_no_north"),
        (OQ.android_to_global(a, b), "a2g"), (OQ.y_rot_of(qa), "y_rot"), (OQ.north_quat_left_arm(qa), "north_left"),
        (OQ.quat_to_rot9(a), "rot9"), (OQ.quat_to_six(a), "six_of_q"), (OQ.six_to_rot9(six), "rot9_of_six"),
        (OQ.six_to_quat(six), "quat_of_six"), (OQ.hips_sin_cos_to_quat(sc[:, 0], sc[:, 1]), "hips"),
        (OQ.average_quats(near), "average"),
    ]
    for mine, key in pairs:
        np.testing.assert_allclose(mine, g[key], rtol=0, atol=1e-12, err_msg=key)


def test_known_answers():
    # read straight off the reference source (SURVEY.md §4 ii): identity 6D -> identity quaternion -> straight arm
    q = OQ.six_to_quat(np.array([[1.0, 0, 0, 1.0, 0, 0]]))
    np.testing.assert_allclose(q, [[1, 0, 0, 0]], atol=1e-15)
    body = np.array([[-0.22, 0, 0, -0.26, 0, 0, -0.17, 0.43, -0.006]])
    est = OFK.arm_pose_from_nn_targets(np.array([[1.0, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0]]), body, "ORI_CAL_LARM_UARM")
    np.testing.assert_allclose(est[0, 3:6], [-0.43, 0.43, -0.006], atol=1e-15)       # elbow = shoulder + uarm_vec
    np.testing.assert_allclose(est[0, :3], [-0.65, 0.43, -0.006], atol=1e-15)        # hand  = elbow + larm_vec
    msg = OFK.msg_from_est(est, body, "ORI_CAL_LARM_UARM")
    assert msg.shape == (25,) and list(msg[21:]) == [1, 0, 0, 0]
    # quat -> 6D -> quat is the identity up to sign; rotation preserves length
    rng = np.random.default_rng(0)
    qs = rng.normal(size=(16, 4))
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    back = OQ.six_to_quat(OQ.quat_to_six(qs))
    np.testing.assert_allclose(back * np.sign(np.sum(back * qs, 1, keepdims=True)), qs, atol=1e-12)
    v = rng.normal(size=(16, 3))
    np.testing.assert_allclose(np.linalg.norm(OQ.rotate(qs, v), axis=1), np.linalg.norm(v, axis=1), atol=1e-12)
    # degenerate 6D (zero column): the reference raises (LinAlgError from eigh on NaNs)
    with np.errstate(all="ignore"), pytest.raises(np.linalg.LinAlgError):
        OQ.six_to_quat(np.zeros((1, 6)))


@pytest.mark.parametrize("kind", KINDS)
def test_features(kind):
    g = load_golden(f"features_{syn.KIND_NAMES[kind]}.npz")
    lk = syn.kind_spec(kind)["lookup"]
    mine = np.stack([OF.parse_row(syn.KIND_NAMES[kind], r, lk) for r in g["rows"]])
    assert mine.dtype == g["xx"].dtype
    np.testing.assert_allclose(mine, g["xx"], rtol=0, atol=1e-12)


def test_fk_and_msg(body9):
    g = load_golden("fk.npz")
    for tname in OFK.TARGETS:
        for S in (1, 7, 100):
            est = OFK.arm_pose_from_nn_targets(g[f"{tname}__{S}__preds"], body9, tname)
            np.testing.assert_allclose(est, g[f"{tname}__{S}__est"], rtol=0, atol=1e-12)
            np.testing.assert_allclose(OFK.msg_from_est(est, body9, tname), g[f"{tname}__{S}__msg"], rtol=0, atol=1e-12)
    est = OFK.arm_pose_from_nn_targets(g["wide__preds"], body9, "ORI_CAL_LARM_UARM")
    np.testing.assert_allclose(OFK.msg_from_est(est, body9, "ORI_CAL_LARM_UARM"), g["wide__msg"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("kind", KINDS)
def test_lstm_masked_forward(kind):
    g = load_golden(f"lstm_{syn.KIND_NAMES[kind]}.npz")
    spec = syn.kind_spec(kind)
    state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], int(g["weight_seed"]))
    masks = list(unpack_masks(g))
    n = int(g["n"])
    np.testing.assert_allclose(OL.forward_with_masks(state, g["x"]), g["y_eval"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(OL.forward_with_masks(state, np.repeat(g["x"], n, 0), masks, spec["p"]), g["y_mc"],
                               rtol=0, atol=2e-6)
    # the restated cell equations against torch.nn.LSTM itself, eval mode
    import torch
    with torch.no_grad():
        y_t = OL.TorchDropoutLSTM.from_state(state, spec["p"])(torch.from_numpy(g["x"])).numpy()
    np.testing.assert_allclose(y_t, g["y_eval"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", ["watch_only_s3", "pocket_s1", "uarm_s1", "uarm_s4"])
def test_whole_path(name, body9):
    g = load_golden(f"e2e_{name}.npz")
    kind = {"watch_only": syn.KIND_WATCH_ONLY, "pocket": syn.KIND_POCKET, "uarm": syn.KIND_UARM}[name.rsplit("_", 1)[0]]
    spec = syn.kind_spec(kind)
    state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], int(g["weight_seed"]))
    masks = unpack_masks(g)
    orc = OE.OracleEstimator(syn.KIND_NAMES[kind], spec["lookup"], state, spec["stats"], spec["y_targets"].name,
                             spec["T"], int(g["smooth"]), int(g["n"]), body9, spec["p"],
                             mask_source=lambda f: list(masks[f]))
    mine = np.stack([np.asarray(orc.step(r), dtype=np.float64) for r in g["rows"]])
    assert mine.shape == g["msgs"].shape == (len(g["rows"]), 25 + 6 * int(g["n"]) * int(g["smooth"]))
    np.testing.assert_allclose(mine, g["msgs"], rtol=0, atol=5e-6)
    np.testing.assert_allclose(mine[:, :25], g["last"], rtol=0, atol=5e-6)
