"""Generate the golden fixtures by running the UNMODIFIED reference in this container.

Run HERE only (``python tests/golden/make_golden.py``): imports ``wear_mocap_ape`` from /root/reference/src with
the three harness shims of SURVEY.md §8c (an ``aenum`` stand-in, a redirected ``config.PATHS["deploy"]``, seeded
synthetic ``checkpoint.pt`` files) and writes small ``.npz`` fixtures next to this file.  While doing so it checks
the oracle (``oracle/``) against the reference's outputs and aborts on any mismatch, so a committed fixture set
means "oracle == reference on these inputs".  The GPU box never runs this file; it reads the ``.npz`` only.
"""
import enum
import json
import shutil
import sys
import tempfile
import types
import warnings
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
warnings.filterwarnings("ignore")


def _install_aenum_shim():
    class _Dict(enum._EnumDict):
        def __setitem__(self, k, v):
            if k != "_settings_":
                super().__setitem__(k, v)

    class _Meta(enum.EnumMeta):
        @classmethod
        def __prepare__(m, name, bases, **kw):
            base = super().__prepare__(name, bases, **kw)
            d = _Dict()
            d.__dict__.update(base.__dict__)
            for k in base:
                dict.__setitem__(d, k, base[k])
            return d

    class Enum(enum.Enum, metaclass=_Meta):
        pass

    mod = types.ModuleType("aenum")
    mod.Enum, mod.NoAlias = Enum, object()
    sys.modules["aenum"] = mod


_install_aenum_shim()
sys.path.insert(0, "/root/reference/src")

from wear_mocap_ape import config as ref_config                      # noqa: E402
from wear_mocap_ape.data_deploy.nn import deploy_models as ref_dm    # noqa: E402
from wear_mocap_ape.data_types import messaging as ref_msg           # noqa: E402
from wear_mocap_ape.estimate import compose_msg as ref_cm            # noqa: E402
from wear_mocap_ape.estimate import estimate_joints as ref_ej        # noqa: E402
from wear_mocap_ape.estimate import nn_models as ref_nn              # noqa: E402
from wear_mocap_ape.estimate.watch_only import WatchOnlyNN           # noqa: E402
from wear_mocap_ape.estimate.watch_phone_pocket_nn import WatchPhonePocketNN   # noqa: E402
from wear_mocap_ape.estimate.watch_phone_uarm_nn import WatchPhoneUarmNN       # noqa: E402
from wear_mocap_ape.utility import transformations as ts             # noqa: E402
from wear_mocap_ape.utility.names import NNS_INPUTS as REF_IN, NNS_TARGETS as REF_TG   # noqa: E402

from arm_pose_estimation_b200 import synthetic as syn                # noqa: E402
from arm_pose_estimation_b200.data_types import messaging as my_msg  # noqa: E402
from arm_pose_estimation_b200.utility import names as my_names       # noqa: E402
from oracle import estimator as OE, features as OF, fk as OFK, lstm as OL, quat as OQ   # noqa: E402

BODY = np.array([[-0.22, 0, 0, -0.26, 0, 0, -0.1704612, 0.4309841, -0.00670862]])
KINDS = (syn.KIND_WATCH_ONLY, syn.KIND_POCKET, syn.KIND_UARM)


def close(a, b, tol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = float(np.max(np.abs(a - b))) if a.size else 0.0
    assert a.shape == b.shape and err <= tol, f"{what}: shape {a.shape} vs {b.shape}, max err {err:g} > {tol:g}"
    print(f"  ok {what}: max |d| = {err:.3g}")


def make_ref_deploy(tmp):
    """Copy of the reference deploy dir + seeded synthetic checkpoints (same weights the package synthesises)."""
    dst = Path(tmp) / "deploy"
    shutil.copytree(ref_config.PATHS["deploy"], dst)
    for kind in KINDS:
        h = syn.KIND_HASH[kind]
        params = json.loads((dst / "nn" / h / "results.json").read_text())
        sd = syn.synth_state_dict(len(params["x_inputs_v"]), params["hidden_layer_size"],
                                  params["hidden_layer_count"], len(params["y_targets_v"]), 1234 + kind)
        (dst / "nn" / h).chmod(0o755)
        torch.save(({k: torch.from_numpy(v) for k, v in sd.items()}, None), dst / "nn" / h / "checkpoint.pt")
    ref_config.PATHS["deploy"] = dst
    return dst


def ref_estimator(kind, **kw):
    if kind == syn.KIND_WATCH_ONLY:
        return WatchOnlyNN(**kw)
    if kind == syn.KIND_POCKET:
        return WatchPhonePocketNN(model_hash=ref_dm.LSTM.WATCH_PHONE_POCKET.value, **kw)
    return WatchPhoneUarmNN(**kw)


def golden_tables():
    print("tables")
    assert dict(ref_msg.WATCH_ONLY_IMU_LOOKUP) == dict(my_msg.WATCH_ONLY_IMU_LOOKUP)
    assert dict(ref_msg.WATCH_PHONE_IMU_LOOKUP) == dict(my_msg.WATCH_PHONE_IMU_LOOKUP)
    for mem in my_names.NNS_INPUTS:
        assert list(REF_IN[mem.name].value) == list(mem.value), mem.name
    for mem in my_names.NNS_TARGETS:
        assert list(REF_TG[mem.name].value) == list(mem.value), mem.name
    tables = {
        "watch_only_lookup": dict(ref_msg.WATCH_ONLY_IMU_LOOKUP), "watch_phone_lookup": dict(ref_msg.WATCH_PHONE_IMU_LOOKUP),
        "inputs": {m.name: list(REF_IN[m.name].value) for m in my_names.NNS_INPUTS},
        "targets": {m.name: list(REF_TG[m.name].value) for m in my_names.NNS_TARGETS},
    }
    (HERE / "tables.json").write_text(json.dumps(tables, indent=1) + "\n")


def golden_quat():
    print("quaternion leaves (transformations.py)")
    rng = np.random.default_rng(7)
    n = 48
    a, b = rng.normal(size=(n, 4)), rng.normal(size=(n, 4))
    qa = a / np.linalg.norm(a, axis=1, keepdims=True)
    v, e = rng.normal(size=(n, 3)), rng.uniform(-3, 3, size=(n, 3))
    six = rng.normal(size=(n, 6))
    s_c = rng.normal(size=(n, 2))
    near = qa[0] + 0.2 * rng.normal(size=(40, 4))
    near[::3] *= -1.0                                                    # exercise the hemisphere flip
    out = dict(
        a=a, b=b, qa=qa, v=v, e=e, six=six, s_c=s_c, near=near,
        hamilton=ts.hamilton_product(a, b),
        rotate=ts.quat_rotate_vector(a, v),
        rotate_single_vec=ts.quat_rotate_vector(a, v[0]),
        invert=ts.quat_invert(a),
        euler=ts.euler_to_quat(e),
        a2g_no_north=ts.android_quat_to_global_no_north(a),
        a2g=ts.android_quat_to_global(a, b),
        y_rot=ts.reduce_global_quat_to_y_rot(qa),
        north_left=ts.calib_watch_left_to_north_quat(qa),
        rot9=ts.quat_to_rot_mat_1x9(a),
        six_of_q=ts.quat_to_6drr_1x6(a),
        rot9_of_six=ts.six_drr_1x6_to_rot_mat_1x9(six),
        quat_of_six=ts.six_drr_1x6_to_quat(six),
        hips=ts.hips_sin_cos_to_quat(s_c[:, 0], s_c[:, 1]),
        average=ts.average_quaternions(near),
    )
    close(OQ.hamilton(a, b), out["hamilton"], 1e-14, "hamilton")
    close(OQ.rotate(a, v), out["rotate"], 1e-13, "rotate")
    close(OQ.rotate(a, v[0]), out["rotate_single_vec"], 1e-13, "rotate single vec")
    close(OQ.invert(a), out["invert"], 1e-14, "invert")
    close(OQ.euler_to_quat(e), out["euler"], 1e-15, "euler_to_quat")
    close(OQ.android_to_global_no_north(a), out["a2g_no_north"], 0, "android no north")
    close(OQ.android_to_global(a, b), out["a2g"], 1e-14, "android_to_global")
    close(OQ.y_rot_of(qa), out["y_rot"], 1e-14, "reduce_global_quat_to_y_rot")
    close(OQ.north_quat_left_arm(qa), out["north_left"], 1e-14, "calib_watch_left_to_north_quat")
    close(OQ.quat_to_rot9(a), out["rot9"], 1e-14, "quat_to_rot_mat_1x9")
    close(OQ.quat_to_six(a), out["six_of_q"], 1e-14, "quat_to_6drr_1x6")
    close(OQ.six_to_rot9(six), out["rot9_of_six"], 1e-13, "six_drr_1x6_to_rot_mat_1x9")
    close(OQ.six_to_quat(six), out["quat_of_six"], 1e-13, "six_drr_1x6_to_quat")
    close(OQ.hips_sin_cos_to_quat(s_c[:, 0], s_c[:, 1]), out["hips"], 1e-15, "hips_sin_cos_to_quat")
    close(OQ.average_quats(near), out["average"], 1e-15, "average_quaternions")
    np.savez_compressed(HERE / "quat_leaves.npz", **out)


def golden_features():
    print("stage 1: parse_row_to_xx")
    for kind in KINDS:
        est = ref_estimator(kind, smooth=1, monte_carlo_samples=2)
        rows = syn.synth_rows(kind, 4, 16, config_id=90 + kind).reshape(64, -1)
        rng = np.random.default_rng(kind)
        rows[1::7, :] *= rng.uniform(0.5, 2.0)                           # non-unit quaternions, odd magnitudes
        xx = np.stack([np.asarray(est.parse_row_to_xx(r)) for r in rows])
        lk = my_msg.WATCH_ONLY_IMU_LOOKUP if kind == syn.KIND_WATCH_ONLY else my_msg.WATCH_PHONE_IMU_LOOKUP
        mine = np.stack([OF.parse_row(syn.KIND_NAMES[kind], r, lk) for r in rows])
        assert mine.dtype == xx.dtype, (mine.dtype, xx.dtype)
        close(mine, xx, 1e-12 if xx.dtype == np.float64 else 0, f"features {syn.KIND_NAMES[kind]} ({xx.dtype})")
        np.savez_compressed(HERE / f"features_{syn.KIND_NAMES[kind]}.npz", rows=rows, xx=xx)


def golden_fk():
    print("stage 3: arm_pose_from_nn_targets + msg_from_nn_targets_est")
    rng = np.random.default_rng(11)
    out = {}
    for tname, O in (("ORI_CAL_LARM_UARM", 12), ("ORI_CAL_LARM_UARM_HIPS", 14), ("ORI_POS_CAL_LARM_UARM_HIPS", 20)):
        for S in (1, 7, 100):
            base = rng.normal(size=(1, O))
            preds = base + (0.05 if S > 1 else 0.0) * rng.normal(size=(S, O))   # MC-like spread round one pose
            est = ref_ej.arm_pose_from_nn_targets(preds, BODY, REF_TG[tname])
            msg = ref_cm.msg_from_nn_targets_est(est, BODY, REF_TG[tname])
            my_est = OFK.arm_pose_from_nn_targets(preds, BODY, tname)
            close(my_est, est, 1e-12, f"est {tname} S={S}")
            close(OFK.msg_from_est(my_est, BODY, tname), msg, 1e-12, f"msg {tname} S={S}")
            out[f"{tname}__{S}__preds"], out[f"{tname}__{S}__est"], out[f"{tname}__{S}__msg"] = preds, est, msg
    # wide-spread samples: quaternions far apart, sign flips in the average
    preds = rng.normal(size=(64, 12))
    est = ref_ej.arm_pose_from_nn_targets(preds, BODY, REF_TG["ORI_CAL_LARM_UARM"])
    msg = ref_cm.msg_from_nn_targets_est(est, BODY, REF_TG["ORI_CAL_LARM_UARM"])
    close(OFK.msg_from_est(OFK.arm_pose_from_nn_targets(preds, BODY, "ORI_CAL_LARM_UARM"), BODY, "ORI_CAL_LARM_UARM"),
          msg, 1e-12, "msg wide spread")
    out["wide__preds"], out["wide__est"], out["wide__msg"] = preds, est, msg
    out["body"] = BODY
    np.savez_compressed(HERE / "fk.npz", **out)


def golden_lstm():
    print("stage 2: DropoutLSTM.forward / monte_carlo_predictions with replayed masks")
    for kind in KINDS:
        spec = syn.kind_spec(kind)
        model, params = ref_nn.load_deployed_model_from_hash(spec["hash"])
        I, H, L, T, O, p = (spec[k] for k in "IHLTOp")
        n = 100 if kind == syn.KIND_UARM else 24
        rng = np.random.default_rng(50 + kind)
        x = rng.normal(size=(1, T, I)).astype(np.float32)
        with torch.no_grad():
            y_eval = model(torch.from_numpy(x)).numpy()
            seed = 4242 + kind
            torch.manual_seed(seed)
            y_mc = model.monte_carlo_predictions(n_samples=n, x=torch.from_numpy(x)).numpy()
        model.eval()
        masks = OL.replay_torch_masks(seed, T, n, H, L, p)
        state = syn.synth_state_dict(I, H, L, O, 1234 + kind)
        close(OL.forward_with_masks(state, x), y_eval, 2e-6, f"lstm eval {syn.KIND_NAMES[kind]}")
        close(OL.forward_with_masks(state, np.repeat(x, n, 0), masks, p), y_mc, 2e-6, f"lstm mc n={n} {syn.KIND_NAMES[kind]}")
        np.savez_compressed(HERE / f"lstm_{syn.KIND_NAMES[kind]}.npz", x=x, y_eval=y_eval, y_mc=y_mc, n=n, seed=seed,
                            weight_seed=1234 + kind, masks=np.packbits(np.stack(masks)), masks_shape=np.stack(masks).shape)


def golden_e2e():
    print("whole path: the three calls of estimator.py:174-176, frame by frame")
    F_, n = 14, 20
    for kind, smooth in ((syn.KIND_WATCH_ONLY, 3), (syn.KIND_POCKET, 1), (syn.KIND_UARM, 1), (syn.KIND_UARM, 4)):
        spec = syn.kind_spec(kind)
        I, H, L, T, O, p = (spec[k] for k in "IHLTOp")
        est = ref_estimator(kind, smooth=smooth, monte_carlo_samples=n, add_mc_samples=True)
        est.reset()
        rows = syn.synth_rows(kind, 1, F_, config_id=70 + kind)[0]
        seed0 = 9000 + 10 * kind + smooth
        msgs, masks, last = [], [], []
        for f, row in enumerate(rows):
            xx = est.parse_row_to_xx(row)
            torch.manual_seed(seed0 + f)
            pred = est.add_xx_to_row_hist_and_make_prediction(xx)
            msgs.append(np.asarray(est.msg_from_pred(pred, True), dtype=np.float64))
            last.append(est.get_last_msg())
            masks.append(np.stack(OL.replay_torch_masks(seed0 + f, T, n, H, L, p)))
        msgs, masks = np.stack(msgs), np.stack(masks)
        # the oracle, fed the replayed masks, must reproduce the reference frame by frame
        state = syn.synth_state_dict(I, H, L, O, 1234 + kind)
        orc = OE.OracleEstimator(syn.KIND_NAMES[kind], spec["lookup"], state, spec["stats"], spec["y_targets"].name,
                                 T, smooth, n, BODY, p, mask_source=lambda f: list(masks[f]))
        mine = np.stack([np.asarray(orc.step(r), dtype=np.float64) for r in rows])
        close(mine, msgs, 5e-6, f"e2e {syn.KIND_NAMES[kind]} smooth={smooth}")
        # and with torch's own RNG (the mode used as the timed CPU baseline) bit-for-bit the same stream of masks
        orc_t = OE.OracleEstimator(syn.KIND_NAMES[kind], spec["lookup"], state, spec["stats"], spec["y_targets"].name,
                                   T, smooth, n, BODY, p, mask_source="torch")
        outs = []
        for f, r in enumerate(rows):
            xx = orc_t.parse_row_to_xx(r)
            torch.manual_seed(seed0 + f)
            outs.append(np.asarray(orc_t.msg_from_pred(orc_t.add_xx_to_row_hist_and_make_prediction(xx)), dtype=np.float64))
        close(np.stack(outs), msgs, 1e-9, f"e2e torch-RNG {syn.KIND_NAMES[kind]} smooth={smooth}")
        np.savez_compressed(HERE / f"e2e_{syn.KIND_NAMES[kind]}_s{smooth}.npz", rows=rows, msgs=msgs, last=np.stack(last),
                            masks=np.packbits(masks), masks_shape=masks.shape, n=n, smooth=smooth, seed0=seed0,
                            weight_seed=1234 + kind)


def golden_ff():
    print("feed-forward MC regressors: DropoutFF / DropoutFF2D with the dropout mask torch drew")
    from oracle import ff as OFF
    out = {}
    for name, ctor, x_shape in (("ff", lambda: ref_nn.DropoutFF(output_size=12, hidden_layer_size=96, hidden_layer_count=2, input_size=20, dropout=0.2), (1, 20)),
                                ("ff2d", lambda: ref_nn.DropoutFF2D(output_size=14, hidden_layer_size=64, hidden_layer_count=1, input_size=22, seq_len=5, dropout=0.3), (1, 5, 22))):
        torch.manual_seed(77)
        model = ctor().eval()
        p = model._do.p
        state = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
        rng = np.random.default_rng(5)
        x = rng.normal(size=x_shape).astype(np.float32)
        xb = rng.normal(size=(6,) + x_shape[1:]).astype(np.float32)
        n = 40
        H = state["_input_layer.weight"].shape[0]
        with torch.no_grad():
            y_eval = model(torch.from_numpy(xb)).numpy()
            torch.manual_seed(4321)
            y_mc = model.monte_carlo_predictions(n_samples=n, x=torch.from_numpy(x)).numpy()
            # the mask torch drew: nn.Dropout on the [n, 1, H] activations, replayed from the same seed
            torch.manual_seed(4321)
            masks = torch.empty(n, 1, H).bernoulli_(1 - p).numpy().astype(np.uint8).reshape(n, H)
        model.eval()
        close(OFF.forward_with_masks(state, xb.reshape(6, -1)), y_eval.reshape(6, -1), 2e-6, f"{name} eval")
        close(OFF.forward_with_masks(state, np.repeat(x.reshape(1, -1), n, 0), masks, p), y_mc.reshape(n, -1), 2e-6, f"{name} mc n={n}")
        out.update({f"{name}__x": x, f"{name}__xb": xb, f"{name}__y_eval": y_eval, f"{name}__y_mc": y_mc, f"{name}__masks": masks,
                    f"{name}__p": np.float32(p)})
        out.update({f"{name}__state__{k}": v for k, v in state.items()})
    np.savez_compressed(HERE / "ff.npz", **out)


def golden_imu_pose_lstm():
    print("ImuPoseLSTM: Linear + relu -> LSTM(256, 256, 2) -> Linear, eval mode")
    from oracle import imu_pose_lstm as OI
    model = ref_nn.ImuPoseLSTM(input_size=20, hidden_layer_size=0, hidden_layer_count=0, output_size=12).eval()
    state = syn.synth_imu_pose_state_dict(20, 12, 99)           # (1 M weights: the fixture carries the seed, not the values)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    x = np.random.default_rng(6).normal(size=(5, 7, 20)).astype(np.float32)
    with torch.no_grad():
        y = model(torch.from_numpy(x)).numpy()
        y_mc = model.monte_carlo_predictions(n_samples=9, x=torch.from_numpy(x[:1])).numpy()
    close(OI.forward(state, x), y, 2e-6, "ImuPoseLSTM eval")
    close(y_mc, y[:1], 1e-6, "ImuPoseLSTM monte_carlo_predictions == forward (up to batch-size dependent rounding)")
    np.savez_compressed(HERE / "imu_pose_lstm.npz", x=x, y=y, y_mc=y_mc, weight_seed=99)


def golden_csv():
    print("file formats: the reference's own recorders write the watch-only golden rows / messages")
    import queue
    import threading
    import time
    from wear_mocap_ape.record.arm_pose_to_csv import arm_pose_to_csv
    from wear_mocap_ape.record.est_output import EstOutputRecorder
    from arm_pose_estimation_b200.record import replay
    g = np.load(HERE / "e2e_watch_only_s3.npz")
    rows, last = g["rows"], g["last"]
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        # raw IMU recording (record/arm_pose_to_csv.py:9-32): rows arrive as tuples of Python floats (struct.unpack in the
        # listener); the recorder loops forever, so a non-iterable sentinel ends its thread and closes the file
        q = queue.Queue()
        th = threading.Thread(target=arm_pose_to_csv, args=(q, tmp), daemon=True)
        th.start()
        for r in rows:
            q.put(tuple(float(v) for v in r))
        q.put(None)                                                   # (skipped by the recorder: arm_pose_to_csv.py:22)
        q.put(0)                                                      # map(str, 0) raises -> the with-block closes the file
        th.join(timeout=20)
        (imu_file,) = list(tmp.glob("arm_pose_rec_*.csv"))
        imu_text = imu_file.read_text()
        # pose estimates (record/est_output.py:10-57)
        rec = EstOutputRecorder(tmp / "est.csv")
        mq = queue.Queue()
        rec.record_in_thread(mq)
        for m in last:
            mq.put(np.asarray(m))
        while not mq.empty():
            time.sleep(0.05)
        time.sleep(0.2)
        rec.terminate()
        time.sleep(2.5)                                               # (its queue.get times out after 2 s, then the loop ends)
        est_text = (tmp / "est.csv").read_text()
        # the package's reader / writer against them
        back, layout = replay.read_imu_csv(imu_file)
        assert layout == my_msg.LAYOUT_WATCH_ONLY
        close(back, rows, 0, "read_imu_csv(reference recording)")
        replay.write_imu_csv(tmp / "mine.csv", rows)
        assert (tmp / "mine.csv").read_text() == imu_text, "write_imu_csv differs from the reference recorder's bytes"
        times, msgs = replay.read_pose_csv(tmp / "est.csv")
        close(msgs, last, 0, "read_pose_csv(reference EstOutputRecorder file)")
        replay.write_pose_csv(tmp / "mine_est.csv", last, times=times)
        assert (tmp / "mine_est.csv").read_text() == est_text, "write_pose_csv differs from EstOutputRecorder's bytes"
    (HERE / "arm_pose_rec_watch_only_s3.csv").write_text(imu_text)
    (HERE / "est_output_watch_only_s3.csv").write_text(est_text)
    print("  ok csv fixtures written")


def main():
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    with tempfile.TemporaryDirectory() as tmp:
        make_ref_deploy(tmp)
        if only == "ff":
            golden_ff()
            return
        if only == "imu_pose_lstm":
            golden_imu_pose_lstm()
            return
        if only == "csv":
            golden_csv()
            return
        golden_tables()
        golden_quat()
        golden_features()
        golden_fk()
        golden_lstm()
        golden_e2e()
        golden_ff()
        golden_imu_pose_lstm()
        golden_csv()
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
