"""Recorded-IMU CSV replay front-end: file formats on the CPU (against files written by the reference's own recorders),
relabelling on the GPU against the reference's golden messages (injected masks) and against the streaming path."""
from pathlib import Path

import numpy as np
import pytest

from conftest import load_golden, unpack_masks

from arm_pose_estimation_b200 import synthetic as syn
from arm_pose_estimation_b200.data_types import messaging
from arm_pose_estimation_b200.record import replay


def test_imu_csv_round_trip(tmp_path):
    rows = syn.synth_rows(syn.KIND_WATCH_ONLY, 1, 9, config_id=2)[0]
    p = tmp_path / "arm_pose_rec.csv"
    replay.write_imu_csv(p, rows)
    assert p.read_text().split("\n")[0] == ",".join(messaging.WATCH_ONLY_IMU_LOOKUP.keys())     # arm_pose_to_csv.py:11-12
    back, layout = replay.read_imu_csv(p)
    assert layout == messaging.LAYOUT_WATCH_ONLY
    np.testing.assert_array_equal(back, rows)                        # str(float32) round-trips exactly
    # the watch_raw_record.py variant: a leading timestamp column
    lines = p.read_text().strip().split("\n")
    q = tmp_path / "raw.csv"
    q.write_text("timestamp," + lines[0] + "\n" + "\n".join(f"2024-01-01 00:00:0{i}," + ln for i, ln in enumerate(lines[1:])) + "\n")
    np.testing.assert_array_equal(replay.read_imu_csv(q)[0], rows)
    rows55 = syn.synth_rows(syn.KIND_UARM, 1, 3, config_id=2)[0]
    replay.write_imu_csv(tmp_path / "wp.csv", rows55, messaging.LAYOUT_WATCH_PHONE)
    back55, layout55 = replay.read_imu_csv(tmp_path / "wp.csv")
    assert layout55 == messaging.LAYOUT_WATCH_PHONE
    np.testing.assert_array_equal(back55, rows55)
    (tmp_path / "bad.csv").write_text("a,b,c\n1,2,3\n")
    with pytest.raises(UserWarning):
        replay.read_imu_csv(tmp_path / "bad.csv")
    (tmp_path / "empty.csv").write_text(",".join(messaging.WATCH_ONLY_IMU_LOOKUP.keys()) + "\n")
    assert replay.read_imu_csv(tmp_path / "empty.csv")[0].shape == (0, 28)


def test_pose_csv_round_trip(tmp_path):
    msgs = np.random.default_rng(0).normal(size=(5, 25))
    p = tmp_path / "est.csv"
    replay.write_pose_csv(p, msgs, times=[f"t{i}" for i in range(5)])
    assert p.read_text().split("\n")[0].split(",") == replay.EST_OUTPUT_HEADER and len(replay.EST_OUTPUT_HEADER) == 26
    times, back = replay.read_pose_csv(p)
    assert times == [f"t{i}" for i in range(5)]
    np.testing.assert_array_equal(back, msgs)
    with pytest.raises(UserWarning):
        replay.write_pose_csv(tmp_path / "nope" / "est.csv", msgs)    # est_output.py:27-28


def test_pad_recordings():
    a, b = np.ones((3, 28), np.float32), 2 * np.ones((5, 28), np.float32)
    rows, lengths = replay.pad_recordings([a, b, np.zeros((0, 28), np.float32)])
    assert rows.shape == (3, 5, 28) and list(lengths) == [3, 5, 0]
    assert (rows[0, 3:] == 1).all() and (rows[2] == 0).all()


@pytest.mark.gpu
def test_relabel_matches_streaming(tmp_path):
    from arm_pose_estimation_b200 import _native as N
    from arm_pose_estimation_b200.estimate.batched import BatchedEstimator
    kind = syn.KIND_WATCH_ONLY
    spec = syn.kind_spec(kind)
    state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234)
    recs = [syn.synth_rows(kind, 1, F, config_id=6, first_stream=i)[0] for i, F in enumerate((11, 7, 16))]
    for i, r in enumerate(recs):                                          # through real files, in the reference's format
        replay.write_imu_csv(tmp_path / f"rec{i}.csv", r)
    recs = [replay.read_imu_csv(tmp_path / f"rec{i}.csv")[0] for i in range(3)]

    def make(n_streams, fpc):
        return BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"],
                                stats=spec["stats"], n_streams=n_streams, mc_samples=10, smooth=3, dropout=spec["p"],
                                frames_per_call=fpc, mask_mode=N.MASK_PHILOX, philox_seed=3)
    res = replay.relabel_recordings(recs, make, frames_per_call=5)
    stream = make(3, 1)
    rows, lengths = replay.pad_recordings(recs)
    for f in range(rows.shape[1]):
        out = stream.step(rows[:, f:f + 1])
        for i in range(3):
            if f < lengths[i]:
                np.testing.assert_array_equal(res[i]["msg"][f], out.msg[i, 0].astype(np.float64))
                np.testing.assert_array_equal(res[i]["std"][f], out.std[i, 0].astype(np.float64))
    assert [len(r["msg"]) for r in res] == [11, 7, 16]
    replay.write_pose_csv(tmp_path / "est0.csv", res[0]["msg"])
    np.testing.assert_array_equal(replay.read_pose_csv(tmp_path / "est0.csv")[1], res[0]["msg"])


GOLDEN = Path(__file__).resolve().parent / "golden"


def test_reference_written_files_parse_and_round_trip(tmp_path):
    # fixtures written by the reference's own arm_pose_to_csv / EstOutputRecorder (tests/golden/make_golden.py::golden_csv)
    g = load_golden("e2e_watch_only_s3.npz")
    rows, layout = replay.read_imu_csv(GOLDEN / "arm_pose_rec_watch_only_s3.csv")
    assert layout == messaging.LAYOUT_WATCH_ONLY
    np.testing.assert_array_equal(rows, g["rows"])
    replay.write_imu_csv(tmp_path / "mine.csv", rows)                    # our writer produces the reference recorder's bytes
    assert (tmp_path / "mine.csv").read_text() == (GOLDEN / "arm_pose_rec_watch_only_s3.csv").read_text()
    times, msgs = replay.read_pose_csv(GOLDEN / "est_output_watch_only_s3.csv")
    np.testing.assert_array_equal(msgs, g["last"])
    replay.write_pose_csv(tmp_path / "mine_est.csv", msgs, times=times)
    assert (tmp_path / "mine_est.csv").read_text() == (GOLDEN / "est_output_watch_only_s3.csv").read_text()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["fp32", "tc"])
def test_relabel_of_the_reference_recording_gives_the_reference_messages(tmp_path, variant):
    # the recording the REFERENCE wrote -> relabel_recordings (masks the reference drew, injected) -> the messages the reference
    # computed frame by frame (golden), and the pose file its EstOutputRecorder wrote
    from arm_pose_estimation_b200 import _native as N
    from arm_pose_estimation_b200.estimate.batched import BatchedEstimator
    from test_gpu_parity import msg_close, POS_TOL
    g = load_golden("e2e_watch_only_s3.npz")
    kind = syn.KIND_WATCH_ONLY
    spec = syn.kind_spec(kind)
    state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], int(g["weight_seed"]))
    n, smooth = int(g["n"]), int(g["smooth"])
    masks = unpack_masks(g)                                              # [F, L-1, T, n, H]
    rows, _ = replay.read_imu_csv(GOLDEN / "arm_pose_rec_watch_only_s3.csv")

    def make(n_streams, fpc):
        return BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"],
                                stats=spec["stats"], n_streams=n_streams, mc_samples=n, smooth=smooth, dropout=spec["p"],
                                frames_per_call=fpc, mask_mode=N.MASK_INJECTED, lstm_variant=variant)
    # two copies of the recording, one cut short: ragged recordings, calls of 4 + 4 + 4 + 2 frames
    res = replay.relabel_recordings([rows, rows[:9]], make, frames_per_call=4, keep_samples=True,
                                    masks=np.stack([masks, masks]))
    assert [len(r["msg"]) for r in res] == [len(rows), 9]
    worst = 0.0
    for r in res:
        F = len(r["msg"])
        worst = max(worst, msg_close(r["msg"], g["msgs"][:F, :25]))
        worst = max(worst, float(np.abs(r["samples"].reshape(F, -1) - g["msgs"][:F, 25:]).max()))
    print(f"relabelled reference recording [{variant}]: worst position error vs the reference's messages {worst:.3g} m")
    assert worst <= POS_TOL
    replay.write_pose_csv(tmp_path / "est.csv", res[0]["msg"])
    msg_close(replay.read_pose_csv(tmp_path / "est.csv")[1], replay.read_pose_csv(GOLDEN / "est_output_watch_only_s3.csv")[1])
