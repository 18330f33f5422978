"""Multi-stream front-end: host logic on CPU with a stand-in engine; on the GPU every stream must see exactly
the estimates a dedicated single-stream estimator produces from the rows it kept."""
import queue
import time

import numpy as np
import pytest

from arm_pose_estimation_b200.estimate.multi_stream import MultiStreamEstimator


class FakeOut:
    def __init__(self, msg, std, samples, status):
        self.msg, self.std, self.samples, self.status = msg, std, samples, status


class FakeEngine:
    """Records what the front-end asks for; message[0] = first column of the row, message[1] = the stream's frame number."""
    def __init__(self, B, ncols=28, S=3):
        self.B, self.ncols, self.S, self.calls = B, ncols, S, []

    def step(self, rows, stream_frames=None):
        self.calls.append((rows.copy(), stream_frames.copy()))
        msg = np.zeros((self.B, 1, 25), np.float32)
        msg[:, 0, 0] = rows[:, 0, 0]
        msg[:, 0, 1] = stream_frames
        return FakeOut(msg, np.zeros((self.B, 1, 6), np.float32), np.ones((self.B, 1, self.S, 6), np.float32),
                       np.zeros((self.B, 1), np.int32))

    step_graph = step                                       # the front-end ticks through the engine's CUDA-graph entry point


def test_collect_applies_the_reference_freshness_policy_per_stream():
    eng = FakeEngine(3)
    ms = MultiStreamEstimator(eng, max_backlog=5)
    qs = [queue.Queue() for _ in range(3)]
    for k in range(9):                                      # stream 0: 9 rows waiting -> rows 0..2 are shed, row 3 is taken, 5 stay
        qs[0].put(np.full(28, k, np.float32))
    qs[1].put(np.full(28, 100, np.float32))                 # stream 1: one row; stream 2: nothing
    active = ms.collect(qs)
    assert active.tolist() == [True, True, False]
    assert ms._rows[0, 0, 0] == 3 and qs[0].qsize() == 5 and ms.dropped.tolist() == [3, 0, 0]
    assert ms._rows[1, 0, 0] == 100 and qs[1].qsize() == 0


def test_streams_advance_independently_and_inactive_streams_are_skipped():
    eng = FakeEngine(3)
    ms = MultiStreamEstimator(eng, add_mc_samples=False)
    qs = [queue.Queue() for _ in range(3)]
    qs[0].put(np.full(28, 1, np.float32)); qs[2].put(np.full(28, 7, np.float32))
    msgs = ms.tick(ms.collect(qs))
    assert sorted(msgs) == [0, 2] and eng.calls[-1][1].tolist() == [0, -1, 0]
    assert isinstance(msgs[0], np.ndarray) and msgs[0].shape == (25,) and msgs[0].dtype == np.float64
    qs[0].put(np.full(28, 2, np.float32)); qs[1].put(np.full(28, 5, np.float32))
    msgs = ms.tick(ms.collect(qs))
    assert sorted(msgs) == [0, 1] and eng.calls[-1][1].tolist() == [1, 0, -1]
    assert ms.frames.tolist() == [2, 1, 1] and msgs[0][1] == 1 and msgs[1][1] == 0
    assert ms.tick(ms.collect(qs)) == {} and len(eng.calls) == 2           # nothing waiting: no launch
    ms.reset_stream(0)
    qs[0].put(np.full(28, 3, np.float32))
    ms.tick(ms.collect(qs))
    assert eng.calls[-1][1].tolist() == [0, -1, -1]


def test_thread_fan_out_and_message_format_with_samples():
    eng = FakeEngine(2, S=3)
    ms = MultiStreamEstimator(eng, add_mc_samples=True)
    qs = [queue.Queue() for _ in range(2)]
    out_qs = ms.process_in_thread(qs)
    for k in range(4):
        qs[k % 2].put(np.full(28, k, np.float32))
        time.sleep(0.01)
    got = [out_qs[0].get(timeout=5), out_qs[1].get(timeout=5), out_qs[0].get(timeout=5), out_qs[1].get(timeout=5)]
    ms.terminate()
    ms._thread.join(timeout=5)
    assert not ms.is_active() and all(isinstance(m, list) and len(m) == 25 + 6 * 3 for m in got)      # estimator.py:131-136
    assert [m[0] for m in got] == [0.0, 1.0, 2.0, 3.0] and [m[1] for m in got] == [0.0, 0.0, 1.0, 1.0]
    import struct
    struct.pack("f" * len(got[0]), *got[0])                 # what PoseEstPublisherUDP does with it (pose_est_udp.py:47)
    with pytest.raises(UserWarning):
        ms.process_in_thread(qs[:1])


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["fp32", "tc"])
def test_every_stream_matches_its_own_single_stream_estimator(variant):
    from arm_pose_estimation_b200 import _native as N, synthetic as syn
    from arm_pose_estimation_b200.estimate.batched import BatchedEstimator
    kind, B, n, smooth, F = syn.KIND_POCKET, 4, 20, 3, 9
    spec = syn.kind_spec(kind)
    state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)

    def engine(n_streams, first_stream):
        return BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"],
                                stats=spec["stats"], n_streams=n_streams, mc_samples=n, smooth=smooth, dropout=spec["p"],
                                frames_per_call=1, mask_mode=N.MASK_PHILOX, philox_seed=11, first_stream=first_stream,
                                lstm_variant=variant)

    rows = syn.synth_rows(kind, B, F, config_id=17)
    # arrival pattern: which streams have a row at each tick (stream 3 joins late, stream 1 stalls in the middle)
    rng = np.random.default_rng(5)
    ms = MultiStreamEstimator(engine(B, 0), add_mc_samples=True)
    qs = [queue.Queue() for _ in range(B)]
    fed = [0] * B
    got = [[] for _ in range(B)]
    for tick in range(40):
        for b in range(B):
            arrives = rng.random() < (0.8, 0.4, 0.9, 0.6)[b] and not (b == 3 and tick < 6) and not (b == 1 and 10 <= tick < 20)
            if arrives and fed[b] < F:
                qs[b].put(rows[b, fed[b]])
                fed[b] += 1
        for b, m in ms.tick(ms.collect(qs)).items():
            got[b].append(np.asarray(m))
    assert [len(g) for g in got] == fed and min(fed) >= 5
    for b in range(B):
        solo = engine(1, b)                                 # a dedicated estimator for global stream id b
        for f in range(fed[b]):
            out = solo.step(rows[b:b + 1, f:f + 1])
            want = np.concatenate([out.msg[0, 0].astype(np.float64), out.samples[0, 0].astype(np.float64).ravel()])
            np.testing.assert_array_equal(got[b][f], want)
    # ... and the ORACLE's single-stream loop (estimator.py:155-178 restated), fed the masks the kernels drew for that stream
    # (ape_philox_masks exports them: keyed by global stream id and the stream's own frame counter)
    import torch
    from oracle import estimator as OE
    L, T, H = spec["L"], spec["T"], spec["H"]
    worst = 0.0
    for b in range(B):
        masks = torch.empty((1, fed[b], L - 1, T, n, H), dtype=torch.uint8, device="cuda")
        N.check(N.load().ape_philox_masks(11, b, 1, fed[b], 0, L, T, n, H, spec["p"], N.ptr(masks), N.current_stream_ptr()), "masks")
        masks = masks.cpu().numpy()[0]
        orc = OE.OracleEstimator(syn.KIND_NAMES[kind], spec["lookup"], state, spec["stats"], spec["y_targets"].name,
                                 T, smooth, n, None, spec["p"], mask_source=lambda f, m=masks: list(m[f]))
        for f in range(fed[b]):
            want = np.asarray(orc.step(rows[b, f]), dtype=np.float64)
            for a, c in ((4, 7), (11, 14), (18, 21)):                          # hand, elbow, shoulder positions
                worst = max(worst, float(np.abs(got[b][f][a:c] - want[a:c]).max()))
            worst = max(worst, float(np.abs(got[b][f][25:] - want[25:]).max()))
    print(f"multi-stream front-end [{variant}] vs the oracle's per-stream loop: worst position error {worst:.3g} m")
    assert worst <= 1e-4
