"""GPU (-m gpu): the small-batch cluster kernel (csrc/ape_lstm_tcl.cu: all layers of a call in one launch, one 8-CTA cluster per 128
rows, hidden units split across the cluster) - what a single-stream estimator runs per frame (BASELINE configs[1]) and what a few
dozen real-time streams run per tick.
It keeps the layer kernels' operand rounding points, accumulation order and Philox keys, so it must be BIT-identical to them;
and like them it is checked against the reference's own messages and the oracle."""
import numpy as np
import pytest

from conftest import load_golden, unpack_masks
from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200 import synthetic as syn
from oracle import estimator as OE
from test_gpu_parity import msg_close, POS_TOL
from test_gpu_tc import make

pytestmark = pytest.mark.gpu
KINDS = [syn.KIND_WATCH_ONLY, syn.KIND_POCKET, syn.KIND_UARM]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("B,n,nF,mode", [(1, 100, 1, N.MASK_PHILOX), (1, 1, 1, N.MASK_PHILOX), (3, 40, 1, N.MASK_PHILOX), (2, 16, 4, N.MASK_INJECTED),
                                         (1, 128, 1, N.MASK_INJECTED),
                                         # several clusters in one launch (one per 128 rows): estimates that straddle two clusters, more
                                         # estimates than a cluster has rows, several frames per call, the largest call the estimator gives it
                                         (3, 100, 1, N.MASK_PHILOX), (130, 1, 1, N.MASK_PHILOX), (20, 7, 2, N.MASK_INJECTED), (16, 128, 1, N.MASK_PHILOX)])
def test_small_batch_kernel_is_bit_identical_to_the_layer_kernels(kind, B, n, nF, mode):
    rows = syn.synth_rows(kind, B, 3 * nF, config_id=6)
    kw = dict(frames_per_call=nF, mask_mode=mode, philox_seed=11, smooth=2)
    small, spec, _ = make(kind, B, n, "tc", **kw)
    layers, _, _ = make(kind, B, n, "tc", small_batch_kernel=False, **kw)
    assert small.small_batch and small.tc_flags == 4 and not layers.small_batch
    rng = np.random.default_rng(2)
    for c in range(3):
        masks = None
        if mode == N.MASK_INJECTED:
            masks = (rng.random(size=(B, nF, spec["L"] - 1, spec["T"], n, spec["H"])) < 0.8).astype(np.uint8)
        a = small.step(rows[:, c * nF:(c + 1) * nF], masks=masks)
        b = layers.step(rows[:, c * nF:(c + 1) * nF], masks=masks)
        assert np.isfinite(a.msg).all()
        np.testing.assert_array_equal(a.samples, b.samples)
        np.testing.assert_array_equal(a.msg, b.msg)
        np.testing.assert_array_equal(a.std, b.std)
    assert small.launches == 3 * 3                                            # stage 1, ONE LSTM launch, stage 3 per call


@pytest.mark.parametrize("name", ["watch_only_s3", "pocket_s1", "uarm_s1", "uarm_s4"])
def test_small_batch_whole_path_against_reference_messages(name):
    g = load_golden(f"e2e_{name}.npz")
    kind = {"watch_only": syn.KIND_WATCH_ONLY, "pocket": syn.KIND_POCKET, "uarm": syn.KIND_UARM}[name.rsplit("_", 1)[0]]
    n, smooth = int(g["n"]), int(g["smooth"])
    masks = unpack_masks(g)
    rows, F = g["rows"], len(g["rows"])
    be, spec, _ = make(kind, 1, n, "tc", smooth=smooth, frames_per_call=1, mask_mode=N.MASK_INJECTED)
    assert be.small_batch
    worst = 0.0
    for f in range(F):                                                        # frame by frame: the streaming call pattern
        out = be.step(rows[None, f:f + 1], masks=masks[None, f:f + 1])
        worst = max(worst, msg_close(out.msg[0, 0], g["msgs"][f, :25]))
        err = float(np.abs(out.samples[0, 0] - g["msgs"][f, 25:].reshape(n * smooth, 6)).max())
        assert err <= POS_TOL
        worst = max(worst, err)
    print(f"{name} small-batch kernel: worst position error vs the reference's messages over {F} frames: {worst:.3g} m")


def test_small_batch_graph_path_equals_eager_and_ragged_rows_against_oracle():
    kind, B, n = syn.KIND_POCKET, 1, 37                                       # 37 rows of a 128-row tile
    spec = syn.kind_spec(kind)
    rows = syn.synth_rows(kind, B, 5, config_id=12)
    graph, _, state = make(kind, B, n, "tc", mask_mode=N.MASK_PHILOX, philox_seed=4)
    eager, _, _ = make(kind, B, n, "tc", mask_mode=N.MASK_PHILOX, philox_seed=4)
    assert graph.small_batch
    import torch
    for f in range(5):
        a, b = graph.step_graph(rows[:, f:f + 1]), eager.step(rows[:, f:f + 1])
        np.testing.assert_array_equal(a.msg, b.msg)
        np.testing.assert_array_equal(a.samples, b.samples)
    # the Philox masks exported and replayed through the oracle
    masks = torch.empty((B, 1, spec["L"] - 1, spec["T"], n, spec["H"]), dtype=torch.uint8, device="cuda")
    orc = None
    ref, _, _ = make(kind, B, n, "tc", mask_mode=N.MASK_PHILOX, philox_seed=4)
    got = []
    ms = []
    for f in range(3):
        N.check(N.load().ape_philox_masks(4, 0, B, 1, f, spec["L"], spec["T"], n, spec["H"], spec["p"], N.ptr(masks), N.current_stream_ptr()), "masks")
        ms.append(masks.cpu().numpy()[0, 0].copy())
        got.append(ref.step(rows[:, f:f + 1]))
    orc = OE.OracleEstimator(syn.KIND_NAMES[kind], spec["lookup"], state, spec["stats"], spec["y_targets"].name, spec["T"], 1, n, None,
                             spec["p"], mask_source=lambda f: list(ms[f]))
    for f in range(3):
        want = np.asarray(orc.step(rows[0, f]))
        msg_close(got[f].msg[0, 0], want[:25])
        assert np.abs(got[f].samples[0, 0].ravel() - want[25:]).max() <= POS_TOL
