"""GPU (-m gpu): the two-layer wavefront tensor-core kernel (csrc/ape_lstm_tcw.cu: layers 1 and 2 of the H = 128 model in one
launch, streamed weights, h in tensor memory, dropout of the inner gap applied by the epilogue warps) against the reference's
golden messages, the oracle with injected masks, and the one-layer-per-launch tensor-core path it replaces (same operand
rounding, same Philox keys, same MMA order: the two must agree to the last bit).  Tolerance vs oracle / reference: 1e-4 m."""
import numpy as np
import pytest

from conftest import load_golden, unpack_masks
from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200 import synthetic as syn
from oracle import estimator as OE
from test_gpu_parity import msg_close, POS_TOL
from test_gpu_tc import make

pytestmark = pytest.mark.gpu
PAIRS, SINGLE = 2, 1             # ape_lstm_args.tc_flags


@pytest.mark.parametrize("name", ["uarm_s1", "uarm_s4"])
def test_wavefront_whole_path_against_reference_messages(name):
    g = load_golden(f"e2e_{name}.npz")
    n, smooth = int(g["n"]), int(g["smooth"])
    masks = unpack_masks(g)
    rows, F = g["rows"], len(g["rows"])
    be, spec, _ = make(syn.KIND_UARM, 1, n, "tc", smooth=smooth, frames_per_call=F, mask_mode=N.MASK_INJECTED, tc_flags=PAIRS)
    out = be.step(rows[None], masks=masks[None])
    worst = msg_close(out.msg[0], g["msgs"][:, :25])
    err = np.abs(out.samples[0].reshape(F, -1) - g["msgs"][:, 25:]).max()
    print(f"{name} wavefront kernel: worst position error vs the reference's messages {max(worst, err):.3g} m")
    assert err <= POS_TOL


@pytest.mark.parametrize("B,nF,n", [(1, 1, 7), (3, 2, 70), (5, 3, 100)])
def test_wavefront_ragged_rows_against_oracle(B, nF, n):
    # injected masks on both gaps; rows not a multiple of the 256-row tile; several tiles per call
    kind = syn.KIND_UARM
    be, spec, state = make(kind, B, n, "tc", frames_per_call=nF, mask_mode=N.MASK_INJECTED, tc_flags=PAIRS)
    rng = np.random.default_rng(B * 100 + n)
    rows = syn.synth_rows(kind, B, nF, config_id=43)
    masks = (rng.random(size=(B, nF, spec["L"] - 1, spec["T"], n, spec["H"])) < 0.8).astype(np.uint8)
    out = be.step(rows, masks=masks)
    worst = 0.0
    for b in range(B):
        orc = OE.OracleEstimator("uarm", spec["lookup"], state, spec["stats"], spec["y_targets"].name, spec["T"], 1, n, None,
                                 spec["p"], mask_source=lambda f, b=b: list(masks[b, f]))
        for f in range(nF):
            want = np.asarray(orc.step(rows[b, f]))
            worst = max(worst, msg_close(out.msg[b, f], want[:25]), float(np.abs(out.samples[b, f].ravel() - want[25:]).max()))
    print(f"wavefront kernel vs oracle, {B} x {nF} x {n}: worst position error {worst:.3g} m")
    assert worst <= POS_TOL


@pytest.mark.parametrize("B,n,mode", [(2, 100, N.MASK_PHILOX), (200, 100, N.MASK_PHILOX), (37, 64, N.MASK_INJECTED), (1024, 100, N.MASK_PHILOX)])
def test_wavefront_equals_one_layer_launches(B, n, mode):
    # 200 x 100 rows = 79 tiles > 74 CTA pairs: some pairs carry two tiles and the wavefront runs across the tile boundary;
    # 1024 x 100 is the benchmark shape (400 tiles, 5.4 rounds)
    kind, nF = syn.KIND_UARM, 2
    rows = np.tile(syn.synth_rows(kind, min(B, 8), nF, config_id=6), ((B + 7) // 8, 1, 1))[:B]
    kw = dict(frames_per_call=1, mask_mode=mode, philox_seed=91, smooth=2)
    pair, spec, _ = make(kind, B, n, "tc", tc_flags=PAIRS, **kw)
    single, _, _ = make(kind, B, n, "tc", tc_flags=SINGLE, **kw)
    rng = np.random.default_rng(3)
    for f in range(nF):
        masks = None
        if mode == N.MASK_INJECTED:
            masks = (rng.random(size=(B, 1, spec["L"] - 1, spec["T"], n, spec["H"])) < 0.8).astype(np.uint8)
        oa, ob = pair.step(rows[:, f:f + 1], masks=masks), single.step(rows[:, f:f + 1], masks=masks)
        assert np.isfinite(oa.msg).all()
        diff = float(np.abs(oa.samples - ob.samples).max())
        print(f"wavefront vs one-layer launches, {B} x {n}, frame {f}: max |difference| {diff:.3g} m")
        # same operands, same Philox keys; the wavefront kernel adds the bias inside the accumulator (an extra K step against a
        # tile of ones) instead of after it, so the pre-activations differ in the last fp32 bits - and tanh.approx, a piecewise
        # approximation good to 2^-11, turns a last-bit difference at a segment boundary into a 1e-4 relative step of that gate
        # (measured: 1e-5 m; the bound is the tensor-core kernels' own distance from the fp32 kernel)
        assert diff <= 5e-5
        assert float(np.abs(oa.msg - ob.msg).max()) <= 5e-5 and float(np.abs(oa.std - ob.std).max()) <= 5e-5


def test_wavefront_is_the_default_when_the_batch_fills_the_gpu_and_repeats_bitwise():
    kind, B, n = syn.KIND_UARM, 1024, 100
    rows = np.tile(syn.synth_rows(kind, 32, 2, config_id=8), (32, 1, 1))
    be, _, _ = make(kind, B, n, "tc", mask_mode=N.MASK_PHILOX, philox_seed=5)          # tc_flags = 0: automatic
    ref, _, _ = make(kind, B, n, "tc", mask_mode=N.MASK_PHILOX, philox_seed=5, tc_flags=SINGLE)
    want = first = None
    for rep in range(10):
        be.reset()
        be.step(rows[:, 0:1])
        out = be.step(rows[:, 1:2])
        if want is None:
            ref.step(rows[:, 0:1])
            want = ref.step(rows[:, 1:2])
            assert be.launches == 2 * (2 + 2) and ref.launches == 2 * (2 + 3)      # features, layer 0, ONE pair launch, stage 3
            first = (out.msg.copy(), out.samples.copy())
        assert float(np.abs(out.samples - want.samples).max()) <= 5e-5             # (see above)
        np.testing.assert_array_equal(out.msg, first[0])                           # the wavefront launch itself repeats bit for bit
        np.testing.assert_array_equal(out.samples, first[1])
