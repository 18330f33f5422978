"""GPU (-m gpu): the split-precision tensor-core kernel (csrc/ape_lstm_tcx.cu, H = 128: every operand an fp16 pair hi + lo, three
tcgen05 passes per product, ex2 / rcp cell update) - the tensor-core path for models whose weights the single-pass fp16 kernels
cannot carry.  Checked against the oracle with injected masks at weight scales 1x .. 8x and biased forget gates (bound 1e-4 m,
measured ~1e-6 .. 2e-5 m), against the reference's own messages, against the fp32 kernel on ragged / multi-tile shapes under
Philox, through the pipeline, and as what "auto" selects when the single-pass probe fails."""
import numpy as np
import pytest

from conftest import load_golden, unpack_masks
from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200 import synthetic as syn
from oracle import estimator as OE
from test_gpu_parity import msg_close, POS_TOL
from test_gpu_tc import make, _scaled_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["uarm_s1", "uarm_s4"])
def test_split_whole_path_against_reference_messages(name):
    g = load_golden(f"e2e_{name}.npz")
    n, smooth = int(g["n"]), int(g["smooth"])
    masks = unpack_masks(g)
    rows, F = g["rows"], len(g["rows"])
    be, spec, _ = make(syn.KIND_UARM, 1, n, "tcx", smooth=smooth, frames_per_call=F, mask_mode=N.MASK_INJECTED)
    assert be.lstm_variant == "tc" and be.tc_split
    out = be.step(rows[None], masks=masks[None])
    worst = msg_close(out.msg[0], g["msgs"][:, :25])
    err = np.abs(out.samples[0].reshape(F, -1) - g["msgs"][:, 25:]).max()
    print(f"{name} split-precision path: worst position error vs the reference's messages {max(worst, err):.3g} m (probe {be.tcx_probe_error_m:.3g} m)")
    assert err <= 1e-5                                                      # fp32-grade: ten times inside north_star's bound


@pytest.mark.parametrize("scale,forget_bias", [(1.0, 0.0), (1.0, 1.0), (2.0, 1.0), (4.0, 0.0), (8.0, 1.0)])
def test_split_weight_scale_sweep_against_oracle(scale, forget_bias):
    # the sweep that sends the single-pass kernels to 2.5e-4 / 6.6e-4 / 1.7e-2 m at 2x / 4x / 8x (test_gpu_tc.py): the split-precision
    # kernel must hold north_star's 1e-4 m at every scale, and "auto" must pick a tensor-core path (single pass or split) - not the
    # 22x slower fp32 kernel - for this H = 128 model at every scale
    kind, B, nF, n = syn.KIND_UARM, 2, 2, 48
    state = _scaled_state(kind, scale, forget_bias)
    spec = syn.kind_spec(kind)
    rng = np.random.default_rng(int(scale * 10) + kind)
    rows = syn.synth_rows(kind, B, nF, config_id=44)
    masks = (rng.random(size=(B, nF, spec["L"] - 1, spec["T"], n, spec["H"])) < 0.8).astype(np.uint8)
    want = []
    for b in range(B):
        orc = OE.OracleEstimator(syn.KIND_NAMES[kind], spec["lookup"], state, spec["stats"], spec["y_targets"].name, spec["T"], 1, n,
                                 None, spec["p"], mask_source=lambda f, b=b: list(masks[b, f]))
        want.append([np.asarray(orc.step(rows[b, f])) for f in range(nF)])
    errs = {}
    for variant in ("tcx", "auto"):
        be, _, _ = make(kind, B, n, variant, state=state, frames_per_call=nF, mask_mode=N.MASK_INJECTED)
        out = be.step(rows, masks=masks)
        worst = 0.0
        for b in range(B):
            for f in range(nF):
                w = want[b][f]
                for a, c in ((4, 7), (11, 14), (18, 21)):
                    worst = max(worst, float(np.abs(out.msg[b, f, a:c] - w[a:c]).max()))
                worst = max(worst, float(np.abs(out.samples[b, f].ravel() - w[25:]).max()))
        errs[variant] = (worst, be.lstm_variant, be.tc_split, be.tc_probe_error_m, be.tcx_probe_error_m)
    print(f"uarm weights x{scale} forget bias +{forget_bias}: split precision {errs['tcx'][0]:.3g} m (probe {errs['tcx'][4]:.3g} m); "
          f"auto -> {'split' if errs['auto'][2] else 'single pass'} {errs['auto'][0]:.3g} m (single-pass probe {errs['auto'][3]:.3g} m)")
    assert errs["tcx"][0] <= POS_TOL
    assert errs["auto"][1] == "tc" and errs["auto"][0] <= POS_TOL            # a tensor-core path at every scale, within the bound
    if scale >= 2.0:
        assert errs["auto"][2]                                              # ... the split one once the single pass fails its probe


@pytest.mark.parametrize("B,n", [(1, 1), (3, 70), (5, 100), (200, 100)])
def test_split_matches_fp32_kernel_with_philox(B, n):
    # ragged row counts, one row, and 79 tiles > 74 CTA pairs (items run on across tile boundaries); the Philox keys do not depend
    # on the kernel, so both variants see the same masks
    kind, nF = syn.KIND_UARM, 2
    rows = np.tile(syn.synth_rows(kind, min(B, 8), 3 * nF, config_id=7), ((B + 7) // 8, 1, 1))[:B]
    kw = dict(frames_per_call=nF, mask_mode=N.MASK_PHILOX, philox_seed=3, smooth=2)
    x, spec, _ = make(kind, B, n, "tcx", **kw)
    r, _, _ = make(kind, B, n, "fp32", **kw)
    for c in range(3):
        a, b = x.step(rows[:, c * nF:(c + 1) * nF]), r.step(rows[:, c * nF:(c + 1) * nF])
        d = max(float(np.abs(a.samples - b.samples).max()), float(np.abs(a.msg - b.msg).max()), float(np.abs(a.std - b.std).max()))
        print(f"split precision vs fp32 kernel, {B} x {nF} x {n}, call {c}: max |difference| {d:.3g} m")
        assert np.isfinite(a.msg).all() and d <= 5e-6


def test_split_pipeline_is_bitwise_invisible_and_repeats():
    # submitted back to back through the native pipeline (two lanes, side stream) vs one call at a time without any pipelining
    kind, B, n, calls = syn.KIND_UARM, 300, 100, 6
    rows = np.tile(syn.synth_rows(kind, 8, calls, config_id=9), (38, 1, 1))[:B]
    kw = dict(frames_per_call=1, mask_mode=N.MASK_PHILOX, philox_seed=8, smooth=3)
    piped, _, _ = make(kind, B, n, "tcx", **kw)
    plain, _, _ = make(kind, B, n, "tcx", pipeline=False, **kw)
    assert piped._pipe is not None and plain._pipe is None
    pend = [piped.submit(rows[:, k:k + 1]) for k in range(calls)]
    for k, p in enumerate(pend):
        got, want = p.result(), plain.step(rows[:, k:k + 1])
        np.testing.assert_array_equal(got.msg, want.msg, err_msg=f"call {k}")
        np.testing.assert_array_equal(got.samples, want.samples, err_msg=f"call {k}")
        np.testing.assert_array_equal(got.std, want.std, err_msg=f"call {k}")
