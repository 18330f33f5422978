"""Bring-up of the small-batch cluster kernel (csrc/ape_lstm_tcl.cu): against the layer kernels on the same Philox masks, and the
single-stream frame latency with and without it.  python tools/tcl_probe.py"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator


def make(kind, B, n, small, **kw):
    spec = syn.kind_spec(kind)
    state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
    return BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                            n_streams=B, mc_samples=n, smooth=kw.pop("smooth", 1), dropout=spec["p"], frames_per_call=kw.pop("nF", 1),
                            mask_mode=N.MASK_PHILOX, philox_seed=11, lstm_variant="tc", small_batch_kernel=small, **kw)


for kind in (syn.KIND_POCKET, syn.KIND_WATCH_ONLY, syn.KIND_UARM):
    for B, n, nF in ((1, 100, 1), (1, 1, 1), (3, 40, 1), (2, 16, 4)):
        rows = syn.synth_rows(kind, B, 3 * nF, config_id=6)
        x, r = make(kind, B, n, True, nF=nF, smooth=2), make(kind, B, n, False, nF=nF, smooth=2)
        assert x.small_batch and not r.small_batch
        for c in range(3):
            a, b = x.step(rows[:, c * nF:(c + 1) * nF]), r.step(rows[:, c * nF:(c + 1) * nF])
            print(f"{syn.KIND_NAMES[kind]} B={B} n={n} nF={nF} call {c}: small-batch vs layer kernels max |d| {np.abs(a.samples - b.samples).max():.3g} m, "
                  f"msg {np.abs(a.msg - b.msg).max():.3g}, finite {np.isfinite(a.msg).all()}", flush=True)
for kind in (syn.KIND_POCKET, syn.KIND_UARM):
    for small in (False, True):
        be = make(kind, 1, 100, small)
        rows = syn.synth_rows(kind, 1, 320, config_id=2)
        lat = []
        for f in range(320):
            t0 = time.perf_counter()
            be.step_graph(rows[:, f:f + 1])
            lat.append(time.perf_counter() - t0)
        lat = np.asarray(lat[20:]) * 1e3
        print(f"{syn.KIND_NAMES[kind]} 1 x 100, small_batch_kernel={small}: frame latency p50 {np.percentile(lat, 50):.4f} ms, p99 {np.percentile(lat, 99):.4f} ms")
