"""Cluster kernel (one 8-CTA cluster per 128 rows) against the layer kernels for calls of a few hundred to a few thousand rows - the
multi-stream real-time regime: python tools/tcl_crossover.py   (per-call device time of the whole three-stage call, CUDA events)"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator

for kind in (syn.KIND_POCKET, syn.KIND_UARM):
    spec = syn.kind_spec(kind)
    state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
    for B in (1, 2, 4, 8, 16, 32, 64):
        n = 100
        res = {}
        for name, rows_cap, pipe in (("cluster", 1 << 20, True), ("layers", 0, False), ("layers+pipeline", 0, True)):
            BatchedEstimator.SMALL_BATCH_ROWS = rows_cap
            if B * n > 128 * 64 and name == "cluster":
                continue
            be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                                  n_streams=B, mc_samples=n, smooth=1, dropout=spec["p"], frames_per_call=1, mask_mode=N.MASK_PHILOX, philox_seed=7,
                                  lstm_variant="tc", pipeline=pipe)
            rows = syn.synth_rows(kind, B, 64, config_id=2)
            dev = [torch.from_numpy(np.ascontiguousarray(rows[:, f:f + 1])).cuda() for f in range(64)]
            for f in range(8):
                be.step_device(dev[f], raw_ready=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # latency: one call at a time
            lat = []
            for f in range(8, 40):
                e0.record(); be.step_device(dev[f], raw_ready=True); e1.record(); e1.synchronize()
                lat.append(e0.elapsed_time(e1) * 1e3)
            # throughput: back to back
            e0.record()
            for f in range(40, 64):
                be.step_device(dev[f], raw_ready=True)
            e1.record(); e1.synchronize()
            res[name] = (float(np.median(lat)), e0.elapsed_time(e1) * 1e3 / 24, be.small_batch)
        print(syn.KIND_NAMES[kind], "streams", B, "rows", B * n, {k: f"latency {v[0]:.1f} us, back-to-back {v[1]:.1f} us/call (small_batch={v[2]})" for k, v in res.items()}, flush=True)
