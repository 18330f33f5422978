"""Single-stream frame latency (BASELINE configs[1]) with and without the cross-call pipeline machinery."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator

kind = syn.KIND_POCKET
spec = syn.kind_spec(kind)
state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
rows = syn.synth_rows(kind, 1, 400, config_id=2)
for pipeline in (True, False):
    be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                          n_streams=1, mc_samples=100, smooth=1, dropout=spec["p"], frames_per_call=1, mask_mode=N.MASK_PHILOX, philox_seed=7,
                          pipeline=pipeline)
    lat = []
    for f in range(320):
        t0 = time.perf_counter(); be.step(rows[:, f:f + 1]); lat.append(time.perf_counter() - t0)
    lat = np.asarray(lat[20:]) * 1e3
    print(f"pipeline={pipeline} variant={be.lstm_variant}: step() p50 {np.percentile(lat, 50):.3f} ms p99 {np.percentile(lat, 99):.3f} ms", flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for f in range(200): be.step(rows[:, f:f + 1])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
