"""Host-side cost of one BatchedEstimator.submit() (uarm 1024 x 100): wall time per call with the device kept busy, and a cProfile
of the submit path.  python tools/host_overhead.py [calls]"""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from arm_pose_estimation_b200 import _native as N, synthetic as syn  # noqa: E402
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 300
kind = syn.KIND_UARM
spec = syn.kind_spec(kind)
state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                      n_streams=1024, mc_samples=100, smooth=1, dropout=spec["p"], frames_per_call=1, mask_mode=N.MASK_PHILOX,
                      philox_seed=2026, emit_samples=True)
rows = syn.synth_rows(syn.KIND_UARM, 1024, 8, config_id=3)
pend = []
for k in range(20):
    pend.append(be.submit(rows[:, k % 8:k % 8 + 1]))
for p in pend:
    p.result()
torch.cuda.synchronize()
t0 = time.perf_counter()
host = 0.0
pend = []
for k in range(calls):
    a = time.perf_counter()
    pend.append(be.submit(rows[:, k % 8:k % 8 + 1]))
    host += time.perf_counter() - a
    if len(pend) >= be.N_SLOTS - 1:
        pend.pop(0).result()
for p in pend:
    p.result()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(f"{calls} calls: wall {1e3 * wall / calls:.3f} ms/call, host time inside submit() {1e3 * host / calls:.3f} ms/call")
pr = cProfile.Profile()
pr.enable()
pend = []
for k in range(calls):
    pend.append(be.submit(rows[:, k % 8:k % 8 + 1]))
    if len(pend) >= be.N_SLOTS - 1:
        pend.pop(0).result()
for p in pend:
    p.result()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

# the device-resident entry point: host time per step_device() call (no host wait inside)
rows_dev = torch.from_numpy(rows).cuda()
frames = [rows_dev[:, k:k + 1].contiguous() for k in range(8)]
be.reset()
for k in range(10):
    be.step_device(frames[k % 8], raw_ready=True)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
ev0.record()
for k in range(calls):
    be.step_device(frames[k % 8], raw_ready=True)
ev1.record()
host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"step_device: host enqueue {1e3 * host / calls:.3f} ms/call, device {ev0.elapsed_time(ev1) / calls:.3f} ms/call")
