"""Where a single-stream frame's wall time goes (configs[1], step_graph): host staging / graph launch / wait / result views.
python tools/lat_host_breakdown.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from arm_pose_estimation_b200 import _native as N, synthetic as syn             # noqa: E402
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator          # noqa: E402

kind = syn.KIND_POCKET
spec = syn.kind_spec(kind)
state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                      n_streams=1, mc_samples=100, smooth=1, dropout=spec["p"], frames_per_call=1, mask_mode=N.MASK_PHILOX, philox_seed=7,
                      lstm_variant="tc")
rows = syn.synth_rows(kind, 1, 400, config_id=2)
for f in range(50):
    be.step_graph(rows[:, f:f + 1])
t = []
for f in range(50, 350):
    t0 = time.perf_counter()
    be.step_graph(rows[:, f:f + 1])
    t.append(time.perf_counter() - t0)
print("step_graph p50 %.1f us  p99 %.1f us" % (np.percentile(t, 50) * 1e6, np.percentile(t, 99) * 1e6))
# the bare graph: launch + wait, nothing else
st = torch.cuda.current_stream()
g = be._g.graph
tl, tw = [], []
for _ in range(300):
    t0 = time.perf_counter()
    g.replay()
    t1 = time.perf_counter()
    st.synchronize()
    t2 = time.perf_counter()
    tl.append(t1 - t0)
    tw.append(t2 - t0)
print("graph.replay() returns after p50 %.1f us; replay + synchronize p50 %.1f us  p99 %.1f us" % (
    np.percentile(tl, 50) * 1e6, np.percentile(tw, 50) * 1e6, np.percentile(tw, 99) * 1e6))
# device time of the graph: events around the replay
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d = []
for _ in range(100):
    e0.record()
    g.replay()
    e1.record()
    e1.synchronize()
    d.append(e0.elapsed_time(e1) * 1e3)
print("graph device time (events) p50 %.1f us" % np.percentile(d, 50))
