"""Bring-up of the split-precision kernel: tcx vs the fp32 kernel on a few shapes (same Philox masks), timing of both.
python tools/tcx_probe.py"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator

kind = syn.KIND_UARM
spec = syn.kind_spec(kind)


def make(B, n, variant, scale=1.0, **kw):
    state = {k: v.copy() for k, v in syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind).items()}
    for k in state:
        if k.startswith("lstm.weight_"):
            state[k] = (state[k] * np.float32(scale)).astype(np.float32)
    return BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                            n_streams=B, mc_samples=n, smooth=1, dropout=spec["p"], frames_per_call=1, mask_mode=N.MASK_PHILOX,
                            philox_seed=11, lstm_variant=variant, **kw)


for B, n, scale in ((2, 16, 1.0), (3, 100, 1.0), (200, 100, 1.0), (64, 100, 4.0), (64, 100, 8.0), (1024, 100, 1.0)):
    rows = np.tile(syn.synth_rows(kind, min(B, 8), 2, config_id=6), ((B + 7) // 8, 1, 1))[:B]
    x, r = make(B, n, "tcx", scale), make(B, n, "fp32", scale)
    for f in range(2):
        a, b = x.step(rows[:, f:f + 1]), r.step(rows[:, f:f + 1])
        print(f"B={B} n={n} scale={scale} frame {f}: tcx vs fp32 max |d| samples {np.abs(a.samples - b.samples).max():.3g} m, msg {np.abs(a.msg - b.msg).max():.3g}, "
              f"finite {np.isfinite(a.msg).all()}, probe {x.tcx_probe_error_m:.3g}", flush=True)
be = make(1024, 100, "tcx")
rows = np.tile(syn.synth_rows(kind, 8, 8, config_id=6), (128, 1, 1))
dev = [torch.from_numpy(np.ascontiguousarray(rows[:, f:f + 1])).cuda() for f in range(8)]
for f in range(5):
    be.step_device(dev[f], raw_ready=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for f in range(50):
    be.step_device(dev[f % 8], raw_ready=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
lm = np.zeros(3, np.float32)
time.sleep(0.3)
be.step_device(dev[0], layer_ms=lm)
print(f"tcx 1024 x 100: {ms:.3f} ms/step = {1024 / ms * 1e3:.3g} est/s; layer_ms {lm}")
