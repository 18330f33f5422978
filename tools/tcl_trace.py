"""Per-step timeline of the small-batch cluster kernel (CTA 0): python tools/tcl_trace.py"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator
kind = syn.KIND_POCKET
spec = syn.kind_spec(kind)
state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                      n_streams=1, mc_samples=100, smooth=1, dropout=spec["p"], frames_per_call=1, mask_mode=N.MASK_PHILOX, philox_seed=7)
rows = syn.synth_rows(kind, 1, 8, config_id=2)
dev = [torch.from_numpy(np.ascontiguousarray(rows[:, f:f + 1])).cuda() for f in range(8)]
for f in range(4):
    be.step_device(dev[f])
tr = torch.zeros(768, dtype=torch.int64, device="cuda")
be.step_device(dev[4], trace=tr, trace_layer=-2)
torch.cuda.synchronize()
st = tr.cpu().numpy()[: spec["L"] * spec["T"] * 8].reshape(-1, 8).astype(np.float64)
t0 = st[0, 0]
names = ["x ready", "h landed", "committed", "acc ready", "cell done", "barrier"]
print("step  " + "  ".join(f"{n:>10s}" for n in names) + "   (us since the first stamp)")
for i, r in enumerate(st):
    print(f"{i:4d}  " + "  ".join(f"{(v - t0) / 1e3:10.2f}" if v > 0 else f"{'-':>10s}" for v in r[:6]))
