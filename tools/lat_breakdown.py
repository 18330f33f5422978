import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator
kind = syn.KIND_POCKET
spec = syn.kind_spec(kind)
state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                      n_streams=1, mc_samples=100, smooth=1, dropout=spec["p"], frames_per_call=1, mask_mode=N.MASK_PHILOX, philox_seed=7,
                      lstm_variant="tc", small_batch_kernel=True)
rows = syn.synth_rows(kind, 1, 40, config_id=2)
for f in range(40):
    be.step_graph(rows[:, f:f + 1])
lm = np.zeros(2, np.float32)
dev = torch.from_numpy(rows[:, :1].copy()).cuda()
for _ in range(3):
    be.step_device(dev, layer_ms=lm)
print("tcl launch ms", lm)
