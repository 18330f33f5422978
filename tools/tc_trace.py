"""Timeline of one CTA of the tensor-core LSTM kernel (SM-clock stamps written by the kernel when a trace buffer is given).
Needs a library built with -DAPE_TC_TRACE=1 (the stamps are compiled out by default): APE_B200_LIB=<that .so> python tools/tc_trace.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 1
kind, B, n = syn.KIND_UARM, 1024, 100
spec = syn.kind_spec(kind)
state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1236)
be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                      n_streams=B, mc_samples=n, dropout=spec["p"], mask_mode=N.MASK_PHILOX, lstm_variant="tc")
rows = torch.from_numpy(np.tile(syn.synth_rows(kind, 64, 4, config_id=3), (16, 1, 1))).cuda()
for f in range(3):
    be.step_device(rows[:, f:f + 1].contiguous())
trace = torch.zeros(768, dtype=torch.int64, device="cuda")
be.step_device(rows[:, 3:4].contiguous(), trace=trace, trace_layer=layer)
torch.cuda.synchronize()
tr = trace.cpu().numpy().reshape(3, 16, 16)
t0 = tr[tr > 0].min()
names = {0: ["step top"] + [f"{e} c{c}" for c in range(4) for e in ("ACC_READY seen", "drained", "h published")] + ["output layer start", "output layer done"],
         1: ["loads issued", "masks applied", "X_DONE seen", "X_READY arrived"],
         2: ["top", "X_READY seen"] + [f"{e} c{c}" for c in range(4) for e in ("SLOT_FREE seen", "x+old pieces issued", "H_READY seen")] + ["step issued"]}
for role, rn in enumerate(["epilogue warp 0 (chunk 0)", "loader warp", "MMA issuer"]):
    print(f"== {rn}")
    for t in range(spec["T"]):
        ev = [(int(tr[role, t, e] - t0), names[role][e] if e < len(names[role]) else str(e)) for e in range(16) if tr[role, t, e] > 0]
        print(f"  t={t}: " + "; ".join(f"{c} {nm}" for c, nm in ev))
