"""Soak test of the tensor-core kernels: two estimators with the same seed run N steps at the full bench shape; the device
results must be bit-identical at every step (a race between the loader / issuer / producer / epilogue roles, the two CTAs of a
pair or the two streams of the cross-call pipeline would show up as a difference), finite, and unit-norm."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for kind in (syn.KIND_WATCH_ONLY, syn.KIND_POCKET, syn.KIND_UARM):
    spec = syn.kind_spec(kind)
    state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
    mk = lambda: BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                                  n_streams=1024, mc_samples=100, smooth=1, dropout=spec["p"], frames_per_call=1, mask_mode=N.MASK_PHILOX,
                                  philox_seed=99, emit_samples=True)
    a, b = mk(), mk()
    rows = torch.from_numpy(np.tile(syn.synth_rows(kind, 64, 64, config_id=7), (16, 1, 1))).cuda()
    frames = [rows[:, f:f + 1].contiguous() for f in range(64)]
    bad = 0
    for s in range(steps):
        oa = a.step_device(frames[s % 64], raw_ready=True)
        ob = b.step_device(frames[s % 64], raw_ready=True)
        if s % 50 == 49 or s == steps - 1:
            torch.cuda.synchronize()
            same = torch.equal(oa.msg, ob.msg) and torch.equal(oa.samples, ob.samples) and torch.equal(oa.std, ob.std)
            fin = bool(torch.isfinite(oa.msg).all()) and bool(torch.isfinite(oa.samples).all())
            qn = float((oa.msg[..., 0:4].norm(dim=-1) - 1).abs().max())
            if not (same and fin and qn < 1e-4):
                bad += 1
                print(f"{syn.KIND_NAMES[kind]} step {s}: identical={same} finite={fin} |q|-1={qn:.2g}", flush=True)
    print(f"{syn.KIND_NAMES[kind]} ({a.lstm_variant}): {steps} steps x 2 estimators, {bad} bad checks", flush=True)
