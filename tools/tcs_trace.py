"""Timeline of one step of CTA 0 of the streamed-weights tensor-core LSTM kernel (H = 256): per weight piece the issuer's
wait / issue stamps and the producer's slot-free / copy-issued stamps, plus the epilogue's half-pass starts.
The stamps need a library built with -DAPE_TCS_TRACE=1 (APE_B200_LIB=<that .so>)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator

kind, B, n = syn.KIND_WATCH_ONLY, 1024, 100
spec = syn.kind_spec(kind)
state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234)
be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                      n_streams=B, mc_samples=n, dropout=spec["p"], mask_mode=N.MASK_PHILOX, lstm_variant="tc")
rows = torch.from_numpy(np.tile(syn.synth_rows(kind, 64, 4, config_id=3), (16, 1, 1))).cuda()
for f in range(3):
    be.step_device(rows[:, f:f + 1].contiguous())
trace = torch.zeros(768, dtype=torch.int64, device="cuda")
be.step_device(rows[:, 3:4].contiguous(), trace=trace, trace_layer=1)
torch.cuda.synchronize()
tr = trace.cpu().numpy()
t0 = tr[tr > 0].min()
rel = lambda v: int(v - t0) if v > 0 else -1
print("issuer: X_READY seen", rel(tr[576]))
print("issuer: SLOT_FREE seen per chunk", [rel(v) for v in tr[560:568]])
print("issuer: H_READY seen per slice", [rel(v) for v in tr[568:576]])
print("loader: first batch loaded", rel(tr[580]), "X_DONE seen", rel(tr[581]), "X_READY arrived", rel(tr[582]))
print("epilogue half-pass starts, step T-1:", [rel(v) for v in tr[512:528]])
print("epilogue half-pass starts, step T  :", [rel(v) for v in tr[528:544]])
iw, ii = tr[0:256:2], tr[1:256:2]
pw, pi = tr[256:512:2], tr[257:512:2]
print("piece: issuer wait-done, issue-done (dt) | producer slot-free, copy-issued | copy-issued -> issuer wait-done")
for p in range(128):
    if iw[p] == 0:
        break
    print(f"  {p:3d}: {rel(iw[p]):7d} {rel(ii[p]):7d} ({int(ii[p]-iw[p]):4d}) | {rel(pw[p]):7d} {rel(pi[p]):7d} | {int(iw[p]-pi[p]):6d} | FULL spins {int(tr[600 + p])}")
d = np.diff(iw[:128][iw[:128] > 0])
print("issuer wait-done deltas: mean", d.mean(), "median", np.median(d), "max", d.max())
