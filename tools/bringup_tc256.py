"""Bring-up probe of the H = 256 streamed-weights tensor-core kernel: tensor-core vs fp32 kernel on growing batches."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate import batched

def make(kind, B, n, variant, **kw):
    spec = syn.kind_spec(kind)
    state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
    return batched.BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"],
                                    stats=spec["stats"], n_streams=B, mc_samples=n, dropout=spec["p"], lstm_variant=variant, **kw)

for kind in (syn.KIND_POCKET, syn.KIND_WATCH_ONLY):
    for B, n in ((1, 64), (3, 100), (40, 100), (1024, 100)):
        rows = np.tile(syn.synth_rows(kind, 8, 2, config_id=6), (128, 1, 1))[:B]
        a = make(kind, B, n, "tc", mask_mode=N.MASK_PHILOX, philox_seed=78)
        b = make(kind, B, n, "fp32", mask_mode=N.MASK_PHILOX, philox_seed=78)
        for f in range(2):
            oa, ob = a.step(rows[:, f:f + 1]), b.step(rows[:, f:f + 1])
            print(syn.KIND_NAMES[kind], B, n, "frame", f, "probe", a.tc_probe_error_m, "max |d samples|", float(np.abs(oa.samples - ob.samples).max()),
                  "max |d msg|", float(np.abs(oa.msg - ob.msg).max()), flush=True)
        if B == 1024:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                a.step_device(a.raw, 1)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 20
            print(syn.KIND_NAMES[kind], "tc step", dt * 1e3, "ms ->", B / dt, "est/s", flush=True)
