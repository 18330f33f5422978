"""Where the end-to-end leg loses against the device-resident one (uarm 1024 x 100): ms per call of the native pipeline with
host rows in / results out, rows in only, results out only, neither.  python tools/e2e_probe.py [calls]"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from arm_pose_estimation_b200 import _native as N, synthetic as syn  # noqa: E402
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator  # noqa: E402

import time
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 200
kind = syn.KIND_UARM
spec = syn.kind_spec(kind)
state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
rows = syn.synth_rows(kind, 1024, 8, config_id=3)
rows_dev = torch.from_numpy(rows).cuda()
frames = [rows_dev[:, k:k + 1].contiguous() for k in range(8)]
hframes = [np.ascontiguousarray(rows[:, k:k + 1]) for k in range(8)]
for emit in (True, False, True):
    be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                          n_streams=1024, mc_samples=100, smooth=1, dropout=spec["p"], frames_per_call=1, mask_mode=N.MASK_PHILOX,
                          philox_seed=2026, emit_samples=emit)
    for name, host_in, d2h in (("host rows in, results out", True, True), ("host rows in only", True, False),
                               ("results out only", False, True), ("device resident", False, False)):
        def call(k):
            flags = (N.PIPE_D2H if d2h else 0)
            return be._native_call(ctypes.c_void_p(hframes[k % 8].ctypes.data) if host_in else None,
                                   None if host_in else ctypes.c_void_p(frames[k % 8].data_ptr()), 1, None, flags)[0]
        time.sleep(0.5)                                         # burst conditions for every variant (the board leaves its power cap)
        for k in range(12):
            call(k)
        N.check(be.lib.ape_pipeline_sync(be._pipe), "sync")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        slots = []
        for k in range(calls):
            slots.append(call(k))
            if d2h and len(slots) >= 5:
                N.check(be.lib.ape_pipeline_wait(be._pipe, slots.pop(0)), "wait")
        N.check(be.lib.ape_pipeline_fence(be._pipe, None), "fence")      # (nothing: keeps the API exercised)
        N.check(be.lib.ape_pipeline_sync(be._pipe), "sync")
        e1.record()
        torch.cuda.synchronize()
        print(f"emit_samples={emit}: {name}: {e0.elapsed_time(e1) / calls:.4f} ms/call")
