"""A/B of the two-lane cross-call pipeline (BatchedEstimator(lanes=...)): device ms per step and host enqueue ms per step.

    python tools/lanes_ab.py [workload] [steps]
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from arm_pose_estimation_b200 import _native as N, synthetic as syn            # noqa: E402
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator          # noqa: E402
from bench import WORKLOADS                                                     # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "uarm_1024x100"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    kind, B, n, smooth = WORKLOADS[workload]
    spec = syn.kind_spec(kind)
    I, H, L, T, O = (spec[k] for k in "IHLTO")
    state = syn.synth_state_dict(I, H, L, O, 1234 + kind)
    rows = syn.synth_rows(kind, min(64, B), 8, config_id=3, first_stream=0)
    rows = np.ascontiguousarray(np.tile(rows, (-(-B // rows.shape[0]), 1, 1))[:B])
    rows_dev = torch.from_numpy(rows).cuda()
    frames = [rows_dev[:, f:f + 1].contiguous() for f in range(8)]
    for lanes in (False, True, False, True):
        be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=T, y_targets=spec["y_targets"],
                              stats=spec["stats"], n_streams=B, mc_samples=n, smooth=smooth, dropout=spec["p"],
                              mask_mode=N.MASK_PHILOX, philox_seed=2026, lstm_variant="tc", lanes=lanes)
        for f in range(10):
            be.step_device(frames[f % 8], raw_ready=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for f in range(K):
            be.step_device(frames[f % 8], raw_ready=True)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"{workload} lanes={lanes}: {ms:.4f} ms/step device ({B / ms * 1e3:.4g} est/s), host enqueue {(t1 - t0) / K * 1e3:.4f} ms/step", flush=True)
        del be


if __name__ == "__main__":
    main()
