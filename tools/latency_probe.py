"""Where does the single-stream frame latency (BASELINE configs[1]) go?  Host wall clock around the pieces of BatchedEstimator.step."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator

kind = syn.KIND_POCKET
spec = syn.kind_spec(kind)
state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
rows = syn.synth_rows(kind, 1, 400, config_id=2)
for variant in ("fp32", "tc"):
    be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                          n_streams=1, mc_samples=100, smooth=1, dropout=spec["p"], frames_per_call=1, mask_mode=N.MASK_PHILOX, philox_seed=7,
                          lstm_variant=variant)
    lat = []
    for f in range(320):
        t0 = time.perf_counter(); be.step(rows[:, f:f + 1]); lat.append(time.perf_counter() - t0)
    lat = np.asarray(lat[20:]) * 1e3
    dev = torch.from_numpy(rows).cuda()
    fr = [dev[:, f:f + 1].contiguous() for f in range(100)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f in range(100):
        be.step_device(fr[f], raw_ready=True)
    t_enq = (time.perf_counter() - t0) / 100 * 1e3
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 100 * 1e3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for f in range(100):
        be.step_device(fr[f], raw_ready=True)
    e1.record(); torch.cuda.synchronize()
    layer_ms = np.zeros(spec["L"], np.float32)
    be.step_device(fr[0], layer_ms=layer_ms)
    print(f"{variant}: step() p50 {np.percentile(lat, 50):.3f} ms p99 {np.percentile(lat, 99):.3f} ms | step_device enqueue {t_enq:.3f} ms/frame, "
          f"enqueue+drain {t_all:.3f} ms/frame, device {e0.elapsed_time(e1) / 100:.3f} ms/frame | LSTM layers {layer_ms.tolist()} ms", flush=True)
