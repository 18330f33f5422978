"""BASELINE configs[3]: batched offline relabelling of synthetic recordings (60 s @ 60 Hz, watch-only model, 100 MC samples)
through record/replay.py::relabel_recordings - host rows in, host messages out (pinned H2D / D2H every call).  Times a bounded
sample of R recordings per GPU and extrapolates to the 10 000-recording job (recordings are independent: the job is sharded
with shard_streams, no collective).  Run under torchrun for N > 1: every rank relabels its own shard of R recordings (weak)."""
import os, sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator
from arm_pose_estimation_b200.record.replay import relabel_recordings

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
F = int(sys.argv[2]) if len(sys.argv) > 2 else 3600
fpc = int(sys.argv[3]) if len(sys.argv) > 3 else 4
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
kind = syn.KIND_WATCH_ONLY
spec = syn.kind_spec(kind)
state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
base = syn.synth_rows(kind, 16, F, config_id=4, first_stream=rank * R)
rows = np.ascontiguousarray(np.tile(base, (-(-R // 16), 1, 1))[:R])


def make(n_streams, frames_per_call):
    return BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"], stats=spec["stats"],
                            n_streams=n_streams, mc_samples=100, smooth=1, dropout=spec["p"], frames_per_call=frames_per_call,
                            mask_mode=N.MASK_PHILOX, philox_seed=4, first_stream=rank * R, emit_samples=False)


relabel_recordings(rows[:, :64], make, frames_per_call=fpc)          # warm-up (allocations, probe)
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
t0 = time.perf_counter()
res = relabel_recordings(rows, make, frames_per_call=fpc)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    dt = float(t.item())
if rank == 0:
    est = R * F * world
    print(json.dumps({"workload": f"relabel {R} recordings per GPU x {F} frames, watch-only model, 100 MC samples, {fpc} frames per call, host rows in / host messages out",
                      "n_gpus": world, "estimates": est, "seconds": dt, "estimates_per_s": est / dt,
                      "job_10000_recordings_x_3600_frames_s": 10000 * 3600 / (est / dt), "finite": bool(np.isfinite(res[0]["msg"]).all())}))
if world > 1:
    torch.distributed.destroy_process_group()
