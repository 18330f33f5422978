"""Benchmark of the per-frame arm-pose estimation hot path (BASELINE.json's metric: MC-sampled arm-pose
estimates/sec at 1/2/4/8 B200 vs the reference CPU path, and % of roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (configs[2] of BASELINE.json, the largest single-GPU configuration and the one the throughput metric
is quoted on): watch+phone upper-arm estimator (I38 H128 L3 T6 O12) with quaternion FK, 1024 concurrent streams
x 100 MC samples PER GPU (weak scaling: streams are sharded, no collective), synthetic IMU rows, seeded
random-init weights.  One step = one frame of every stream = 1024 estimates per GPU.

  value      estimates/s, whole job, inputs resident in HBM, CUDA events, max over ranks
  e2e        the same metric through BatchedEstimator.step() with HOST rows: pinned H2D of the raw rows and D2H
             of messages + std + per-sample positions inside the timed region, every step
  roofline   the dominant kernel (an LSTM layer >= 1 launch): algorithmic flops / its device time (events
             recorded by the library around each layer launch), against MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the oracle's restatement of the reference CPU path (torch.nn.LSTM on the CPU +
             numpy FK, exactly the arithmetic the reference executes) timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (kind, streams per GPU, MC samples, smooth)
    "uarm_1024x100": (2, 1024, 100, 1),
    "pocket_1x100": (1, 1, 100, 1),
    "watch_only_1024x100": (0, 1024, 100, 1),
    "pocket_1024x100": (1, 1024, 100, 1),
}
WORKLOAD_TEXT = {
    "uarm_1024x100": "watch+phone upper-arm estimator with quaternion FK, 1024 concurrent streams x 100 MC samples on 1 B200 (BASELINE configs[2])",
    "pocket_1x100": "watch+phone pocket LSTM estimator, 1 stream x 100 MC samples (BASELINE configs[1])",
    "watch_only_1024x100": "watch-only LSTM estimator (the model of BASELINE configs[0] / [3]), 1024 concurrent streams x 100 MC samples",
    "pocket_1024x100": "watch+phone pocket LSTM estimator (the model of BASELINE configs[1]), 1024 concurrent streams x 100 MC samples",
}
METRIC, UNIT = "mc_sampled_arm_pose_estimates_per_sec", "estimates/s"


def tc_kernel_name(H):
    if H == 256:
        return "lstm_layer_tcs_kernel<256> (one layer >= 1 launch; tcgen05 cta_group::2, fp16 operands, fp32 accumulate, TMA-streamed weights, h_t in TMEM)"
    return f"lstm_layer_tc_kernel<{H}> (one layer >= 1 launch; tcgen05 cta_group::2, fp16 operands, fp32 accumulate, weights resident in shared memory)"


def algorithmic_flops_per_estimate(I, H, L, T, O, n):
    """F(n) of SURVEY.md §8d: layer 0 once per estimate, layers >= 1 per MC sample, output layer on the last step."""
    return T * 2 * 4 * H * (I + H) + n * (L - 1) * T * 2 * 4 * H * (2 * H) + n * 2 * H * O


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = max(smax, float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- the reference CPU path (oracle port), used by cpu_baseline and --impl reference --------------------------------
_W = {}


def _cpu_worker_init(kind, n, smooth, threads):
    """Per-process set-up: one single-stream oracle estimator (torch CPU LSTM + numpy FK) and its synthetic rows."""
    import torch
    torch.set_num_threads(threads)
    from arm_pose_estimation_b200 import synthetic as syn
    from oracle import estimator as OE
    spec = syn.kind_spec(kind)
    state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
    _W["orc"] = OE.OracleEstimator(syn.KIND_NAMES[kind], spec["lookup"], state, spec["stats"], spec["y_targets"].name,
                                   spec["T"], smooth, n, None, spec["p"], mask_source="torch")
    _W["rows"] = syn.synth_rows(kind, 1, 256, config_id=3, first_stream=os.getpid() % 1000)[0]
    _W["pos"] = 0
    for r in _W["rows"][:3]:
        _W["orc"].step(r)                                     # warm-up frames


def _cpu_worker_frames(n_frames):
    orc, rows = _W["orc"], _W["rows"]
    t0 = time.perf_counter()
    for _ in range(n_frames):
        orc.step(rows[_W["pos"] % len(rows)])                 # the three calls of estimator.py:174-176
        _W["pos"] += 1
    return n_frames, time.perf_counter() - t0


class CpuReference:
    """The reference CPU path on `workers` independent single-stream processes (the reference is single-stream by
    construction, nn_models.py:201-202): `rate(frames)` = summed estimates/s over one bounded sample."""

    def __init__(self, kind, n, smooth, workers, threads_per_worker=1):
        import multiprocessing as mp
        self.workers = workers
        self.pool = mp.get_context("spawn").Pool(workers, initializer=_cpu_worker_init,
                                                 initargs=(kind, n, smooth, threads_per_worker))

    def rate(self, frames_per_worker):
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker_frames, [frames_per_worker] * self.workers, chunksize=1)
        wall = time.perf_counter() - t0
        return sum(f for f, _ in res) / max(dt for _, dt in res), wall

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args, rank, world):
    if rank != 0:
        return
    kind, B, n, smooth = WORKLOADS[args.workload]
    workers = max(1, min(os.cpu_count() or 1, 64))
    ref = CpuReference(kind, n, smooth, workers)
    for _ in range(args.warmup):
        ref.rate(2)
    rates, t_all = [], time.perf_counter()
    for _ in range(args.steps):
        rates.append(ref.rate(args.ref_frames)[0])
    wall = time.perf_counter() - t_all
    ref.close()
    value = float(np.mean(rates))
    sample = (f"{workers} single-stream worker processes (1 thread each) x {args.ref_frames} frames per step of the {args.workload} "
              f"workload (n={n} MC samples, smooth={smooth}); torch.nn.LSTM on CPU + numpy FK (oracle port of the reference path)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "streams_per_gpu": B, "mc_samples": n, "smooth": smooth},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- our arm -----------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from arm_pose_estimation_b200 import _native as N, synthetic as syn
    from arm_pose_estimation_b200.estimate.batched import BatchedEstimator, shard_streams

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    kind, B, n, smooth = WORKLOADS[args.workload]
    spec = syn.kind_spec(kind)
    I, H, L, T, O = (spec[k] for k in "IHLTO")
    state = syn.synth_state_dict(I, H, L, O, 1234 + kind)
    first, count = shard_streams(B * world, world, rank)      # weak scaling: B streams per GPU, global stream ids
    be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=T, y_targets=spec["y_targets"],
                          stats=spec["stats"], n_streams=count, mc_samples=n, smooth=smooth, dropout=spec["p"],
                          frames_per_call=1, mask_mode=N.MASK_PHILOX, philox_seed=2026, first_stream=first,
                          emit_samples=True, lstm_variant=args.lstm)
    K, W = args.steps, args.warmup
    # synthetic rows: 64 distinct seeded streams tiled over the shard (generation cost only), K+W frames
    base = syn.synth_rows(kind, min(64, count), K + W, config_id=3, first_stream=first)
    rows = np.ascontiguousarray(np.tile(base, (-(-count // base.shape[0]), 1, 1))[:count])
    rows_dev = torch.from_numpy(rows).cuda()
    frames_dev = [rows_dev[:, f:f + 1].contiguous() for f in range(K + W)]
    frames_host = [np.ascontiguousarray(rows[:, f:f + 1]) for f in range(K + W)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value") ----
    be.reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        t_wait = time.time()
        while not clk.lines and time.time() - t_wait < 10.0:  # nvidia-smi start-up must not land in the timed region
            time.sleep(0.05)
        for f in range(W):
            be.step_device(frames_dev[f], raw_ready=True)
        barrier()
        launches0 = be.launches
        ev0.record()
        for f in range(W, W + K):
            be.step_device(frames_dev[f], raw_ready=True)     # inputs resident in HBM before the timed region
        ev1.record()
        barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = be.launches - launches0
    clocks = clk.summary()

    # ---- end to end through the public host-facing call ----
    be.reset()
    for f in range(W):
        be.step(frames_host[f])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    checksum, pending = 0.0, []
    for f in range(W, W + K):
        pending.append(be.submit(frames_host[f]))             # pinned H2D -> 3 stages -> pinned D2H, enqueued
        if len(pending) >= be.N_SLOTS - 1:                    # (the estimator has N_SLOTS staging slots)
            checksum += float(pending.pop(0).result().msg[0, 0, 4])   # host read of an earlier step's result
    for p in pending:
        checksum += float(p.result().msg[0, 0, 4])
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))

    # ---- roofline leg: per-layer device time of the LSTM stage (events recorded by the library) ----
    time.sleep(0.5)                                           # (launches timed alone against the burst peak: leave the power cap first)
    layer_ms = np.zeros(L, np.float32)
    acc = np.zeros(L, np.float64)
    reps = max(3, min(K, 10))
    for f in range(reps):
        be.step_device(frames_dev[W + f % K], layer_ms=layer_ms)
        acc += layer_ms
    torch.cuda.synchronize()
    acc /= reps
    peaks, peak_src = measured_peaks()
    rows_mc = count * n
    dom_flops = rows_mc * T * 2 * 4 * H * (2 * H)             # one layer >= 1 launch, algorithmic
    dom_ms = float(np.mean(acc[1:-1])) if L > 2 else float(acc[-1])   # a middle layer (no output GEMM); L=2: the last
    fp32_peak_tflops = 148 * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12
    total_est = count * K * world
    tensor = be.lstm_variant == "tc"
    if tensor:
        roofline = {"bound": "tensor", "kernel": tc_kernel_name(H),
                    "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
                    "peak_source": f"MEASURED_PEAKS.json bf16_tflops (burst) [{peak_src}]; fp16 and bf16 tcgen05.mma run at the same rate"}
    else:
        roofline = {"bound": "fp32_ffma", "kernel": "lstm_layer_fma_kernel (one layer >= 1 launch)", "achieved": achieved,
                    "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": achieved / fp32_peak_tflops,
                    "peak_source": f"148 SM x 128 FFMA/clk x 2 x {peaks.get('sm_max_mhz', 1965.0):.0f} MHz (nominal fp32 FMA peak at max SM clock; "
                                   f"MEASURED_PEAKS.json [{peak_src}] has no fp32 figure)",
                    "frac_of_bf16_tensor_peak": achieved / peaks["bf16_tflops"]}
    # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture (profiles/r1i_ncu_summary.md:
    # dram__bytes_read.sum + dram__bytes_write.sum of launch 1); null for configurations that were not captured
    traffic = {("tc", 128, 1024, 100): 1.923840e6 + 98.600960e6, ("tc", 256, 1024, 100): 5.389056e6 + 46848.0}.get(
        (be.lstm_variant, H, count, n)) if args.workload in ("uarm_1024x100", "watch_only_1024x100") else None
    roofline.update({"flops_per_launch": dom_flops, "ms_per_launch": dom_ms, "layer_ms": [float(v) for v in acc], "traffic": traffic,
                     "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1i_ncu_summary.md)"})

    line = {
        "metric": METRIC, "value": total_est / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("f16 operands, f32 accumulate/state" if tensor else "f32"), "data": "synthetic",
        "config": {"workload": args.workload, "baseline_config": WORKLOAD_TEXT[args.workload],
                   "model": {"I": I, "H": H, "L": L, "T": T, "O": O, "dropout": spec["p"]}, "streams_per_gpu": count,
                   "mc_samples": n, "smooth": smooth, "frames_per_step": 1, "estimates_per_step_per_gpu": count,
                   "lstm_variant": ("tcgen05_fp16_operands_fp32_accumulate" if tensor else "fp32_ffma"),
                   "tc_probe_error_m": be.tc_probe_error_m, "parity_tolerance_m": 1e-4, "rng": "philox4x32-10",
                   "l2": f"per-step working set (inter-layer sequences {rows_mc * T * H * 4 / 1e6:.0f} MB) exceeds the 126 MB L2"},
        "e2e": {"value": total_est / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": be.h2d_bytes_per_frame,
                "d2h_bytes_per_step": be.d2h_bytes_per_frame, "ms_per_step": e2e_ms / K},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "checksum": checksum,
    }
    if rank == 0 and world == 1 and not args.no_other_models and args.workload == "uarm_1024x100":
        del be                                                # free its buffers before the other models are set up
        line["other_models"] = {w: quick_throughput(w, args.lstm, BatchedEstimator, N, syn, torch)
                                for w in ("watch_only_1024x100", "pocket_1024x100")}
    if rank == 0 and world == 1 and not args.no_realtime:
        line["realtime"] = realtime_latency(BatchedEstimator, N, syn)
    if rank == 0 and world == 1 and not args.no_other_models:
        line["fk_roofline"] = fk_standalone(N, syn, torch, kind, n, measured_peaks()[0])
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        workers = max(1, min(os.cpu_count() or 1, 64))
        ref = CpuReference(kind, n, smooth, workers)
        ref.rate(2)
        rate, wall = ref.rate(args.cpu_frames)
        ref.close()
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": workers, "kind": "port",
                                "sample": f"{workers} single-stream worker processes (1 thread each) x {args.cpu_frames} frames of the "
                                          f"{args.workload} workload (n={n}); torch.nn.LSTM on CPU + numpy FK; {wall:.1f} s of CPU work"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def fk_standalone(N, syn, torch, kind, n, peaks, E=32768, reps=20):
    """Stage 3 on its own (judged separately against the HBM roofline): `ape_fk_reduce` over E estimates x n MC rows
    of random network targets - 157 MB of predictions in, messages + std + per-sample positions out, far beyond the 126 MB L2.
    Algorithmic bytes per estimate: 4 S O in + 100 (message) + 24 (std) + 24 S (sample positions)."""
    spec = syn.kind_spec(kind)
    O, tgt = spec["O"], {12: N.TARGET_ORI_CAL_LARM_UARM, 14: N.TARGET_ORI_CAL_LARM_UARM_HIPS}[spec["O"]]
    g = torch.Generator(device="cuda").manual_seed(3)
    preds = torch.randn((E, 1, n, O), generator=g, device="cuda", dtype=torch.float32)
    yy_m = torch.as_tensor(np.asarray(spec["stats"]["yy_m"], np.float32)).cuda()
    yy_s = torch.as_tensor(np.asarray(spec["stats"]["yy_s"], np.float32)).cuda()
    from arm_pose_estimation_b200.data_types.bone_map import body_measurements_row
    body = torch.as_tensor(body_measurements_row(None).astype(np.float32).ravel()).cuda()
    msg = torch.empty((E, 25), device="cuda"); std = torch.empty((E, 6), device="cuda")
    samples = torch.empty((E, n, 6), device="cuda"); status = torch.zeros(E, dtype=torch.int32, device="cuda")
    lib, st = N.load(), N.current_stream_ptr()

    def run():
        N.check(lib.ape_fk_reduce(N.ptr(preds), 1, N.ptr(yy_m), N.ptr(yy_s), N.ptr(body), tgt, O, E, 1, 0, None, n, 1,
                                  N.ptr(msg), N.ptr(samples), N.ptr(std), None, N.ptr(status), st), "ape_fk_reduce")
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_per_est = 4 * n * O + 100 + 24 + 24 * n
    gbs = E * bytes_per_est / (ms * 1e-3) / 1e9
    return {"kernel": "fk_reduce_kernel (stage 3 standalone)", "estimates": E, "mc_rows": n, "ms_per_launch": ms,
            "bytes_per_estimate": bytes_per_est, "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": gbs / peaks["hbm_gbs"], "estimates_per_s": E / (ms * 1e-3),
            # DRAM bytes of one launch from the committed ncu capture (profiles/r1i_ncu_summary.md, last section: 157.3 MB read + 54.6 MB
            # written - part of the 83 MB of results is still in the L2 when the kernel ends); only for the captured shape
            "traffic": (157.311744e6 + 54.591232e6) if (E == 32768 and n == 100 and O == 12) else None,
            "l2": f"{E * bytes_per_est / 1e6:.0f} MB per launch: larger than the 126 MB L2"}


def quick_throughput(workload, lstm, BatchedEstimator, N, syn, torch, steps=50, warmup=5):
    """Device-resident throughput + roofline of the dominant kernel for the other two deployed models (both H = 256), same
    shape as the headline workload: 1024 streams x 100 MC samples, one frame of every stream per step."""
    from arm_pose_estimation_b200.estimate.batched import shard_streams
    kind, B, n, smooth = WORKLOADS[workload]
    spec = syn.kind_spec(kind)
    I, H, L, T, O = (spec[k] for k in "IHLTO")
    state = syn.synth_state_dict(I, H, L, O, 1234 + kind)
    be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=T, y_targets=spec["y_targets"], stats=spec["stats"],
                          n_streams=B, mc_samples=n, smooth=smooth, dropout=spec["p"], frames_per_call=1, mask_mode=N.MASK_PHILOX,
                          philox_seed=2026, emit_samples=True, lstm_variant=lstm)
    base = syn.synth_rows(kind, 64, steps + warmup, config_id=3)
    rows_dev = torch.from_numpy(np.ascontiguousarray(np.tile(base, (B // 64, 1, 1)))).cuda()
    frames = [rows_dev[:, f:f + 1].contiguous() for f in range(steps + warmup)]
    for f in range(warmup):
        be.step_device(frames[f], raw_ready=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for f in range(warmup, warmup + steps):
        be.step_device(frames[f], raw_ready=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    time.sleep(0.5)                                           # the roofline leg times launches ALONE against the burst peak: let the
    layer_ms, acc = np.zeros(L, np.float32), np.zeros(L, np.float64)   # board leave the power cap the throughput leg drove it into
    for f in range(5):
        be.step_device(frames[warmup + f], layer_ms=layer_ms)
        acc += layer_ms
    torch.cuda.synchronize()
    acc /= 5
    # end to end: host rows in (pinned H2D), host results out (pinned D2H), every step inside the timed region
    host_frames = [np.ascontiguousarray(np.tile(base[:, f:f + 1], (B // 64, 1, 1))) for f in range(steps + warmup)]
    be.reset()                                                # (after the roofline leg: a long loaded run lowers the clocks of what follows)
    for f in range(warmup):
        be.step(host_frames[f])
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    pending, checksum = [], 0.0
    for f in range(warmup, warmup + steps):
        pending.append(be.submit(host_frames[f]))
        if len(pending) >= be.N_SLOTS - 1:
            checksum += float(pending.pop(0).result().msg[0, 0, 4])
    for p in pending:
        checksum += float(p.result().msg[0, 0, 4])
    g1.record()
    torch.cuda.synchronize()
    e2e_ms = g0.elapsed_time(g1)
    peaks, _ = measured_peaks()
    dom_ms = float(np.mean(acc[1:-1])) if L > 2 else float(acc[-1])
    flops = B * n * T * 2 * 4 * H * (2 * H)
    tensor = be.lstm_variant == "tc"
    peak = peaks["bf16_tflops"] if tensor else 148 * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12
    return {"config": WORKLOAD_TEXT[workload], "model": {"I": I, "H": H, "L": L, "T": T, "O": O}, "value": B * steps / (ms * 1e-3), "unit": UNIT,
            "ms_per_step": ms / steps, "steps": steps, "lstm_variant": be.lstm_variant, "tc_probe_error_m": be.tc_probe_error_m,
            "e2e": {"value": B * steps / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / steps, "h2d_bytes_per_step": be.h2d_bytes_per_frame,
                    "d2h_bytes_per_step": be.d2h_bytes_per_frame, "checksum": checksum},
            "roofline": {"bound": "tensor" if tensor else "fp32_ffma", "kernel": tc_kernel_name(H) if tensor else "lstm_layer_fma_kernel",
                         "achieved": flops / (dom_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / (dom_ms * 1e-3) / 1e12 / peak, "flops_per_launch": flops, "layer_ms": [float(v) for v in acc]}}


def realtime_latency(BatchedEstimator, N, syn, frames=300):
    """BASELINE configs[1]: watch+phone pocket LSTM estimator, 1 stream x 100 MC samples, per-frame latency through the
    host-facing call (one 55-float row in pinned memory -> message + std + 100 sample positions back on the host)."""
    kind = syn.KIND_POCKET
    spec = syn.kind_spec(kind)
    state = syn.synth_state_dict(spec["I"], spec["H"], spec["L"], spec["O"], 1234 + kind)
    be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=spec["T"], y_targets=spec["y_targets"],
                          stats=spec["stats"], n_streams=1, mc_samples=100, smooth=1, dropout=spec["p"], frames_per_call=1,
                          mask_mode=N.MASK_PHILOX, philox_seed=7)
    rows = syn.synth_rows(kind, 1, frames + 20, config_id=2)
    out = {"workload": "watch+phone pocket LSTM estimator (I22 H256 L2 T6 O14), 1 stream x 100 MC samples, frame by frame",
           "lstm_variant": be.lstm_variant, "frames": frames}
    for key, fn, what in (("", be.step_graph, "BatchedEstimator.step_graph (one CUDA-graph launch: H2D + 3 stages + D2H, then sync)"),
                          ("eager_", be.step, "BatchedEstimator.step (H2D + 3 stages + D2H enqueued call by call, then sync)")):
        be.reset()
        lat = []
        for f in range(frames + 20):
            t0 = time.perf_counter()
            fn(rows[:, f:f + 1])
            lat.append(time.perf_counter() - t0)
        lat = np.asarray(lat[20:]) * 1e3
        out.update({key + "p50_ms": float(np.percentile(lat, 50)), key + "p99_ms": float(np.percentile(lat, 99)),
                    key + "timing": "host wall clock around " + what})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="uarm_1024x100", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-frames", type=int, default=400, help="frames per worker of the bounded cpu_baseline sample")
    ap.add_argument("--ref-frames", type=int, default=50, help="--impl reference: frames per worker per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-realtime", action="store_true", help="skip the configs[1] single-stream latency leg")
    ap.add_argument("--no-other-models", action="store_true", help="skip the throughput legs of the two H = 256 models")
    ap.add_argument("--lstm", default="auto", choices=["auto", "fp32", "tc"], help="LSTM kernel variant (auto: probe-gated tensor cores)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the ONE JSON line and nothing else: whatever libraries write to file descriptor 1 (NCCL prints its
    # version banner there) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world != args.gpus and world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
