"""BASELINE configs[4]: MC-sample / stream sweep mapping the FFMA vs tensor-core crossover of the LSTM stage.
Device time of the LSTM stage alone (all layers, CUDA events around ape_mc_lstm_*): the fp32 FFMA kernel, the tensor-core layer kernels
and (calls of <= 8192 rows) the cluster kernel (tc_flags = 4: one 8-CTA cluster per 128 rows), same Philox masks.
Writes a markdown table (stdout) - committed as profiles/r1_crossover.md (round 1) / r2_crossover.md (round 2)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from arm_pose_estimation_b200 import _native as N, synthetic as syn
from arm_pose_estimation_b200.estimate import nn_models

lib = N.load()
max_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20


def lstm_ms(spec, w32, wtc, x, n, variant, reps):
    E = x.shape[0]
    I, H, L, T, O = (spec[k] for k in "IHLTO")
    ws = torch.empty(N.workspace_bytes(I, H, L, T, O, E, n, tensor_core=(variant != "fp32")) + 4096, dtype=torch.uint8, device="cuda")
    preds = torch.zeros((E, 1, n, O), dtype=torch.float32, device="cuda")
    a = N.LstmArgs()
    a.weights, a.weights_tc = w32.data_ptr(), wtc.data_ptr()
    a.I, a.H, a.L, a.T, a.O = I, H, L, T, O
    a.dropout_p = spec["p"]
    a.x_dense, a.feat_ring_buf, a.feat_ring = x.data_ptr(), None, 0
    a.B, a.nF, a.frame0, a.n_samples = E, 1, 0, n
    a.mask_mode, a.philox_seed, a.stream_id0 = N.MASK_PHILOX, 0x5EED, 0
    a.workspace = ws.data_ptr()
    a.preds, a.pred_ring, a.all_steps = preds.data_ptr(), 1, 0
    fn = lib.ape_mc_lstm_tc if variant != "fp32" else lib.ape_mc_lstm_fma
    a.tc_flags = 4 if variant == "cluster" else 0
    st = N.current_stream_ptr()
    for _ in range(2):
        N.check(fn(a, st), variant)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        N.check(fn(a, st), variant)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for kind in (syn.KIND_UARM, syn.KIND_POCKET):
    spec = syn.kind_spec(kind)
    I, H, L, T, O = (spec[k] for k in "IHLTO")
    state = syn.synth_state_dict(I, H, L, O, 1234 + kind)
    w32 = torch.from_numpy(nn_models.pack_lstm_weights(state)).cuda()
    wtc = torch.from_numpy(nn_models.pack_lstm_weights_tc(state)).cuda()
    print(f"\n### {syn.KIND_NAMES[kind]} model (I{I} H{H} L{L} T{T} O{O}): LSTM stage, device ms per call  fp32 FFMA | tcgen05 layer kernels | cluster kernel  (fp32 / best tensor-core)\n")
    ns = (1, 4, 16, 64, 100, 256, 1024)
    print("| streams \\ MC samples | " + " | ".join(str(n) for n in ns) + " |")
    print("|---|" + "---|" * len(ns))
    g = torch.Generator(device="cpu").manual_seed(1)
    for B in (1, 4, 16, 64, 256, 1024, 4096, 16384, 65536):
        x = torch.randn((B, T, I), generator=g, dtype=torch.float32).cuda()
        cells = []
        for n in ns:
            rows = B * n
            if rows > max_rows:
                cells.append("-")
                continue
            reps = 20 if rows <= 1 << 14 else (5 if rows <= 1 << 18 else 2)
            f, t = lstm_ms(spec, w32, wtc, x, n, "fp32", reps), lstm_ms(spec, w32, wtc, x, n, "tc", reps)
            c = lstm_ms(spec, w32, wtc, x, n, "cluster", reps) if rows <= 8192 else None
            best = t if c is None else min(t, c)
            cells.append(f"{f:.3f} \\| {t:.3f} \\| {'-' if c is None else format(c, '.3f')} ({f / best:.1f}x)")
        print(f"| {B} | " + " | ".join(cells) + " |", flush=True)
