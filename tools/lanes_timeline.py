"""CTA timeline of the big layer kernels over a few consecutive pipelined calls: which SM ran which layer launch when.

    python tools/lanes_timeline.py [workload] [lanes 0|1] > gpurun_out/timeline.txt

Every CTA of every tensor-core layer launch stamps %globaltimer at entry, after its set-up and at exit (ape_lstm_args.trace with
trace_layer < 0).  Prints, per call and layer, the launch's first entry / last exit relative to the first stamped call, the
spread of entry and exit times, and how much of the (SMs x wall time) rectangle of the traced calls was covered by CTAs.
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from arm_pose_estimation_b200 import _native as N, synthetic as syn            # noqa: E402
from arm_pose_estimation_b200.estimate.batched import BatchedEstimator          # noqa: E402
from bench import WORKLOADS                                                     # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "uarm_1024x100"
    lanes = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
    kind, B, n, smooth = WORKLOADS[workload]
    spec = syn.kind_spec(kind)
    I, H, L, T, O = (spec[k] for k in "IHLTO")
    state = syn.synth_state_dict(I, H, L, O, 1234 + kind)
    rows = syn.synth_rows(kind, min(64, B), 8, config_id=3, first_stream=0)
    rows = np.ascontiguousarray(np.tile(rows, (-(-B // rows.shape[0]), 1, 1))[:B])
    rows_dev = torch.from_numpy(rows).cuda()
    frames = [rows_dev[:, f:f + 1].contiguous() for f in range(8)]
    be = BatchedEstimator(kind=kind, layout=spec["layout"], state=state, seq_len=T, y_targets=spec["y_targets"],
                          stats=spec["stats"], n_streams=B, mc_samples=n, smooth=smooth, dropout=spec["p"],
                          mask_mode=N.MASK_PHILOX, philox_seed=2026, lstm_variant="tc", lanes=lanes)
    NC = 8
    tl = torch.zeros((NC, L, 160, 4), dtype=torch.int64, device="cuda")
    for f in range(10):
        be.step_device(frames[f % 8], raw_ready=True)
    for k in range(NC):
        be.step_device(frames[k % 8], raw_ready=True, timeline=tl[k])
    torch.cuda.synchronize()
    t = tl.cpu().numpy()
    used = t[..., 0] > 0
    t0 = t[..., 0][used].min()
    print(f"# {workload} lanes={lanes}: times in us relative to the first stamped entry")
    print("# call layer  CTAs  first_entry last_entry  first_exit last_exit  mean_setup_us  mean_busy_us")
    for k in range(NC):
        for l in range(L):
            u = used[k, l]
            if not u.any():
                continue
            en, su, ex = [(t[k, l, :, i][u] - t0) / 1e3 for i in range(3)]
            print(f"{k:5d} {l:5d} {int(u.sum()):5d} {en.min():11.1f} {en.max():10.1f} {ex.min():11.1f} {ex.max():9.1f} {np.mean(su - en):14.2f} {np.mean(ex - en):13.1f}")
    # coverage of the big layers (>= 1) over calls 2 .. NC-2 (steady state)
    sel = used[2:NC - 1, 1:]
    en = t[2:NC - 1, 1:, :, 0][sel]
    ex = t[2:NC - 1, 1:, :, 2][sel]
    span = (ex.max() - en.min()) / 1e3
    busy = float(np.sum(ex - en)) / 1e3
    n_sm = int(t[..., 3][used].max()) + 1
    print(f"# steady state (calls 2..{NC - 2}, layers >= 1): span {span:.1f} us = {span / (NC - 3):.1f} us per call; CTA-busy {busy / n_sm:.1f} us per SM "
          f"= {busy / n_sm / span:.3f} of the span ({n_sm} SMs)")
    # per-SM gaps between consecutive big-layer CTAs
    evs = {}
    for k in range(NC):
        for l in range(1, L):
            for c in np.nonzero(used[k, l])[0]:
                evs.setdefault(int(t[k, l, c, 3]), []).append((int(t[k, l, c, 0]), int(t[k, l, c, 2]), k, l))
    gaps = []
    for sm, lst in evs.items():
        lst.sort()
        for (e0, x0, *_), (e1, x1, *_) in zip(lst, lst[1:]):
            gaps.append((e1 - x0) / 1e3)
    gaps = np.array(gaps)
    print(f"# per-SM gap between one big-layer CTA's exit and the next one's entry: median {np.median(gaps):.1f} us, mean {gaps.mean():.1f}, p90 {np.percentile(gaps, 90):.1f}, max {gaps.max():.1f}")


if __name__ == "__main__":
    main()
