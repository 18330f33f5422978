"""Stage 3 standalone against the HBM roofline (bench.py's fk_roofline leg on its own): python tools/fk_bench.py [E] [n]
APE_B200_LIB selects the library build."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from arm_pose_estimation_b200 import _native as N, synthetic as syn            # noqa: E402
import bench                                                                    # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for kind in (syn.KIND_UARM, syn.KIND_POCKET):
    r = bench.fk_standalone(N, syn, torch, kind, n, bench.measured_peaks()[0], E=E)
    print(json.dumps({k: r[k] for k in ("ms_per_launch", "achieved", "frac", "bytes_per_estimate")} | {"kind": syn.KIND_NAMES[kind]}))
