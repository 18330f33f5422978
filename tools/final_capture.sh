#!/bin/bash
# One GPU call that produces a round's measured artefacts under gpurun_out/<prefix>_*: bash tools/final_capture.sh r2j
# (tests -> bench legs without a profiler -> ncu launch list -> one `ncu --set full` capture per kernel)
P=${1:-rX}
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/${P}_pytest.log 2>&1; tail -2 $O/${P}_pytest.log
python bench.py > $O/${P}_bench.json 2> $O/${P}_bench.err
python bench.py --steps 20 --warmup 3 > $O/${P}_bench_steps20.json 2> /dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $O/${P}_reference_arm_bench.json 2> /dev/null
LIGHT="--steps 3 --warmup 1 --no-cpu-baseline --no-realtime --no-other-models --no-sustained --no-relabel"
NCU="ncu --set full --clock-control none --import-source on"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${P}_uarm_launches.csv python bench.py $LIGHT > /dev/null 2>&1
$NCU --kernel-name-base mangled -k regex:fk_reduce_kernelILi0 -c 1 -f -o $O/${P}_fk12 python tools/fk_bench.py > /dev/null 2>&1
$NCU --kernel-name-base mangled -k regex:fk_reduce_kernelILi1 -c 1 -f -o $O/${P}_fk14 python tools/fk_bench.py > /dev/null 2>&1
$NCU -k regex:lstm_pair_tcw -c 1 -f -o $O/${P}_tcw python bench.py $LIGHT > /dev/null 2>&1
$NCU -k regex:lstm_small -c 1 -f -o $O/${P}_tcl python tools/lat_breakdown.py > /dev/null 2>&1
$NCU -k regex:mc_ff_kernel -c 1 -f -o $O/${P}_ff python -c "
import sys; sys.path.insert(0, '.')
import torch, bench
from arm_pose_estimation_b200 import _native as N
bench.ff_leg(N, torch, reps=2)" > /dev/null 2>&1
python tools/lat_host_breakdown.py > $O/${P}_latency_breakdown.txt 2>&1
ls -la $O | grep ${P}_
