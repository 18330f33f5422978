#!/bin/bash
# Timing experiments on the tensor-core LSTM kernel: rebuild the library with one cost removed (results are then wrong;
# only the layer times matter) and run the bench's LSTM leg.  Usage (on a GPU box): bash tools/tc_experiments.sh "1 2 3 4 5"
set -e
cd "$(dirname "$0")/.."
SRCS=$(ls arm_pose_estimation_b200/csrc/*.cu)
for e in ${1:-0 1 2 3 4 5}; do
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -DAPE_EXP=$e -o /tmp/libape_exp$e.so $SRCS
  echo -n "APE_EXP=$e: "
  APE_B200_LIB=/tmp/libape_exp$e.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline --lstm tc 2>&1 | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value']),[round(v,4) for v in d['roofline']['layer_ms']])"
done
