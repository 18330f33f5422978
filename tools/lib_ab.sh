#!/bin/bash
# A/B of library builds on one GPU box: bash tools/lib_ab.sh "<lib.so> <lib.so> ..." [workload] [repeats]
# Prints device-resident estimates/s, end-to-end estimates/s and the per-layer times of each build (bench.py legs).
cd "$(dirname "$0")/.."
for rep in $(seq 1 ${3:-2}); do
  for lib in $1; do
    echo -n "$lib: "
    APE_B200_LIB=$PWD/$lib python bench.py --workload ${2:-uarm_1024x100} --steps 100 --warmup 5 --no-cpu-baseline --no-realtime --no-other-models --no-sustained --no-relabel 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(round(d['value']),round(d['e2e']['value']),[round(v,4) for v in d['roofline']['layer_ms']],d['clocks'])"
  done
done
