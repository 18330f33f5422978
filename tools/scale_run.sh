#!/bin/bash
# The driver's 1 -> 8 scaling run on one box: bash tools/scale_run.sh "1 2 4 8" [extra bench flags] -> gpurun_out/scale_n<N>.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in ${1:-1 2 4 8}; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 ${@:2} > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n ${@:2} > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - <<PY
import json
d = json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
print($n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "sustained", round(d.get("sustained", {}).get("value_all_gpus", 0)),
      "relabel", round(d.get("relabel", {}).get("value", 0)), d["clocks"])
PY
done
