"""Run HERE only (needs /root/reference): export the reference's deploy constants into the package.

* ``data_deploy/nn/<hash>/results.json`` - the hyper-parameter keys the loader reads
  (``nn_models.py:389-408`` of the reference), copied verbatim as data.
* ``data_deploy/data_stats/<X>_<Y>.json`` - ``xx_m/xx_s/yy_m/yy_s`` of the reference's pickles
  (``data_stats.py:30-40``) as float64 ``repr`` round-trip JSON.
"""
import json
import pickle
import sys
import warnings
from pathlib import Path

REF = Path("/root/reference/src/wear_mocap_ape/data_deploy")
DST = Path(__file__).resolve().parents[2] / "arm_pose_estimation_b200" / "data_deploy"
KEEP = ["model", "hidden_layer_count", "hidden_layer_size", "dropout", "sequence_len", "normalize", "hash",
        "y_targets_n", "x_inputs_n", "y_targets_v", "x_inputs_v"]


def main():
    warnings.filterwarnings("ignore")
    for rj in sorted(REF.glob("nn/*/results.json")):
        params = json.loads(rj.read_text())
        out = DST / "nn" / rj.parent.name / "results.json"
        out.parent.mkdir(parents=True, exist_ok=True)
        out.write_text(json.dumps({k: params[k] for k in KEEP}, indent=1) + "\n")
        print("wrote", out)
    for pk in sorted(REF.glob("data_stats/*.pkl")):
        with open(pk, "rb") as fh:
            d = pickle.load(fh)
        out = DST / "data_stats" / (pk.stem + ".json")
        out.parent.mkdir(parents=True, exist_ok=True)
        js = {k: [float(v) for v in d[k]] for k in ("xx_m", "xx_s", "yy_m", "yy_s")}
        js["x_inputs"], js["y_targets"] = list(d["x_inputs"]), list(d["y_targets"])
        out.write_text(json.dumps(js, indent=1) + "\n")
        print("wrote", out)


if __name__ == "__main__":
    sys.exit(main())
