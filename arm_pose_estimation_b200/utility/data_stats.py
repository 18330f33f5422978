"""Load the per-column normalisation constants (load branch of ``utility/data_stats.py:30-40``).

Looks for ``<deploy>/data_stats/<X>_<Y>.pkl`` first (the reference's own file, so a reference deploy
directory works unchanged) and falls back to the JSON export shipped with this package.  Creating stats
from training data (``data_stats.py:48-188``) is training-side and out of scope.
"""
import json
import logging
import pickle
from pathlib import Path

import numpy as np

from arm_pose_estimation_b200 import config
from arm_pose_estimation_b200.utility.names import NNS_INPUTS, NNS_TARGETS


def get_norm_stats(x_inputs: NNS_INPUTS, y_targets: NNS_TARGETS, data_list: list = None) -> dict:
    stem = "{}_{}".format(x_inputs.name, y_targets.name)
    f_dir = Path(config.PATHS["deploy"]) / "data_stats"
    pkl, jsn = f_dir / (stem + ".pkl"), f_dir / (stem + ".json")
    if pkl.exists():
        with open(pkl, "rb") as handle:
            dat = pickle.load(handle)
        logging.info("loaded data stats from {}".format(pkl))
    elif jsn.exists():
        dat = json.loads(jsn.read_text())
        logging.info("loaded data stats from {}".format(jsn))
    else:
        raise UserWarning(f"no data stats found for {stem} in {f_dir}")
    for k in ("xx_m", "xx_s", "yy_m", "yy_s"):
        dat[k] = np.asarray(dat[k], dtype=np.float64)
    return dat
