"""Paths and port constants (counterpart of ``src/wear_mocap_ape/config.py:5-19``).

``PATHS["deploy"]`` is read at call time by the model loader and the stats loader, exactly like the
reference (``nn_models.py:379``, ``data_stats.py:31``): point it at a ``wear_mocap_ape`` ``data_deploy``
directory that holds real ``checkpoint.pt`` files to run deployed weights.
"""
from pathlib import Path

_pkg = Path(__file__).parent.absolute()

PATHS = {
    "deploy": _pkg / "data_deploy",
    "skeleton": _pkg / "data_deploy",
}

# the surrounding socket code of the reference keeps using these
PORT_PUB_LEFT_ARM = 50003
PORT_LISTEN_WATCH_PHONE_IMU = 65000
PORT_LISTEN_WATCH_IMU = 46000
