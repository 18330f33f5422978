"""Wire layouts of the IMU rows that enter the estimation path.

Same keys and float positions as ``WATCH_ONLY_IMU_LOOKUP`` (28 floats,
``src/wear_mocap_ape/data_types/messaging.py:20-68``) and ``WATCH_PHONE_IMU_LOOKUP`` (55 floats,
``messaging.py:96-187``) of the reference; built here from per-device field groups.  The CUDA feature
kernel addresses the same positions (``csrc/ape_features.cuh``); ``tests/test_oracle_golden.py`` pins the tables
against ``tests/golden/tables.json`` (exported from the reference).
"""


def _device_block(p):
    # one device's sensor block as it is sent by the watch / phone app (23 floats)
    return (
        [f"{p}_dt", f"{p}_h", f"{p}_m", f"{p}_s", f"{p}_ns"]
        + [f"{p}_rotvec_{c}" for c in ("w", "x", "y", "z", "conf")]
        + [f"{p}_gyro_{a}" for a in "xyz"]
        + [f"{p}_lvel_{a}" for a in "xyz"]
        + [f"{p}_lacc_{a}" for a in "xyz"]
        + [f"{p}_pres"]
        + [f"{p}_grav_{a}" for a in "xyz"]
    )


def _forward(p):
    return [f"{p}_forward_{c}" for c in "wxyz"]


def _index(fields):
    return {name: pos for pos, name in enumerate(fields)}


WATCH_ONLY_IMU_LOOKUP = _index(_device_block("sw") + _forward("sw") + ["sw_init_pres"])
watch_only_imu_msg_len = len(WATCH_ONLY_IMU_LOOKUP) * 4  # bytes

WATCH_PHONE_IMU_LOOKUP = _index(
    _device_block("sw") + _device_block("ph") + _forward("sw") + _forward("ph") + ["sw_init_pres"]
)
watch_phone_imu_msg_len = len(WATCH_PHONE_IMU_LOOKUP) * 4  # bytes

# ids shared with the C-ABI (include/ape_b200.h: APE_LAYOUT_*)
LAYOUT_WATCH_ONLY = 0
LAYOUT_WATCH_PHONE = 1
LAYOUT_NCOLS = {LAYOUT_WATCH_ONLY: len(WATCH_ONLY_IMU_LOOKUP), LAYOUT_WATCH_PHONE: len(WATCH_PHONE_IMU_LOOKUP)}
