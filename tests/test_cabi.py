"""CPU: the C-ABI library loads and exports every symbol include/ape_b200.h declares (no compute calls)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200.estimate import nn_models

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "ape_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ape_[a-z0-9_]+)\s*\(", text)))


def test_header_matches_binding_table():
    assert declared_symbols() == sorted(N.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = N.load()
    for name in declared_symbols():
        assert getattr(lib, name) is not None, name
    assert lib.ape_abi_version() == 4


def test_struct_layout_matches_header():
    # field order of struct ape_lstm_args in the header == ctypes Structure
    text = (ROOT / "include" / "ape_b200.h").read_text()
    body = re.search(r"typedef struct ape_lstm_args \{(.*?)\} ape_lstm_args;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", part.strip())[0])
    assert names == [f[0] for f in N.LstmArgs._fields_]


def test_no_gpu_calls_fail_loudly_not_silently():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = nn_models.DropoutLSTM(4, 32, 2, 12)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.forward(torch.zeros(1, 3, 4))


def test_blob_size_mirror():
    for (I, H, L, O) in [(20, 256, 2, 12), (22, 256, 2, 14), (38, 128, 3, 12), (5, 32, 1, 20), (16, 64, 4, 14)]:
        assert N.blob_floats(I, H, L, O) == nn_models.packed_floats(I, H, L, O)


def test_bad_arguments_are_refused():
    lib = N.load()
    assert lib.ape_features(None, 0, 0, None, None, 0, None, 1, 1, 0, None, 1, None) == N.APE_ERR_BAD_ARG
    assert lib.ape_mc_lstm_fma(None, None) == N.APE_ERR_BAD_ARG
    assert lib.ape_fk_reduce(None, 1, None, None, None, 0, 12, 1, 1, 0, None, 1, 1, None, None, None, None, None, None) == N.APE_ERR_BAD_ARG
    out = ctypes.c_uint64(0)
    assert lib.ape_mc_lstm_workspace_bytes(38, 100, 3, 6, 12, 4, 10, ctypes.byref(out)) == N.APE_ERR_UNSUPPORTED   # H % 32 != 0
    assert lib.ape_mc_lstm_workspace_bytes(38, 128, 3, 6, 12, 1024, 100, ctypes.byref(out)) == N.APE_OK and out.value > 0


def test_product_package_never_touches_oracle_or_selfcheck_hooks():
    pkg = ROOT / "arm_pose_estimation_b200"
    for py in pkg.rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py
        if py.name != "_native.py":
            assert "ape_selfcheck" not in src and "ape_selftest" not in src, py


def test_sass_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", str(N.lib_path())], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
