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
    assert lib.ape_abi_version() == 6


def test_struct_layout_matches_header():
    # field order of struct ape_lstm_args in the header == ctypes Structure
    text = (ROOT / "include" / "ape_b200.h").read_text()
    assert _struct_fields(text, "ape_lstm_args") == [f[0] for f in N.LstmArgs._fields_]


def _struct_fields(text, name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[[^\]]*\])?\s*$", part.strip())[0])
    return names


def test_pipeline_desc_layout_matches_header():
    # struct ape_pipeline_desc (ABI 6) embeds ape_lstm_args by value: field order and the array extents must agree with ctypes
    text = (ROOT / "include" / "ape_b200.h").read_text()
    assert _struct_fields(text, "ape_pipeline_desc") == [f[0] for f in N.PipelineDesc._fields_]
    assert int(re.search(r"#define APE_PIPELINE_MAX_SLOTS (\d+)", text).group(1)) == N.PIPELINE_MAX_SLOTS
    flags = {k: int(v) for k, v in re.findall(r"#define (APE_PIPE_[A-Z0-9_]+) +(\d+)", text)}
    assert flags == {"APE_PIPE_INPUT_PENDING": N.PIPE_INPUT_PENDING, "APE_PIPE_CALLER_WAITS": N.PIPE_CALLER_WAITS, "APE_PIPE_D2H": N.PIPE_D2H}
    d = N.PipelineDesc()
    assert ctypes.sizeof(d.out_dev) == 8 * N.PIPELINE_MAX_SLOTS and ctypes.sizeof(d.frames_dev) == 8 * 4
    assert N.PipelineDesc.lstm.size == ctypes.sizeof(N.LstmArgs)
    lib = N.load()
    assert lib.ape_pipeline_create(None, None) == N.APE_ERR_BAD_ARG
    assert lib.ape_pipeline_submit(None, None, None, 1, 0, None, 0, None, None) == N.APE_ERR_BAD_ARG
    assert lib.ape_pipeline_destroy(None) == N.APE_OK


def test_split_precision_blob_layout():
    # pack_lstm_weights_tcx: per layer 2 CTAs x 4 chunks x [Wx_hi + bias K step | Wx_lo | Wh_hi | Wh_lo]; hi + lo reproduce the (halved)
    # weights to ~2^-20, the bias rides in rows kin_pad / kin_pad + 1 of the hi tile, and the size matches the library's
    import numpy as np
    from arm_pose_estimation_b200 import synthetic as syn
    I, H, L, O = 38, 128, 3, 12
    state = syn.synth_state_dict(I, H, L, O, 3)
    blob = nn_models.pack_lstm_weights_tcx(state)
    assert blob.size == N.tcx_blob_bytes(I, H, L) and N.tcx_supported(I, H, L, O) and not N.tcx_supported(20, 256, 2, 12)
    halfs = blob.view(np.float16)
    kin_pad = 48
    p0, p1, ph = (kin_pad + 16) * 64, kin_pad * 64, H * 64                # halfs per piece of layer 0
    # layer 0, CTA 0, chunk 0: columns n = 4 u + g for units 0..15
    hi = halfs[:p0].reshape((kin_pad + 16) // 8, 64, 8).transpose(1, 0, 2).reshape(64, kin_pad + 16).astype(np.float64)
    lo = halfs[p0:p0 + p1].reshape(kin_pad // 8, 64, 8).transpose(1, 0, 2).reshape(64, kin_pad).astype(np.float64)
    n = np.arange(64)
    rows = (n % 4) * H + n // 4
    scale = np.where(n % 4 != 2, 0.5, 1.0)
    w = state["lstm.weight_ih_l0"][rows].astype(np.float64) * scale[:, None]
    assert np.abs(hi[:, :I] + lo[:, :I] - w).max() <= 2.0 ** -20 * np.abs(w).max() and not hi[:, I:kin_pad].any()
    b = (state["lstm.bias_ih_l0"] + state["lstm.bias_hh_l0"])[rows].astype(np.float64) * scale
    assert np.abs(hi[:, kin_pad] + hi[:, kin_pad + 1] - b).max() <= 2.0 ** -20 * np.abs(b).max() and not hi[:, kin_pad + 2:].any()
    wh_hi = halfs[p0 + p1:p0 + p1 + ph].reshape(H // 8, 64, 8).transpose(1, 0, 2).reshape(64, H).astype(np.float64)
    wh_lo = halfs[p0 + p1 + ph:p0 + p1 + 2 * ph].reshape(H // 8, 64, 8).transpose(1, 0, 2).reshape(64, H).astype(np.float64)
    wh = state["lstm.weight_hh_l0"][rows].astype(np.float64) * scale[:, None]
    assert np.abs(wh_hi + wh_lo - wh).max() <= 2.0 ** -20 * np.abs(wh).max()


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


@pytest.mark.parametrize("kgx", [2, 4, 6, 32])
def test_tcs_schedule_covers_every_operand_once(kgx):
    """The H = 256 tensor-core kernel streams its weights in the order of a host-built table (csrc/ape_lstm_tcs.cu: walk_step).
    Every chunk must see its x k-groups [0, kgx) and - after the first step - its recurrent k-groups [0, 32) exactly once, the
    first MMA of a chunk must overwrite the accumulator, recurrent pieces may only need K-slices that are published by then,
    each accumulator slot's last x-part must release the x tile, and a piece must fit two ring slots."""
    import ctypes as C
    lib = N.load()
    NCH, SLICE_KG, SLOT_KG = 8, 4, 16
    for first_step in (1, 0):
        buf = (C.c_uint32 * (2 * 64))()
        n = C.c_int(0)
        assert lib.ape_selfcheck_tcs_schedule(kgx, first_step, buf, 64, C.byref(n)) == N.APE_OK
        ent = [(buf[2 * i], buf[2 * i + 1]) for i in range(n.value)]
        cover_x = {c: [] for c in range(NCH)}
        cover_h = {c: [] for c in range(NCH)}
        begun, done, x_done = [], [], []
        for x, y in ent:
            slot, is_h, first = x & 1, bool(x & 2), bool(x & 4)
            nmma, hneed, src = (x >> 8) & 0x1F, (x >> 13) & 0xF, x >> 17
            assert 1 <= nmma <= 2 * (SLOT_KG // 2)
            chunk_rel, off = divmod(src, kgx + 32)                     # k-group offset inside this CTA's weight tiles
            assert chunk_rel % 2 == slot
            if x & 8:
                begun.append(chunk_rel)
            if is_h:
                kg0 = off - kgx
                assert kg0 >= 0 and y == kg0 * 4 and not first
                assert hneed == -(-(kg0 + 2 * nmma) // SLICE_KG)       # needs exactly the K-slices it multiplies
                cover_h[chunk_rel] += list(range(kg0, kg0 + 2 * nmma))
            else:
                assert off + 2 * nmma <= kgx and y == off * 128 and hneed == 0
                assert first == (off == 0)
                cover_x[chunk_rel] += list(range(off, off + 2 * nmma))
            if x & 16:
                done.append(chunk_rel)
            if x & 32:
                x_done.append(chunk_rel)
        for c in range(NCH):
            assert sorted(cover_x[c]) == list(range(kgx))
            assert sorted(cover_h[c]) == ([] if first_step else list(range(32)))
        assert begun == list(range(NCH)) and sorted(done) == list(range(NCH)) and x_done == [NCH - 2, NCH - 1]
        # the previous tile's output-layer product goes where accumulator slot 1 is first refilled: before chunk 1's first piece of
        # a tile's first step, and nowhere else
        out_before = [i for i, (x, _) in enumerate(ent) if x & 64]
        if first_step:
            assert len(out_before) == 1
            x = ent[out_before[0]][0]
            assert (x >> 17) // (kgx + 32) == 1 and x & 8 and x & 4 and not x & 2
        else:
            assert out_before == []
        if not first_step:                                              # only ONE small piece may follow the last K-slice of h_t
            last_slice = [i for i, (x, _) in enumerate(ent) if (x >> 13) & 0xF == NCH]
            assert ((ent[last_slice[0]][0] >> 8) & 0x1F) == SLICE_KG // 2 and ent[last_slice[0]][0] & 16
    assert lib.ape_selfcheck_tcs_schedule(3, 0, buf, 64, C.byref(n)) == N.APE_ERR_BAD_ARG
