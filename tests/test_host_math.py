"""CPU: the kernels' __host__ __device__ row math (compiled for the host in csrc/ape_selfcheck.cu) against the
Random123 known-answer vectors and the golden fixtures generated from the reference."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden
from arm_pose_estimation_b200 import _native as N
from arm_pose_estimation_b200 import synthetic as syn
from oracle import fk as OFK

KINDS = (syn.KIND_WATCH_ONLY, syn.KIND_POCKET, syn.KIND_UARM)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    assert N.load().ape_selfcheck_philox(c, k, o) == 0
    return [int(v) for v in o]


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_keep_bits_follow_the_documented_counter_layout():
    lib = N.load()
    seed, stream, frame, sample, gap, t, group, p = 0x0123456789ABCDEF, 7, 42, 99, 1, 5, 3, 0.2
    out = C.c_uint32(0)
    assert lib.ape_selfcheck_keep8(seed, stream, frame, sample, gap, t, group, p, C.byref(out)) == 0
    words = philox([stream, frame, sample | (gap << 20) | (t << 24), group], [seed & 0xFFFFFFFF, seed >> 32])
    thr = int((1.0 - np.float32(p)) * 65536.0 + 0.5)
    bits = 0
    for i, w in enumerate(words):
        bits |= int((w & 0xFFFF) < thr) << (2 * i)
        bits |= int((w >> 16) < thr) << (2 * i + 1)
    assert out.value == bits
    # keep rate over many draws ~ 1 - p
    n, kept = 4000, 0
    for g in range(n):
        lib.ape_selfcheck_keep8(seed, 0, 0, 0, 0, 0, g, p, C.byref(out))
        kept += bin(out.value).count("1")
    assert abs(kept / (8 * n) - 0.8) < 0.01
    lib.ape_selfcheck_keep8(seed, 0, 0, 0, 0, 0, 0, 0.0, C.byref(out))
    assert out.value == 0xFF                                   # p = 0 keeps everything


@pytest.mark.parametrize("kind", KINDS)
def test_stage1_row_math_against_reference_features(kind):
    g = load_golden(f"features_{syn.KIND_NAMES[kind]}.npz")
    lib, layout = N.load(), syn.KIND_LAYOUT[kind]
    xx = (C.c_double * 38)()
    nfeat = C.c_int(0)
    for row, want in zip(g["rows"], g["xx"]):
        r = np.ascontiguousarray(row, dtype=np.float32)
        assert lib.ape_selfcheck_features(kind, layout, r.ctypes.data_as(C.POINTER(C.c_float)), xx, C.byref(nfeat)) == 0
        assert nfeat.value == want.shape[0]
        np.testing.assert_allclose(np.array(xx[: nfeat.value]), want.astype(np.float64), rtol=0, atol=1e-12)


def test_stage1_watch_columns_from_the_phone_layout():
    # WatchOnlyNN(watch_phone=True): same features from the 55-float layout (watch_only.py:28-31)
    rows55 = syn.synth_rows(syn.KIND_UARM, 1, 5, config_id=3)[0]
    from arm_pose_estimation_b200.data_types import messaging as M
    lib = N.load()
    a, b = (C.c_double * 38)(), (C.c_double * 38)()
    n = C.c_int(0)
    for r55 in rows55:
        r28 = np.zeros(28, np.float32)
        for k, pos in M.WATCH_ONLY_IMU_LOOKUP.items():
            r28[pos] = r55[M.WATCH_PHONE_IMU_LOOKUP[k]]
        lib.ape_selfcheck_features(0, 0, r28.ctypes.data_as(C.POINTER(C.c_float)), a, C.byref(n))
        lib.ape_selfcheck_features(0, 1, np.ascontiguousarray(r55).ctypes.data_as(C.POINTER(C.c_float)), b, C.byref(n))
        assert list(a[:20]) == list(b[:20])


@pytest.mark.parametrize("use_float,tol", [(0, 1e-11), (1, 3e-5)])
def test_stage3_row_math_against_reference_est_rows(body9, use_float, tol):
    g = load_golden("fk.npz")
    lib = N.load()
    body = np.ascontiguousarray(body9.ravel(), dtype=np.float64)
    for ti, tname in enumerate(OFK.TARGETS):
        W = 14 if ti == 0 else 21
        est = (C.c_double * 23)()
        bad = C.c_int(0)
        for S in (1, 7, 100):
            for p, want in zip(g[f"{tname}__{S}__preds"], g[f"{tname}__{S}__est"]):
                p = np.ascontiguousarray(p, dtype=np.float64)
                assert lib.ape_selfcheck_row_pose(ti, p.ctypes.data_as(C.POINTER(C.c_double)),
                                                  body.ctypes.data_as(C.POINTER(C.c_double)), use_float, est, C.byref(bad)) == 0
                assert bad.value == 0
                got = np.array(est[:W])
                q0 = 6 if W == 14 else 9
                np.testing.assert_allclose(got[:q0], want[:q0], rtol=0, atol=tol)
                for k in range(q0, W, 4):                      # quaternions: equal up to sign (w ~ 0 can flip in float)
                    d = min(np.abs(got[k:k + 4] - want[k:k + 4]).max(), np.abs(got[k:k + 4] + want[k:k + 4]).max())
                    assert d <= tol
    # degenerate 6D pair is flagged (the reference raises LinAlgError)
    z = np.zeros(12)
    lib.ape_selfcheck_row_pose(0, z.ctypes.data_as(C.POINTER(C.c_double)), body.ctypes.data_as(C.POINTER(C.c_double)), 1, est, C.byref(bad))
    assert bad.value == 1
