"""The C ABI used from plain C (examples/cabi_consumer.c: gcc, the header, libape_b200.so and the CUDA runtime - no Python, no torch),
as a host written in the reference's would-be FFI language would use it.  CPU: the example compiles and links as C against the header
and the shipped library.  GPU: it runs the three stages on inputs it generates, and the same call replayed through the ctypes binding
(torch-owned device memory) gives bit-identical messages, std and samples."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from arm_pose_estimation_b200 import _native as N

SRC = ROOT / "examples" / "cabi_consumer.c"
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build(tmp_path):
    lib = N.lib_path() if hasattr(N, "lib_path") else ROOT / "arm_pose_estimation_b200" / "lib" / "libape_b200.so"
    exe = tmp_path / "cabi_consumer"
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", f"-I{CUDA}/include", str(SRC), "-o", str(exe),
           f"-L{os.path.dirname(str(lib))}", "-lape_b200", f"-L{CUDA}/lib64", "-lcudart", "-lm", f"-Wl,-rpath,{os.path.dirname(str(lib))}",
           f"-Wl,-rpath,{CUDA}/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no C compiler")
def test_c_consumer_compiles_and_links_as_c(tmp_path):
    N.load()                                           # (the library must have been built: make)
    assert build(tmp_path).exists()


@pytest.mark.gpu
def test_c_consumer_matches_the_ctypes_binding(tmp_path):
    import torch
    exe, dump = build(tmp_path), tmp_path / "out.bin"
    r = subprocess.run([str(exe), str(dump)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    buf = dump.read_bytes()
    B, nF, n, ncols, I, H, L, T, O, smooth, blob_floats, _ = np.frombuffer(buf, np.int32, 12)
    E, S, off = B * nF, smooth * n, 48

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(buf, dtype, count, off)
        off += a.nbytes
        return a
    raw, blob = take(np.float32, E * ncols), take(np.float32, blob_floats)
    xx_m, xx_s, yy_m, yy_s, body9 = take(np.float64, I), take(np.float64, I), take(np.float32, O), take(np.float32, O), take(np.float32, 9)
    msg_c, std_c, smp_c = take(np.float32, E * 25).reshape(E, 25), take(np.float32, E * 6).reshape(E, 6), take(np.float32, E * S * 6).reshape(E, S, 6)
    assert off == len(buf)

    dev = lambda a: torch.from_numpy(np.array(a)).cuda()         # (a writable copy of the read-only file view)
    lib, st = N.load(), N.current_stream_ptr()
    feat_ring, pred_ring = nF + T - 1, nF + smooth - 1
    d_raw, d_blob, d_xm, d_xs, d_ym, d_ys, d_body = dev(raw), dev(blob), dev(xx_m), dev(xx_s), dev(yy_m), dev(yy_s), dev(body9)
    feats = torch.zeros((B, feat_ring, I), device="cuda")
    preds = torch.zeros((B, pred_ring, n, O), device="cuda")
    msg, std, smp = torch.empty((E, 25), device="cuda"), torch.empty((E, 6), device="cuda"), torch.empty((E, S, 6), device="cuda")
    status = torch.empty(E, dtype=torch.int32, device="cuda")
    N.check(lib.ape_features(N.ptr(d_raw), N.LAYOUT_WATCH_ONLY, 0, N.ptr(d_xm), N.ptr(d_xs), 1, N.ptr(feats), B, nF, 0, None, feat_ring, st), "ape_features")
    import ctypes
    nbytes = ctypes.c_uint64(0)
    N.check(lib.ape_mc_lstm_workspace_bytes(I, H, L, T, O, E, n, ctypes.byref(nbytes)), "workspace")
    ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device="cuda")
    a = N.LstmArgs()
    a.weights, a.I, a.H, a.L, a.T, a.O, a.dropout_p = d_blob.data_ptr(), I, H, L, T, O, 0.2
    a.feat_ring_buf, a.feat_ring, a.B, a.nF, a.frame0, a.n_samples = feats.data_ptr(), feat_ring, B, nF, 0, n
    a.mask_mode, a.philox_seed, a.stream_id0 = N.MASK_PHILOX, 0x1234abcd, 7
    a.workspace = (ws.data_ptr() + 255) & ~255
    a.preds, a.pred_ring = preds.data_ptr(), pred_ring
    N.check(lib.ape_mc_lstm_fma(a, st), "ape_mc_lstm_fma")
    N.check(lib.ape_fk_reduce(N.ptr(preds), pred_ring, N.ptr(d_ym), N.ptr(d_ys), N.ptr(d_body), 0, O, B, nF, 0, None, n, smooth,
                              N.ptr(msg), N.ptr(smp), N.ptr(std), None, N.ptr(status), st), "ape_fk_reduce")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(msg.cpu().numpy(), msg_c)
    np.testing.assert_array_equal(std.cpu().numpy(), std_c)
    np.testing.assert_array_equal(smp.cpu().numpy(), smp_c)
    assert int(status.abs().sum()) == 0
    # forward-kinematics invariants of every MC row: bone lengths (bone_map.py:42-45 defaults used by the example)
    hand, elbow = smp_c[..., 0:3].astype(np.float64), smp_c[..., 3:6].astype(np.float64)
    np.testing.assert_allclose(np.linalg.norm(hand - elbow, axis=-1), 0.22, atol=1e-5)
    np.testing.assert_allclose(np.linalg.norm(elbow - body9[6:9].astype(np.float64), axis=-1), 0.26, atol=1e-5)
