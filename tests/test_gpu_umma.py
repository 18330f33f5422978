"""GPU (-m gpu): the tcgen05 / TMEM plumbing (descriptors, commit, tcgen05.ld lane mapping) against a float64 matmul."""
import numpy as np
import pytest
import torch

from arm_pose_estimation_b200 import _native as N

pytestmark = pytest.mark.gpu


def pack_k_major(m):
    """[rows, K] -> canonical K-major no-swizzle bytes: [K/8][rows][8] halfs (csrc/ape_umma.cuh)."""
    rows, K = m.shape
    return np.ascontiguousarray(m.astype(np.float16).reshape(rows, K // 8, 8).transpose(1, 0, 2))


@pytest.mark.parametrize("cta_group,Nn,K", [(1, 128, 64), (1, 256, 256), (1, 32, 16), (2, 128, 64), (2, 256, 256), (2, 64, 128),
                                            # + 16: A operand through tensor memory (the H = 256 LSTM kernel keeps h_t there)
                                            (17, 128, 64), (17, 128, 256), (18, 128, 64), (18, 128, 256), (18, 256, 256)])
def test_umma_gemm(cta_group, Nn, K):
    mode, cta_group = cta_group, cta_group & 3
    rng = np.random.default_rng(Nn + K + cta_group)
    M = 128 * cta_group
    a = rng.normal(size=(M, K)).astype(np.float16)
    b = rng.normal(size=(Nn, K)).astype(np.float16)
    ap = np.stack([pack_k_major(a[i * 128:(i + 1) * 128]) for i in range(cta_group)])
    nl = Nn // cta_group
    bp = np.stack([pack_k_major(b[i * nl:(i + 1) * nl]) for i in range(cta_group)])
    ad, bd = torch.from_numpy(ap).cuda(), torch.from_numpy(bp).cuda()
    d = torch.full((M, Nn), float("nan"), dtype=torch.float32, device="cuda")
    N.check(N.load().ape_selftest_umma(N.ptr(ad), N.ptr(bd), N.ptr(d), Nn, K, mode, N.current_stream_ptr()), "selftest")
    torch.cuda.synchronize()
    want = a.astype(np.float64) @ b.astype(np.float64).T
    got = d.cpu().numpy()
    err = np.abs(got - want).max()
    print(f"cta_group={cta_group} a_tmem={mode >> 4} N={Nn} K={K}: max |d| = {err:.3g}")
    assert err < 1e-3 * np.sqrt(K)
