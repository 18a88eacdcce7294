"""The one-CTA Jacobi SVD (t10_svd_small) behind Svd / Svd_truncate for matrices that fit one SM, against LAPACK
(numpy): singular values to 1e-13, reconstruction and orthonormality to 1e-13, tall / wide / odd / rank-deficient
shapes, fp32 at 1e-5; the library wrappers return the reference's (U, S, V) triple (linalg.py:544-678)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu
DEV = os.environ.get("T10_TEST_DEVICE", "cuda")


@pytest.fixture(scope="module")
def T():
    import tor10_b200
    return tor10_b200


@pytest.mark.parametrize("shape", [(64, 64), (100, 37), (20, 50), (1, 1), (64, 3), (5, 64), (33, 33), (300, 64)])
def test_svd_small_vs_lapack(T, shape):
    from tor10_b200 import linalg
    rng = np.random.Generator(np.random.PCG64(shape[0] * 1000 + shape[1]))
    a = rng.standard_normal(shape)
    x = torch.from_numpy(a).to(DEV)
    got = linalg._svd_small(x)
    assert got is not None, "matrix should fit the one-CTA kernel"
    u, s, vh = (t.cpu().numpy() for t in got)
    s_ref = np.linalg.svd(a, compute_uv=False)
    k = min(shape)
    assert u.shape == (shape[0], k) and s.shape == (k,) and vh.shape == (k, shape[1])
    assert np.all(np.diff(s) <= 0), "singular values must come out descending"
    assert np.abs(s - s_ref).max() <= 1e-13 * s_ref.max()
    assert np.abs((u * s) @ vh - a).max() <= 1e-13 * np.abs(a).max() * max(shape)
    assert np.abs(u.T @ u - np.eye(k)).max() <= 1e-13 * max(shape)
    assert np.abs(vh @ vh.T - np.eye(k)).max() <= 1e-13 * max(shape)


def test_svd_small_rank_deficient_and_strided(T):
    from tor10_b200 import linalg
    rng = np.random.Generator(np.random.PCG64(5))
    b = rng.standard_normal((48, 10))
    a = b @ rng.standard_normal((10, 40))                 # rank 10
    x = torch.from_numpy(a).to(DEV)
    u, s, vh = (t.cpu().numpy() for t in linalg._svd_small(x))
    s_ref = np.linalg.svd(a, compute_uv=False)
    assert np.abs(s - s_ref).max() <= 1e-12 * s_ref.max()
    assert np.abs((u * s) @ vh - a).max() <= 1e-12 * np.abs(a).max() * 48
    # a non-contiguous view (transposed, then every second row)
    big = torch.from_numpy(rng.standard_normal((80, 60))).to(DEV)
    view = big.t()[::2]                                    # 30 x 80, strides (2, 60)
    u, s, vh = (t.cpu().numpy() for t in linalg._svd_small(view))
    ref = view.cpu().numpy()
    assert np.abs((u * s) @ vh - ref).max() <= 1e-13 * np.abs(ref).max() * 80
    # fp32
    x32 = torch.from_numpy(rng.standard_normal((64, 64)).astype(np.float32)).to(DEV)
    u, s, vh = (t.cpu().numpy().astype(np.float64) for t in linalg._svd_small(x32))
    ref = x32.cpu().numpy().astype(np.float64)
    assert np.abs((u * s) @ vh - ref).max() <= 1e-5 * np.abs(ref).max() * 8
    # too large for one SM: the caller falls back to torch / cuSOLVER
    assert linalg._svd_small(torch.zeros((200, 100), device=DEV, dtype=torch.float64)) is None


def test_svd_truncate_wrapper(T):
    chi = 32
    g = torch.Generator().manual_seed(11)
    m = torch.rand((2 * chi, 2 * chi), generator=g, dtype=torch.float64)
    x = T.UniTensor(bonds=[T.Bond(2 * chi), T.Bond(2 * chi)], rowrank=1, device=torch.device(DEV))
    x.Storage = m.to(DEV)
    u, s, v = T.Svd_truncate(x, chi)
    assert tuple(u.shape) == (2 * chi, chi) and tuple(v.shape) == (chi, 2 * chi) and s.is_diag
    ur, sr, vr = np.linalg.svd(m.numpy(), full_matrices=False)
    assert np.abs(s.Storage.cpu().numpy() - sr[:chi]).max() <= 1e-13 * sr[0]
    rec = (u.Storage * s.Storage) @ v.Storage
    assert np.abs(rec.cpu().numpy() - (ur[:, :chi] * sr[:chi]) @ vr[:chi]).max() <= 1e-12 * sr[0]


def test_svd_small_warm_start_sequence(T):
    """Consecutive calls on one shape start from the previous right singular vectors (warm start, restarted from the
    identity every 64 calls): every result of a slowly drifting sequence -- and of an unrelated matrix thrown in --
    is a valid SVD to the same tolerances as a cold start."""
    from tor10_b200 import linalg
    rng = np.random.Generator(np.random.PCG64(17))
    a = rng.standard_normal((64, 64))
    for it in range(70):
        a = a + 1e-3 * rng.standard_normal((64, 64))
        if it == 40:
            a = rng.standard_normal((64, 64))           # nothing to do with the previous matrix
        x = torch.from_numpy(a).to(DEV)
        u, s, vh = (t.cpu().numpy() for t in linalg._svd_small(x))
        s_ref = np.linalg.svd(a, compute_uv=False)
        assert np.abs(s - s_ref).max() <= 1e-13 * s_ref.max(), it
        assert np.abs((u * s) @ vh - a).max() <= 1e-13 * np.abs(a).max() * 64, it
        assert np.abs(vh @ vh.T - np.eye(64)).max() <= 1e-12, it
        assert np.abs(u.T @ u - np.eye(64)).max() <= 1e-12, it
