"""CPU: the labeler oracle against the golden written from the reference's own methods (oracle/pin_labeler_reference.py)."""
import numpy as np
import torch


def test_match_and_filter_corr_oracle_matches_golden(golden_dir):
    from oracle import labeler_oracle as LO
    from oracle.pin_labeler_reference import labeler_inputs
    g = np.load(f'{golden_dir}/labeler_2pairs.npz')
    C0, F0, C1, F1 = labeler_inputs()
    _, unc = LO.match_and_filter_corr(C0, F0, C1, F1, radius=float(g['radius']), feature_filter='Lowe', spatial_filter='Spherical')
    for i, u in enumerate(unc):
        assert np.array_equal(u.numpy(), g[f'unc_lowe_sph_{i}'])


def test_knn_points_oracle_semantics():
    from oracle import labeler_oracle as LO
    g = torch.Generator().manual_seed(0)
    p1, p2 = torch.randn(2, 40, 5, generator=g), torch.randn(2, 60, 5, generator=g)
    p2[0, 7] = p2[0, 3]
    p1[0, 0] = p2[0, 3]
    out = LO.knn_points(p1, p2, torch.tensor([40, 25]), torch.tensor([60, 33]), K=2)
    d = ((p1[0, :, None] - p2[0, None]) ** 2).sum(-1)
    assert out.idx[0, 0].tolist() == [3, 7] and float(out.dists[0, 0, 0]) == 0.0          # duplicate: lower index first
    assert torch.equal(out.idx[0, :, 0], d.argmin(1))
    assert bool((out.dists[..., 0] <= out.dists[..., 1]).all())
    assert float(out.dists[1, 25:].abs().max()) == 0 and int(out.idx[1, 25:].abs().max()) == 0
    assert int(out.idx[1, :25].max()) < 33


def test_c_helper_equals_numpy_restatement():
    """oracle/csrc/knn_oracle.c (gcc, fmaf) == the numpy kernel-order restatement, indices and distance bits."""
    from oracle import labeler_oracle as LO
    from oracle.build_c import load
    if load() is None:
        import pytest
        pytest.skip('gcc unavailable')
    rng = np.random.default_rng(1)
    for D, K in ((32, 2), (3, 1), (5, 2)):
        a = rng.normal(size=(700, D)).astype(np.float32)
        b = rng.normal(size=(900, D)).astype(np.float32)
        b[10] = b[400]
        a[0] = b[400]
        i0, d0 = LO._knn(a, b, K)
        i1, d1 = LO._knn(a, b, K, force_numpy=True)
        assert np.array_equal(i0, i1) and np.array_equal(d0.view(np.int32), d1.view(np.int32))
