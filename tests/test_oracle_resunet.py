"""CPU: pins the MinkowskiEngine-semantics restatement (oracle/resunet_oracle.py) against dense
F.conv3d / F.conv_transpose3d on a densified grid.  ME itself is absent (parity unpinned against ME)."""
import numpy as np
import pytest
import torch

from oracle import resunet_oracle as RO


def _coords(n, seed, extent=10):
    rng = np.random.default_rng(seed)
    xyz = np.unique(rng.integers(-extent, extent, (4 * n, 3)), axis=0)
    xyz = xyz[rng.permutation(len(xyz))[:n]]
    return np.concatenate([np.zeros((len(xyz), 1), np.int64), xyz], 1)


@pytest.mark.parametrize('ksize', [3, 5])
def test_stride1_conv_matches_dense(ksize):
    c = _coords(400, ksize)
    m = RO.CoordMap(c, 1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(len(c), 4, generator=g, dtype=torch.float64)
    W = torch.randn(ksize ** 3, 4, 6, generator=g, dtype=torch.float64)
    got = RO.sparse_conv(x, W, RO.kernel_map(m, m, ksize))
    want = RO.dense_conv_reference(c, x, c, W, ksize, 1, 1, False)
    assert float((got - want).abs().max()) < 1e-10


def test_stride2_and_transposed_match_dense():
    c = _coords(500, 7, extent=9)
    m1 = RO.CoordMap(c, 1)
    m2 = m1.downsample()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(len(c), 3, generator=g, dtype=torch.float64)
    W = torch.randn(27, 3, 5, generator=g, dtype=torch.float64)
    down = RO.sparse_conv(x, W, RO.kernel_map(m1, m2, 3))
    want = RO.dense_conv_reference(c, x, m2.coords, W, 3, 1, 2, False)
    assert float((down - want).abs().max()) < 1e-10
    Wt = torch.randn(27, 5, 2, generator=g, dtype=torch.float64)
    up = RO.sparse_conv(down, Wt, RO.kernel_map(m2, m1, 3, transposed=True))
    want = RO.dense_conv_reference(m2.coords, down, c, Wt, 3, 2, 2, True)
    assert float((up - want).abs().max()) < 1e-10


def test_downsample_floor_and_first_occurrence():
    c = np.array([[0, -1, -1, -1], [0, 0, 0, 0], [0, 1, 1, 1], [0, -2, -2, -2], [0, 2, 3, -3]], np.int64)
    m2 = RO.CoordMap(c, 1).downsample()
    np.testing.assert_array_equal(m2.coords, [[0, -2, -2, -2], [0, 0, 0, 0], [0, 2, 2, -4]])


def test_forward_shapes_and_unit_norm():
    c = _coords(300, 3, extent=8)
    sd = RO.make_state_dict(1, 32, 5, seed=0)
    F = RO.resunet_forward(c, torch.ones(len(c), 1), sd, True, 5)
    assert F.shape == (len(c), 32)
    assert float((F.norm(dim=1) - 1).abs().max()) < 1e-5
    assert sd['conv1.kernel'].shape == (125, 1, 32) and sd['conv1_tr.kernel'].shape == (96, 64)


def test_instance_norm_statement_matches_torch_instance_norm():
    """oracle._in (the restated ME.MinkowskiInstanceNorm) == torch.nn.functional.instance_norm per cloud (biased variance,
    eps inside the root), then the [1, C] affine."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    from oracle import resunet_oracle as RO
    g = torch.Generator().manual_seed(0)
    x = torch.randn(300, 16, generator=g) * 3 + 1
    batch = np.repeat([0, 1, 2], [120, 100, 80])
    perm = torch.randperm(300, generator=g)
    x, batch = x[perm], batch[perm.numpy()]
    sd = {'n.weight': torch.rand(1, 16, generator=g) + 0.5, 'n.bias': torch.randn(1, 16, generator=g)}
    got = RO._in(x, sd, 'n', batch)
    want = torch.empty_like(x)
    for b in range(3):
        sel = torch.from_numpy(batch == b)
        want[sel] = F.instance_norm(x[sel].T[None], eps=RO.IN_EPS)[0].T
    want = want * sd['n.weight'] + sd['n.bias']
    assert float((got - want).abs().max()) < 1e-5
