"""GPU: the fused first convolution (eyoc_stem_conv: block-occupancy pre-filter + hash probes + 1-channel convolution +
BN / ReLU + split-half packing, csrc/coordmap.cu) against the table path it replaces (eyoc_kernel_map_self with all ksize^3
offsets -> eyoc_sparse_conv -> eyoc_xh_pack): bit-identical rows and an identical 3^3 neighbour table."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _coords(kind, rng):
    from eyoc_b200 import synth
    if kind == 'lidar':
        pair = synth.make_pair(3, distance=10.0, az_step_deg=1.0)
        return synth.collate([pair['coords0'], pair['coords1']])
    if kind == 'dense':                    # a filled cube: every neighbour exists
        g = np.stack(np.meshgrid(np.arange(-7, 9), np.arange(-5, 6), np.arange(-3, 10), indexing='ij'), -1).reshape(-1, 3)
        return np.concatenate([np.zeros((len(g), 1), np.int64), g], 1).astype(np.int32)
    if kind == 'edges':                    # voxels against the ends of the packed 16-bit range, several clouds, negatives
        parts = []
        for b, base in enumerate(([32760, 32760, 32760], [-32768, -32768, -32768], [-3, 32764, -32766], [0, 0, 0])):
            g = rng.integers(0, 8, (300, 3)) + np.array(base)
            g = np.unique(np.clip(g, -32768, 32767), axis=0)
            parts.append(np.concatenate([np.full((len(g), 1), b), g], 1))
        c = np.concatenate(parts).astype(np.int32)
        return c[rng.permutation(len(c))]
    c = rng.integers(-40, 40, (20000, 3))  # 'random': ~4 % occupancy, shuffled rows, 3 clouds
    c = np.concatenate([rng.integers(0, 3, (len(c), 1)), c], 1)
    c = np.unique(c, axis=0).astype(np.int32)
    return c[rng.permutation(len(c))]


@pytest.mark.parametrize('kind', ['lidar', 'dense', 'edges', 'random'])
@pytest.mark.parametrize('ksize', [5, 3])
@pytest.mark.parametrize('mode', ['f16x3', 'tf32x3'])
def test_stem_conv_matches_table_path(kind, ksize, mode):
    from eyoc_b200 import nn as enn
    from eyoc_b200.sparse import SparseTensor
    rng = np.random.default_rng(ksize * 7 + len(kind))
    coords = torch.from_numpy(_coords(kind, rng)).cuda()
    n = len(coords)
    feats = torch.from_numpy(rng.normal(size=(n, 1)).astype(np.float32)).cuda()
    conv = enn.MinkowskiConvolution(1, 32, kernel_size=ksize, stride=1, dimension=3).cuda()
    norm = enn.MinkowskiBatchNorm(32).cuda().eval()
    with torch.no_grad():
        conv.kernel.copy_(torch.from_numpy(rng.normal(size=tuple(conv.kernel.shape)).astype(np.float32)))
        norm.bn.weight.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, 32).astype(np.float32)))
        norm.bn.bias.copy_(torch.from_numpy(rng.normal(size=32).astype(np.float32)))
        norm.bn.running_mean.copy_(torch.from_numpy(rng.normal(size=32).astype(np.float32)))
        norm.bn.running_var.copy_(torch.from_numpy(rng.uniform(0.5, 2.0, 32).astype(np.float32)))
    res = {}
    old_mode, old_fused = enn.CONV_MODE, enn.STEM_FUSED
    try:
        enn.CONV_MODE = mode
        for fused in (False, True):
            enn.STEM_FUSED = fused
            for relu in (False, True):
                x = SparseTensor(feats, coordinates=coords)
                with torch.no_grad():
                    y = enn.conv_bn_act(x, conv, norm, relu=relu)
                mgr = x.coordinate_manager
                res[fused, relu] = (y.Fh.clone() if mode == 'f16x3' else y.F.clone(), mgr.kernel_map(1, 1, 3).clone())
                mgr._check_status()
                if fused and ksize == 5:
                    assert (1, 1, 5, False) not in mgr._maps          # no 125-column table was built
    finally:
        enn.CONV_MODE, enn.STEM_FUSED = old_mode, old_fused
    for relu in (False, True):
        a, b = res[False, relu], res[True, relu]
        assert a[0].dtype == b[0].dtype and torch.equal(a[0], b[0])
        assert torch.equal(a[1], b[1])
    assert float(res[True, False][0].float().abs().max()) > 0


@pytest.mark.parametrize('kind', ['lidar', 'edges'])
@pytest.mark.parametrize('ksize', [5, 3])
def test_stem_conv_all_ones_input(kind, ksize):
    """sparse.ones_features (the reference's occupancy-only input, lib/data_loaders.py:971-972) lets the kernel skip the gather
    of the input value and every probe outside the 3^3 sub-cube: same rows, same neighbour table as a plain ones tensor."""
    from eyoc_b200 import nn as enn
    from eyoc_b200.sparse import SparseTensor, ones_features
    rng = np.random.default_rng(3)
    coords = torch.from_numpy(_coords(kind, rng)).cuda()
    conv = enn.MinkowskiConvolution(1, 32, kernel_size=ksize, stride=1, dimension=3).cuda()
    with torch.no_grad():
        conv.kernel.copy_(torch.from_numpy(rng.normal(size=tuple(conv.kernel.shape)).astype(np.float32)))
    res = []
    for feats in (torch.ones((len(coords), 1), device='cuda'), ones_features(len(coords), 'cuda')):
        x = SparseTensor(feats, coordinates=coords)
        assert x.all_ones == (len(res) == 1)
        with torch.no_grad():
            y = enn.conv_bn_act(x, conv, None, relu=True)
        res.append((y.Fh.clone(), x.coordinate_manager.kernel_map(1, 1, 3).clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
