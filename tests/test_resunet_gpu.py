"""GPU parity: coordinate/kernel maps and the fused ResUNetBN2C forward vs the ME-semantics oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cloud(n, seed, extent=24, batch=1):
    rng = np.random.default_rng(seed)
    out = []
    for b in range(batch):
        # a thin, surface-like blob so that neighbourhood occupancy resembles LiDAR voxels
        xyz = np.stack([rng.integers(-extent, extent, 4 * n), rng.integers(-extent, extent, 4 * n),
                        rng.integers(-3, 4, 4 * n)], 1)
        xyz = np.unique(xyz, axis=0)
        xyz = xyz[rng.permutation(len(xyz))[:n]]
        out.append(np.concatenate([np.full((len(xyz), 1), b), xyz], 1))
    return np.concatenate(out, 0).astype(np.int32)


def test_levels_and_kernel_maps_bit_exact():
    from eyoc_b200.sparse import CoordinateManager
    from oracle import resunet_oracle as RO
    coords = _cloud(2500, 0, batch=2)
    mgr = CoordinateManager(torch.from_numpy(coords).cuda())
    maps = RO.build_maps(coords, 5)
    mgr.ensure_levels(8)
    for ts, m in zip((1, 2, 4, 8), maps['levels']):
        np.testing.assert_array_equal(mgr.levels[ts].coords.cpu().numpy(), m.coords.astype(np.int32))
    np.testing.assert_array_equal(mgr.kernel_map(1, 1, 5).cpu().numpy(), maps['k5'])
    for lvl, ts in enumerate((1, 2, 4, 8)):
        np.testing.assert_array_equal(mgr.kernel_map(ts, ts, 3).cpu().numpy(), maps['s1'][lvl])
    for lvl, ts in enumerate((1, 2, 4)):
        np.testing.assert_array_equal(mgr.kernel_map(ts, 2 * ts, 3).cpu().numpy(), maps['down'][lvl])
        np.testing.assert_array_equal(mgr.kernel_map(2 * ts, ts, 3, transposed=True).cpu().numpy(), maps['up'][lvl])
        perm = mgr.parity_perm(ts).cpu().numpy()
        assert sorted(perm.tolist()) == list(range(mgr.num_rows(ts)))


def test_negative_coordinates_floor():
    from eyoc_b200.sparse import CoordinateManager
    c = np.array([[0, -1, -2, -3], [0, -4, 1, 0], [0, 3, -1, 2], [0, -5, -5, -5]], np.int32)
    mgr = CoordinateManager(torch.from_numpy(c).cuda())
    mgr.ensure_levels(4)
    got2 = mgr.levels[2].coords.cpu().numpy()
    want2 = c.copy()
    want2[:, 1:] = np.floor_divide(c[:, 1:], 2) * 2
    np.testing.assert_array_equal(got2, want2)           # all distinct here, first-occurrence order
    got4 = mgr.levels[4].coords.cpu().numpy()
    want4 = want2.copy()
    want4[:, 1:] = np.floor_divide(want2[:, 1:], 4) * 4
    _, first = np.unique(want4, axis=0, return_index=True)
    np.testing.assert_array_equal(got4, want4[np.sort(first)])


def test_duplicate_and_range_errors():
    from eyoc_b200.sparse import CoordinateManager, SparseTensor
    c = torch.tensor([[0, 1, 2, 3], [0, 1, 2, 3]], dtype=torch.int32).cuda()
    with pytest.raises(RuntimeError):
        CoordinateManager(c).ensure_levels(2)
    c = torch.tensor([[0, 1, 2, 40000]], dtype=torch.int32).cuda()
    with pytest.raises(RuntimeError):
        CoordinateManager(c).ensure_levels(2)
    with pytest.raises(RuntimeError):
        SparseTensor(torch.ones(3, 1), coordinates=torch.zeros(3, 4, dtype=torch.int32))      # CPU: no fallback


@pytest.mark.parametrize('n,batch,seed', [(1500, 1, 1), (3000, 2, 2), (200, 1, 3)])
def test_forward_vs_oracle(n, batch, seed):
    """Feature parity (fp32-accurate path): max |F - F_oracle| <= 1e-5, row order preserved."""
    from eyoc_b200.model import load_model
    from eyoc_b200.sparse import SparseTensor
    from oracle import resunet_oracle as RO
    coords = _cloud(n, seed, batch=batch)
    feats = torch.ones((len(coords), 1), dtype=torch.float32)
    sd = RO.make_state_dict(1, 32, 5, seed=seed)
    Model = load_model('ResUNetBN2C')
    model = Model(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3)
    missing = model.load_state_dict(sd, strict=True)          # MinkowskiEngine key names / shapes
    model = model.cuda().eval()
    sinput = SparseTensor(feats.cuda(), coordinates=torch.from_numpy(coords).cuda())
    out = model(sinput)
    assert len(out) == len(coords) and out.F.shape == (len(coords), 32)
    np.testing.assert_array_equal(out.C.cpu().numpy(), coords)
    want = RO.resunet_forward(coords, feats, sd, True, 5)
    got = out.F.cpu()
    assert float((got - want).abs().max()) <= 1e-5, float((got - want).abs().max())
    cos = (got * want).sum(1)
    assert float(cos.min()) >= 1 - 1e-6
    # un-normalised variant exercises the bias-only epilogue
    model.normalize_feature = False
    det = {}
    RO.resunet_forward(coords, feats, sd, False, 5, detail=det)
    got_raw = model(sinput).F.cpu()
    scale = float(det['pre_norm'].abs().max())
    assert float((got_raw - det['pre_norm']).abs().max()) <= 2e-5 * max(scale, 1.0)


def test_train_mode_rejected_and_module_api():
    from eyoc_b200.model import load_model
    from eyoc_b200.sparse import SparseTensor
    assert load_model('NoSuchNet') is None
    model = load_model('ResUNetBN2C')(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3).cuda()
    coords = torch.from_numpy(_cloud(100, 5)).cuda()
    x = SparseTensor(torch.ones(len(coords), 1).cuda(), coordinates=coords)
    with pytest.raises(NotImplementedError):
        model(x)                                             # training mode: inference path only
    keys = set(model.state_dict().keys())
    assert {'conv1.kernel', 'norm1.bn.running_var', 'block1.conv1.kernel', 'block4_tr.norm2.bn.weight',
            'conv1_tr.kernel', 'final.kernel', 'final.bias'} <= keys
    assert model.state_dict()['conv1.kernel'].shape == (125, 1, 32)
    assert model.state_dict()['conv1_tr.kernel'].shape == (96, 64)
    assert model.state_dict()['final.bias'].shape == (1, 32)


def test_voxelize_gpu_matches_reference_selection():
    """eyoc_voxelize == sparse_quantize(xyz / 0.3, return_index=True) + floor(xyz[sel] / 0.3) (lib/data_loaders.py:940-972),
    clouds batched through the per-point cloud index."""
    from eyoc_b200 import synth
    from eyoc_b200.sparse import voxelize_gpu
    rng = np.random.default_rng(3)
    clouds = [(rng.normal(size=(n, 3)) * np.array([30.0, 30.0, 2.0])).astype(np.float32) for n in (50000, 1, 33333)]
    clouds[0][:100] = clouds[0][100:200]                     # exact duplicates
    xyz = np.concatenate(clouds)
    cloud = np.concatenate([np.full(len(c), b, np.int32) for b, c in enumerate(clouds)])
    coords, sel = voxelize_gpu(torch.from_numpy(xyz).cuda(), 0.3, torch.from_numpy(cloud).cuda())
    want_c, want_s, off = [], [], 0
    for b, c in enumerate(clouds):
        pts, q = synth.voxelize(c, 0.3)
        key = {tuple(r): i for i, r in reversed(list(enumerate(np.floor(c / np.float32(0.3)).astype(np.int64).tolist())))}
        s = np.sort(np.fromiter(key.values(), dtype=np.int64))
        np.testing.assert_array_equal(c[s], pts)
        want_c.append(np.concatenate([np.full((len(q), 1), b, np.int32), q], 1))
        want_s.append(s + off)
        off += len(c)
    np.testing.assert_array_equal(coords.cpu().numpy(), np.concatenate(want_c))
    np.testing.assert_array_equal(sel.cpu().numpy(), np.concatenate(want_s))
