"""GPU: the drop-in inference driver (eyoc_b200/scripts/test_kitti.py) - the reference's per-pair loop and the batched
pipeline give the same poses bit for bit on the same seeded RNG stream, and raw sweeps voxelised on the device feed it."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_per_pair_loop_equals_blocks():
    from eyoc_b200.scripts import test_kitti as tk
    cfg = tk.make_config(tk.parse_args(['--use_RANSAC', 'false']))
    loader = tk.SyntheticPairLoader(range(3))
    np.random.seed(5)
    a = tk.main(cfg, loader)
    np.random.seed(5)
    b = tk.main_blocks(cfg, loader, block=2)                 # 2 + 1 pairs: also a ragged last block
    assert a['count'] == b['count'] == 3
    for Ta, Tb in zip(a['T_est'], b['T_est']):
        assert torch.equal(Ta, Tb)
    assert a['success'] == b['success'] and a['success'] >= 2
    assert len(a['dists_nn']) == 3 and len(a['dists_nn'][0]) == 5000


def test_raw_pair_loader_matches_host_voxelisation():
    """RawPairLoader (eyoc_voxelize) yields the coordinates / point selection the reference's loader computes on the host."""
    from eyoc_b200 import synth
    from eyoc_b200.scripts import test_kitti as tk
    rng = np.random.default_rng(0)
    xyz0 = (rng.normal(size=(20000, 3)) * np.array([20.0, 20.0, 1.5])).astype(np.float32)
    xyz1 = (rng.normal(size=(15000, 3)) * np.array([20.0, 20.0, 1.5])).astype(np.float32)
    d = next(iter(tk.RawPairLoader([(xyz0, xyz1, np.eye(4, dtype=np.float32))])))
    for side, xyz in ((0, xyz0), (1, xyz1)):
        pts, q = synth.voxelize(xyz, 0.3)
        np.testing.assert_array_equal(d[f'sinput{side}_C'][:, 1:].cpu().numpy(), q)
        np.testing.assert_array_equal(d[f'pcd{side}'][0].numpy(), pts)
        assert int(d[f'sinput{side}_C'][:, 0].abs().sum()) == 0 and d[f'sinput{side}_F'].shape == (len(q), 1)
