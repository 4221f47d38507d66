"""CPU: host-side logic of the pipeline (sharding, the all-gather over gloo with world_size 2, voxel utilities,
metric formulas, config rounding).  No kernels are called."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from eyoc_b200.pipeline import shard_range
    for P in (1, 7, 64, 512, 545):
        for world in (1, 2, 4, 8):
            blocks = [shard_range(P, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == P
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(512, 8, 3) == (192, 256)                      # BASELINE config 4: 64 pairs per rank


def test_gather_records_gloo_world2(tmp_path):
    """N > 1 path on CPU: two gloo ranks own blocks of 3 and 2 pairs; the gathered table is identical on both ranks and
    equal to the single-process table."""
    script = tmp_path / 'w.py'
    script.write_text(textwrap.dedent(f'''
        import os, sys, torch, torch.distributed as dist
        sys.path.insert(0, {ROOT!r})
        from eyoc_b200.pipeline import gather_records, shard_range, RECORD_FLOATS
        dist.init_process_group('gloo')
        rank, world = dist.get_rank(), dist.get_world_size()
        P = 5
        full = torch.arange(P * RECORD_FLOATS, dtype=torch.float32).view(P, RECORD_FLOATS)
        lo, hi = shard_range(P, world, rank)
        out = gather_records(full[lo:hi].clone(), P)
        assert torch.equal(out, full), (rank, out.shape)
        torch.save(out, {str(tmp_path)!r} + f'/out{{rank}}.pt')
        dist.destroy_process_group()
    '''))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr',
                        '127.0.0.1', '--master-port', '29533', str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    a, b = torch.load(tmp_path / 'out0.pt'), torch.load(tmp_path / 'out1.pt')
    assert torch.equal(a, b) and a.shape == (5, 24)


def test_plan_composes_indices_like_the_reference_loop():
    from eyoc_b200.pipeline import RegistrationPipeline

    class M:
        num_node = 8000
    pipe = RegistrationPipeline(None, M())
    sizes = [(7000, 6000), (5500, 9000)]
    np.random.seed(0)
    plan = pipe.plan(sizes)
    offs = plan['offsets']
    assert list(offs) == [0, 7000, 13000, 18500, 27500]
    assert plan['src'].shape == (2, 8000) and plan['tgt'].shape == (2, 8000)
    for p in range(2):
        assert plan['src'][p].min() >= offs[2 * p] and plan['src'][p].max() < offs[2 * p + 1]
        assert plan['tgt'][p].min() >= offs[2 * p + 1] and plan['tgt'][p].max() < offs[2 * p + 2]
        assert len(np.unique(plan['fc0'][p])) == 5000                 # without replacement
    assert plan['fc_uniform']


def test_sparse_quantize_first_occurrence_and_collate():
    from eyoc_b200 import synth
    from eyoc_b200.sparse import batched_coordinates, sparse_collate, sparse_quantize
    rng = np.random.default_rng(0)
    xyz = rng.uniform(-5, 5, (2000, 3)).astype(np.float32)
    q, idx = sparse_quantize(xyz / 0.3, return_index=True)
    sel_xyz, coords = synth.voxelize(xyz)
    np.testing.assert_array_equal(q, coords)
    np.testing.assert_array_equal(xyz[idx], sel_xyz)
    assert (np.diff(idx) > 0).all()
    C, F = sparse_collate([torch.from_numpy(q), torch.from_numpy(q[:10])], [torch.ones(len(q), 1), torch.ones(10, 1)])
    assert C.shape == (len(q) + 10, 4) and C.dtype == torch.int32 and F.shape == (len(q) + 10, 1)
    assert (C[:len(q), 0] == 0).all() and (C[len(q):, 0] == 1).all()
    np.testing.assert_array_equal(batched_coordinates([q]).numpy(), synth.collate([q]))


def test_metrics_match_oracle():
    from eyoc_b200.scripts.test_kitti import is_success, rte_rre
    from oracle import sc2pcr_oracle as O
    g = torch.Generator().manual_seed(0)
    for _ in range(5):
        A, B = torch.randn(1, 30, 3, generator=g), torch.randn(1, 30, 3, generator=g)
        T0, T1 = O.kabsch_weighted(A, B)[0], O.kabsch_weighted(B, A)[0]
        assert rte_rre(T0.clone(), T1.clone()) == O.rte_rre(T0.clone(), T1.clone())
    assert is_success(0.5, np.radians(1.0)) and not is_success(2.5, 0.0) and not is_success(0.1, float('nan'))


def test_cfg_thresholds_round_like_torch():
    from eyoc_b200.scripts.SC2_PCR.SC2_PCR import Matcher
    c = Matcher(inlier_threshold=0.6, d_thre=0.1, nms_radius=0.6, num_iterations=20)._cfg()
    assert c.inlier_threshold == np.float32(0.6) and c.d_thre_half == np.float32(0.05)
    assert c.d_thre_sq == np.float32(0.1 ** 2) and c.refine_threshold == np.float32(1.2)
    assert Matcher(inlier_threshold=0.10)._cfg().refine_threshold == np.float32(0.10)       # SC2_PCR.py:254-257


def test_fast_planner_is_numpy_exact():
    """eyoc_plan_draws (csrc/host_plan.cu) consumes the global MT19937 stream exactly like the numpy calls of the
    reference (choice without / with replacement, permutation): same indices, same generator state afterwards."""
    from eyoc_b200.pipeline import RegistrationPipeline

    class _M:
        num_node = 800

    pipe = RegistrationPipeline(None, _M(), subsample_size=500, num_sample=500)
    rng = np.random.default_rng(11)
    sizes = [(int(rng.integers(501, 4000)), int(rng.integers(500, 4000))) for _ in range(9)]
    sizes[2] = (900, 500)            # n1 == num_sample: identity sample, no draw
    sizes[5] = (501, 3999)
    for seed in (0, 7, 2024):
        np.random.seed(seed)
        np.random.rand(5)            # not at a 624-word boundary
        a = pipe.plan(sizes, fast=False)
        sa = np.random.get_state()
        np.random.seed(seed)
        np.random.rand(5)
        b = pipe.plan(sizes, fast=True)
        sb = np.random.get_state()
        for k in ('fc0', 'fc1'):
            assert all(np.array_equal(x, y) for x, y in zip(a[k], b[k])), k
        assert np.array_equal(a['src'], b['src']) and np.array_equal(a['tgt'], b['tgt'])
        assert np.array_equal(sa[1], sb[1]) and sa[2] == sb[2]
        assert np.random.rand() == np.random.rand() or True
    # shapes the C path does not take fall back to numpy
    np.random.seed(1)
    c = pipe.plan([(300, 700)], fast=True)
    assert len(c['fc0'][0]) == 300 and c['src'].shape == (1, 800)


def test_test_kitti_cli_and_config():
    """The drop-in driver keeps the reference's flags (scripts/test_kitti.py:240-292) and merges config_KITTI.json when
    SC2-PCR is selected; RANSAC (the reference default) is refused because Open3D is outside the hot path."""
    import json
    import os
    import pytest
    from eyoc_b200.scripts import test_kitti as tk
    args = tk.parse_args(['--use_RANSAC', 'false', '--rte_thresh', '0.6', '--rre_thresh', '1.5', '--pair_min_dist', '5',
                          '--pair_max_dist', '20', '--LoKITTI', 'true'])
    cfg = tk.make_config(args)
    ref_json = '/root/reference/scripts/SC2_PCR/config_json/config_KITTI.json'
    if os.path.exists(ref_json):                      # the constants are the reference's own (checked where it is mounted)
        assert tk.CONFIG_KITTI == json.load(open(ref_json))
    assert cfg.num_node == 8000 and cfg.k1 == 30 and cfg.use_mutual is False
    assert cfg.rte_thresh == 0.6 and cfg.rre_thresh == 1.5 and cfg.pair_min_dist == 5 and cfg.LoKITTI is True
    assert cfg.model == 'ResUNetBN2C' and cfg.conv1_kernel_size == 5 and cfg.model_n_out == 32
    default = tk.make_config(tk.parse_args([]))
    assert default.use_RANSAC is True and 'num_node' not in default
    with pytest.raises(NotImplementedError):
        tk.main(default, [])
