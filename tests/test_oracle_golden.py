"""CPU: the oracle restatements reproduce the committed golden vectors (which oracle/pin_against_reference.py
generated from the reference's own functions imported from /root/reference, bit-compared there)."""
import dataclasses
import json

import numpy as np
import pytest
import torch

from oracle import matching_oracle as MO
from oracle import sc2pcr_oracle as O


def _cfg(g):
    c = json.loads(str(g['cfg']))
    c.pop('stable_ties', None)
    return O.SC2Config(**c)


@pytest.mark.parametrize('name', ['sc2pcr_n25_s4', 'sc2pcr_n1000_s1', 'sc2pcr_n2000_s3'])
def test_sc2pcr_oracle_reproduces_reference_goldens(golden_dir, name):
    g = np.load(f'{golden_dir}/{name}.npz')
    cfg = _cfg(g)
    src, tgt = torch.from_numpy(g['src'])[None], torch.from_numpy(g['tgt'])[None]
    det = {}
    T, fit = O.sc2_pcr(src, tgt, cfg, det)
    np.testing.assert_array_equal(T[0].numpy(), g['final_trans'])
    np.testing.assert_array_equal(fit[0].numpy(), g['fitness'])
    np.testing.assert_array_equal(det['seeds'][0].numpy(), g['seeds'])
    np.testing.assert_array_equal(det['confidence'][0].numpy(), g['confidence'])
    assert det['global_iters'] == int(g['global_iters']) and det['local_iters'] == int(g['local_iters'])
    # stable-tie mode: what the CUDA path is compared with stage by stage
    ds = {}
    T_s, fit_s = O.sc2_pcr(src, tgt, dataclasses.replace(cfg, stable_ties=True), ds)
    np.testing.assert_array_equal(ds['seeds'][0].numpy(), g['st_seeds'])
    np.testing.assert_array_equal(ds['topk1'][0].numpy(), g['st_topk1'])
    np.testing.assert_array_equal(ds['topk2'][0].numpy(), g['st_topk2'])
    np.testing.assert_array_equal(T_s[0].numpy(), g['st_final_trans'])
    # the two tie rules agree on everything the north star asks for
    assert np.linalg.norm(g['st_final_trans'][:3, :3] - g['final_trans'][:3, :3]) < 1e-4
    assert np.linalg.norm(g['st_final_trans'][:3, 3] - g['final_trans'][:3, 3]) < 1e-3
    np.testing.assert_array_equal(g['st_labels'], g['labels'])


def test_recovers_planted_pose(golden_dir):
    g = np.load(f'{golden_dir}/sc2pcr_n8000_s5.npz')
    T, T_gt = g['final_trans'], g['T_gt']
    assert np.linalg.norm(T[:3, :3] - T_gt[:3, :3]) < 5e-3 and np.linalg.norm(T[:3, 3] - T_gt[:3, 3]) < 0.05
    assert g['labels'].sum() >= 0.9 * g['gt_inlier'].sum()


def test_knn_oracle_golden(golden_dir):
    g = np.load(f'{golden_dir}/knn_1500x1300.npz')
    F0, F1 = torch.from_numpy(g['F0']), torch.from_numpy(g['F1'])
    i, d = MO.find_nn(F0, F1, nn_max_n=500, return_distance=True)
    np.testing.assert_array_equal(i.numpy(), g['idx_sq'])
    np.testing.assert_array_equal(d.numpy(), g['dist_sq'])
    np.testing.assert_array_equal(MO.match_argmin(F0, F1).numpy(), g['idx_cos'])
    # kernel-order restatements agree with the torch-order ones on this (well separated) set
    np.testing.assert_array_equal(MO.knn_sq_seq(g['F0'][:300], g['F1'])[0], g['idx_sq'][:300])
    np.testing.assert_array_equal(MO.knn_cos_seq(g['F0'][:300], g['F1'])[0], g['idx_cos'][:300])


def test_kabsch_and_irls_golden(golden_dir):
    g = np.load(f'{golden_dir}/kabsch_5x20.npz')
    T = O.kabsch_weighted(torch.from_numpy(g['A']), torch.from_numpy(g['B']), torch.from_numpy(g['w']).clone())
    np.testing.assert_array_equal(T.numpy(), g['T'])
    g = np.load(f'{golden_dir}/irls_400.npz')
    np.testing.assert_array_equal(MO.irls_pose(torch.from_numpy(g['p0']), torch.from_numpy(g['p1'])).numpy(), g['T'])


def test_rng_draw_order_matches_pipeline_plan():
    """pipeline.draw_indices makes the reference's six draws in the reference's order."""
    from eyoc_b200.pipeline import draw_indices
    n0, n1 = 7000, 6500
    np.random.seed(5)
    d = draw_indices(n0, n1, 5000, 5000, 8000)
    np.random.seed(5)
    xyz0, xyz1 = np.arange(n0)[:, None].repeat(3, 1), np.arange(n1)[:, None].repeat(3, 1)
    F0, F1 = torch.zeros(n0, 4), torch.zeros(n1, 4)
    a, b = MO.find_corr(xyz0, xyz1, F0, F1, subsample_size=5000)            # draws choice, choice
    np.testing.assert_array_equal(a[:, 0], d['fc0'])
    s0, _ = MO.random_sample(xyz0, F0, 5000)
    s1, _ = MO.random_sample(xyz1, F1, 5000)
    np.testing.assert_array_equal(s0[:, 0], d['rs0'])
    np.testing.assert_array_equal(s1[:, 0], d['rs1'])
    det = {}
    cfg = O.SC2Config(num_node=8000)
    O.match_pair(torch.zeros(1, 5000, 3), torch.zeros(1, 5000, 3), torch.zeros(1, 5000, 4), torch.zeros(1, 5000, 4), cfg, det)
    np.testing.assert_array_equal(det['src_sel'], d['mp0'])
    np.testing.assert_array_equal(det['tgt_sel'], d['mp1'])
