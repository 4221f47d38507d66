"""GPU parity: fused SC2-PCR estimator (csrc/sc2pcr.cu) through the C-ABI vs the pinned oracle goldens.

Goldens come from oracle/pin_against_reference.py: un-prefixed keys are the reference's own CPU output
(bit-identical oracle mode), ``st_*`` keys are the oracle in stable-tie mode (the tie rule the CUDA path
implements).  Integer stages are compared bit-exactly with the oracle's upstream tensors fed through the
C-ABI hooks; floating-point stages use the north-star tolerances (R 1e-4 Frobenius, t 1e-3 m).
"""
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = ['sc2pcr_n1000_s1', 'sc2pcr_n2000_s2', 'sc2pcr_n2000_s3', 'sc2pcr_n25_s4', 'sc2pcr_n8000_s5']


def _matcher(cfg):
    from eyoc_b200.scripts.SC2_PCR.SC2_PCR import Matcher
    return Matcher(inlier_threshold=cfg['inlier_threshold'], num_node=cfg['num_node'], use_mutual=False,
                   d_thre=cfg['d_thre'], num_iterations=cfg['num_iterations'], ratio=cfg['ratio'],
                   nms_radius=cfg['nms_radius'], max_points=cfg['max_points'], k1=cfg['k1'], k2=cfg['k2'])


def _load(golden_dir, name):
    g = np.load(f'{golden_dir}/{name}.npz')
    cfg = json.loads(str(g['cfg']))
    src = torch.from_numpy(g['src'])[None].cuda()
    tgt = torch.from_numpy(g['tgt'])[None].cuda()
    return g, cfg, src, tgt


def _pose_close(T, T_ref):
    T, T_ref = np.asarray(T, np.float64), np.asarray(T_ref, np.float64)
    assert np.linalg.norm(T[:3, :3] - T_ref[:3, :3]) < 1e-4, np.linalg.norm(T[:3, :3] - T_ref[:3, :3])
    assert np.linalg.norm(T[:3, 3] - T_ref[:3, 3]) < 1e-3, np.linalg.norm(T[:3, 3] - T_ref[:3, 3])


@pytest.mark.parametrize('name', CASES)
def test_end_to_end_vs_reference(golden_dir, name):
    """No hooks: the whole estimator vs the REFERENCE's own output (pose tolerance, inlier mask exact)."""
    g, cfg, src, tgt = _load(golden_dir, name)
    m = _matcher(cfg)
    det = {}
    T, fit, labels = m._run(src, tgt, want_labels=True, detail=det)
    _pose_close(T[0].cpu().numpy(), g['final_trans'])
    assert int((labels[0].cpu().numpy() != g['labels']).sum()) == 0
    assert int(det['global_iters'][0]) == int(g['global_iters'])
    np.testing.assert_allclose(det['confidence'][0].cpu().numpy(), g['confidence'], rtol=2e-4, atol=1e-7)
    T2, fit2 = m.SC2_PCR(src, tgt)
    assert torch.equal(T2, T) and torch.equal(fit2, fit)           # deterministic run to run
    assert fit.shape == (1, len(g['fitness'])) and T.shape == (1, 4, 4)
    assert fit.max().item() == g['fitness'].max()


@pytest.mark.parametrize('name', CASES)
def test_stage_parity_with_oracle_upstream(golden_dir, name):
    """Each integer stage fed the oracle's upstream tensor through the hooks must be bit-exact."""
    g, cfg, src, tgt = _load(golden_dir, name)
    m = _matcher(cfg)
    # (a) confidence -> seeds (NMS + stable ranking)
    det = {}
    conf = torch.from_numpy(g['confidence'])[None]
    T, fit, labels = m._run(src, tgt, want_labels=True, detail=det, hooks=dict(confidence=conf))
    np.testing.assert_array_equal(det['seeds'][0].cpu().numpy(), g['st_seeds'])
    # (b) seeds -> top-k1 / top-k2 index sets, stopping iteration, fitness, best seed
    np.testing.assert_array_equal(det['topk1'][0].cpu().numpy(), g['st_topk1'].astype(np.int32))
    np.testing.assert_array_equal(det['topk2'][0].cpu().numpy(), g['st_topk2'].astype(np.int32))
    assert int(det['local_iters'][0]) == int(g['st_local_iters'])
    np.testing.assert_allclose(det['seed_weights'][0].cpu().numpy(), g['st_seed_weights'], rtol=1e-3, atol=1e-6)
    fit_gpu, fit_ref = fit[0].cpu().numpy(), g['st_fitness']
    # a seed hypothesis whose fp32 pose differs in the last bits can move a borderline point across 0.6 m
    assert np.abs(fit_gpu - fit_ref).max() <= 2 and (fit_gpu != fit_ref).mean() < 0.02, \
        (np.abs(fit_gpu - fit_ref).max(), (fit_gpu != fit_ref).mean())
    assert int(det['best_seed'][0]) == int(g['st_best_seed'])
    _pose_close(det['initial_trans'][0].cpu().numpy(), g['st_initial_trans'])
    # (c) refinement + labels
    nfit = int(det['refine_counts'][0, 0])
    np.testing.assert_array_equal(det['refine_counts'][0, 1:1 + nfit].cpu().numpy(), g['st_refine_counts'])
    _pose_close(T[0].cpu().numpy(), g['st_final_trans'])
    assert int((labels[0].cpu().numpy() != g['st_labels']).sum()) == 0
    # (d) refinement alone from the oracle's initial transform
    T3, _, labels3 = m._run(src, tgt, want_labels=True, hooks=dict(initial_trans=torch.from_numpy(g['st_initial_trans'])[None]))
    _pose_close(T3[0].cpu().numpy(), g['st_final_trans'])
    assert int((labels3[0].cpu().numpy() != g['st_labels']).sum()) == 0


def test_first_order_bits_bit_exact(golden_dir):
    """hard / tight compatibility bit matrices vs the oracle's dense fp32 masks (SC2_PCR.py:333-342,357)."""
    from oracle import sc2pcr_oracle as O
    g, cfg, src, tgt = _load(golden_dir, 'sc2pcr_n1000_s1')
    m = _matcher(cfg)
    det = {}
    m._run(src, tgt, want_labels=False, detail=det)
    ocfg = O.SC2Config(**{k: cfg[k] for k in ('inlier_threshold', 'num_node', 'd_thre', 'num_iterations', 'ratio',
                                               'nms_radius', 'max_points', 'k1', 'k2')})
    _, _, _, hard, tight = O.first_order(src.cpu(), tgt.cpu(), ocfg)
    n = src.shape[1]
    for name, want in (('hard_bits', hard), ('tight_bits', tight)):
        words = det[name][0].cpu().numpy().view(np.uint32)
        bits = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(n, -1)[:, :n]
        np.testing.assert_array_equal(bits.astype(np.float32), want[0].numpy())


def test_batched_equals_loop(golden_dir):
    """bs > 1 (extension over the reference): every item equals its own single-pair run."""
    g2, cfg, s2, t2 = _load(golden_dir, 'sc2pcr_n2000_s2')
    g3, _, s3, t3 = _load(golden_dir, 'sc2pcr_n2000_s3')
    m = _matcher(cfg)
    src, tgt = torch.cat([s2, s3, s2]), torch.cat([t2, t3, t2])
    T, fit, lab = m._run(src, tgt, want_labels=True)
    for b, (s, t) in enumerate(((s2, t2), (s3, t3), (s2, t2))):
        T1, f1, l1 = m._run(s, t, want_labels=True)
        assert torch.equal(T[b], T1[0]) and torch.equal(fit[b], f1[0]) and torch.equal(lab[b], l1[0])


def test_estimator_api_and_rng_order(golden_dir):
    """Matcher.estimator: 5-tuple, numpy RNG draw order of match_pair (SC2_PCR.py:288-289)."""
    from oracle import sc2pcr_oracle as O
    from eyoc_b200 import synth
    rng = np.random.default_rng(3)
    n = 1500
    xyz0 = rng.uniform(-40, 40, (n, 3)).astype(np.float32) * np.array([1, 1, 0.1], np.float32)
    T_gt = np.eye(4, dtype=np.float32)
    T_gt[:3, :3] = synth._yaw(0.2)
    T_gt[:3, 3] = [3.0, -2.0, 0.1]
    xyz1 = (xyz0 @ T_gt[:3, :3].T + T_gt[:3, 3]).astype(np.float32)
    f0, f1, hit = synth.planted_descriptors(xyz0, xyz1, T_gt, rng, sigma=0.05)
    from eyoc_b200.scripts.SC2_PCR.SC2_PCR import Matcher
    m = Matcher(inlier_threshold=0.6, num_node=2000, use_mutual=False, d_thre=0.1, num_iterations=20, ratio=0.2,
                nms_radius=0.6, max_points=8000, k1=30, k2=20)
    args = [torch.from_numpy(a)[None] for a in (xyz0, xyz1, f0, f1)]
    np.random.seed(0)
    T, labels, sc, tc, fit = m.estimator(*[a.cuda() for a in args])
    np.random.seed(0)
    ocfg = O.SC2Config(num_node=2000, stable_ties=True)
    det = {}
    T_o, labels_o, sc_o, tc_o, fit_o = O.estimator(*args, ocfg, det, dense_weight=False)
    assert torch.equal(sc.cpu(), sc_o) and torch.equal(tc.cpu(), tc_o)      # identical correspondence sets
    _pose_close(T[0].cpu().numpy(), T_o[0].numpy())
    assert torch.equal(labels.cpu(), labels_o)
    assert T.shape == (1, 4, 4) and labels.shape == (1, 2000) and fit.shape == (1, 400)
    _pose_close(T[0].cpu().numpy(), T_gt)


def test_kabsch_golden(golden_dir):
    from eyoc_b200.scripts.SC2_PCR.common import rigid_transform_3d
    g = np.load(f'{golden_dir}/kabsch_5x20.npz')
    A, B, w = (torch.from_numpy(g[k]).cuda() for k in ('A', 'B', 'w'))
    T = rigid_transform_3d(A, B, w.clone())
    for b in range(5):
        _pose_close(T[b].cpu().numpy(), g['T'][b])
    w2 = w.clone()
    w2[0, :5] = -1.0
    rigid_transform_3d(A, B, w2)
    assert (w2[0, :5] == 0).all()                      # in-place zeroing like common.py:20


def test_degenerate_inputs():
    from eyoc_b200.scripts.SC2_PCR.SC2_PCR import Matcher
    m = Matcher(inlier_threshold=0.6, num_node='all', d_thre=0.1, num_iterations=20, ratio=0.2, nms_radius=0.6)
    with pytest.raises(RuntimeError):
        m.SC2_PCR(torch.zeros(1, 0, 3).cuda(), torch.zeros(1, 0, 3).cuda())
    with pytest.raises(RuntimeError):
        m.SC2_PCR(torch.zeros(1, 3, 3).cuda(), torch.zeros(1, 3, 3).cuda())       # int(3*0.2) = 0 seeds
    with pytest.raises(RuntimeError):
        m.SC2_PCR(torch.zeros(1, 100, 3), torch.zeros(1, 100, 3))                 # CPU tensors: no fallback
    # all-identical points (rank-0 H): must not hang or produce NaN
    T, fit = m.SC2_PCR(torch.ones(1, 50, 3).cuda(), torch.ones(1, 50, 3).cuda())
    assert torch.isfinite(T).all()
