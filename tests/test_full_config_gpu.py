"""GPU: the whole path at the FULL KITTI configuration (scripts/SC2_PCR/config_json/config_KITTI.json: num_node 8000,
1600 seeds; scripts/test_kitti.py: find_corr subsample 5000, random_sample 5000; ~30 k voxels per cloud) through the batched
pipeline, against the CPU oracle pair by pair on the same numpy RNG stream:

* find_corr (scripts/test_kitti.py:28-42): the returned point sets `xyz0[inds0]`, `xyz1[inds1[nn]]`;
* match_pair + SC2-PCR + labels (SC2_PCR.py:280-413): identical correspondence sets and inlier masks, pose within the
  north-star tolerance (1e-4 Frobenius / 1e-3 m).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KITTI = dict(inlier_threshold=0.6, num_node=8000, use_mutual=False, d_thre=0.1, num_iterations=20, ratio=0.2,
             nms_radius=0.6, max_points=8000, k1=30, k2=20)


def _nn_mismatch_is_near_tie(Fq, Fr, got, want):
    """Rows where two nearest-neighbour results differ must be fp32 near-ties (fp64 distances equal to 5e-6 relative: the fp32
    rounding of a 32-term sum of squares)."""
    bad = np.nonzero(got != want)[0]
    for i in bad:
        dg = float(((Fq[i].double() - Fr[got[i]].double()) ** 2).sum())
        dw = float(((Fq[i].double() - Fr[want[i]].double()) ** 2).sum())
        assert abs(dg - dw) <= 5e-6 * max(dg, dw, 1e-12), (int(i), dg, dw)
    return len(bad)


def test_full_kitti_configuration_vs_oracle():
    from eyoc_b200 import synth
    from eyoc_b200.model import load_model
    from eyoc_b200.pipeline import RegistrationPipeline
    from eyoc_b200.scripts.SC2_PCR.SC2_PCR import Matcher
    from oracle import matching_oracle as MO, resunet_oracle as RO, sc2pcr_oracle as O
    pairs = synth.make_pairs([1000, 1001])
    coords, xyz, desc, sizes = synth.collate_pairs(pairs)
    assert min(min(s) for s in sizes) > 20_000
    sd = RO.make_state_dict(1, 32, 5, seed=2)
    model = load_model('ResUNetBN2C')(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3)
    model.load_state_dict(sd)
    pipe = RegistrationPipeline(model.cuda().eval(), Matcher(**KITTI))          # subsample 5000 / sample 5000: the defaults
    dev = torch.device('cuda')
    xyz_d = torch.from_numpy(xyz).to(dev)
    np.random.seed(5)
    out = pipe.run(torch.from_numpy(coords).to(dev), xyz_d, sizes, descriptors=torch.from_numpy(desc).to(dev))
    # ---- the oracle, pair by pair, on the same stream (find_corr -> random_sample x2 -> match_pair)
    np.random.seed(5)
    ocfg = O.SC2Config(**{k: KITTI[k] for k in ('inlier_threshold', 'num_node', 'd_thre', 'num_iterations', 'ratio',
                                                'nms_radius', 'max_points', 'k1', 'k2')}, stable_ties=True)
    for j, p in enumerate(pairs):
        F0, F1 = torch.from_numpy(p['desc0']), torch.from_numpy(p['desc1'])
        state = np.random.get_state()
        fc_src, fc_tgt = MO.find_corr(p['xyz0'], p['xyz1'], F0, F1, subsample_size=5000)
        got_src = xyz_d[out['find_corr_src'][j]].cpu().numpy()
        got_tgt = xyz_d[out['find_corr_tgt'][j]].cpu().numpy()
        assert got_src.shape == (5000, 3) and np.array_equal(got_src, fc_src)
        if not np.array_equal(got_tgt, fc_tgt):            # only fp32 near-ties of the descriptor distance may differ
            after = np.random.get_state()
            np.random.set_state(state)
            inds0 = np.random.choice(len(F0), 5000, replace=False)
            inds1 = np.random.choice(len(F1), 5000, replace=False)
            np.random.set_state(after)
            want_rows = inds1[MO.find_nn(F0[inds0], F1[inds1], nn_max_n=500).numpy()]
            got_rows = out['find_corr_tgt'][j].cpu().numpy() - (sum(sum(s) for s in sizes[:j]) + sizes[j][0])
            assert _nn_mismatch_is_near_tie(F0[inds0], F1, got_rows, want_rows) <= 5
        x0, f0 = MO.random_sample(p['xyz0'], F0, 5000)
        x1, f1 = MO.random_sample(p['xyz1'], F1, 5000)
        det = {}
        T_o, lab_o, sc_o, tc_o, _ = O.estimator(torch.from_numpy(x0)[None], torch.from_numpy(x1)[None], f0[None], f1[None], ocfg,
                                                det, dense_weight=False)
        assert torch.equal(out['src_corr'][j].cpu(), sc_o[0])
        # the oracle's matching goes through torch.matmul (MKL): its summation order - and with it the winner of an fp32
        # near-tie - can change with the alignment of its buffers from run to run (observed: 1 run in 5 flips one row).  Rows
        # whose correspondence differs must be such near-ties (a handful at most); everything else is compared bit for bit.
        same = (out['tgt_corr'][j].cpu() == tc_o[0]).all(1)
        if not bool(same.all()):
            rows = torch.nonzero(~same).flatten().tolist()
            assert len(rows) <= 3, len(rows)
            fs, ft = f0[det['src_sel']].double(), f1[det['tgt_sel']].double()
            tk = torch.from_numpy(x1)[det['tgt_sel']]
            for r in rows:
                c_mine = int(torch.nonzero((tk == out['tgt_corr'][j][r].cpu()).all(1))[0])
                c_or = int(det['nn_idx'][r])
                assert abs(float(fs[r] @ ft[c_mine]) - float(fs[r] @ ft[c_or])) <= 2e-6, (r, c_mine, c_or)
        assert int((out['labels'][j].cpu() != lab_o[0])[same].sum()) == 0
        T = out['trans'][j].cpu()
        assert float(torch.linalg.norm(T[:3, :3] - T_o[0, :3, :3])) < 1e-4
        assert float(torch.linalg.norm(T[:3, 3] - T_o[0, :3, 3])) < 1e-3
        assert out['labels'][j].shape == (8000,) and out['fitness'][j].shape == (1600,)
    # ---- the same block with the NETWORK's descriptors downstream: the chain executes end to end on network output
    np.random.seed(5)
    out_n = pipe.run(torch.from_numpy(coords).to(dev), xyz_d, sizes)
    assert torch.equal(out_n['features'], out['features'])
    assert bool(torch.isfinite(out_n['trans']).all()) and out_n['labels'].shape == (2, 8000)
    # find_corr on network features vs the oracle's NN on the SAME feature tensor (index parity up to fp32 near-ties)
    F = out_n['features'].cpu()
    i0, i1 = out_n['find_corr_src'][0].cpu().numpy(), None
    np.random.seed(5)
    n0, n1 = sizes[0]
    inds0 = np.random.choice(n0, 5000, replace=False)
    inds1 = np.random.choice(n1, 5000, replace=False)
    assert np.array_equal(i0, inds0)
    want_nn = MO.find_nn(F[:n0][inds0], F[n0:n0 + n1][inds1], nn_max_n=500).numpy()
    got_nn_rows = out_n['find_corr_tgt'][0].cpu().numpy() - n0
    _nn_mismatch_is_near_tie(F[:n0][inds0], F[n0:n0 + n1], got_nn_rows, inds1[want_nn])
