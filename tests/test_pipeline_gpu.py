"""GPU: the batched pair pipeline equals a per-pair loop through the drop-in API, and matches the oracle end to end."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(n_pairs=2, num_node=2000):
    from eyoc_b200 import synth
    from eyoc_b200.model import load_model
    from eyoc_b200.scripts.SC2_PCR.SC2_PCR import Matcher
    from oracle import resunet_oracle as RO
    pairs = []
    for i in range(n_pairs):
        p = synth.make_pair(20 + i, distance=6.0 + 3 * i, az_step_deg=1.2)
        f0, f1, _ = synth.planted_descriptors(p['xyz0'], p['xyz1'], p['T_gt'], np.random.default_rng(i), sigma=0.08)
        p.update(desc0=f0, desc1=f1)
        pairs.append(p)
    sd = RO.make_state_dict(1, 32, 5, seed=1)
    model = load_model('ResUNetBN2C')(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3)
    model.load_state_dict(sd)
    matcher = Matcher(inlier_threshold=0.6, num_node=num_node, use_mutual=False, d_thre=0.1, num_iterations=20, ratio=0.2,
                      nms_radius=0.6, max_points=8000, k1=30, k2=20)
    return pairs, sd, model.cuda().eval(), matcher


def test_batched_pipeline_equals_reference_style_loop():
    from eyoc_b200 import synth
    from eyoc_b200.pipeline import RegistrationPipeline
    from eyoc_b200.scripts.test_kitti import find_corr, random_sample
    from eyoc_b200.sparse import SparseTensor
    pairs, sd, model, matcher = _setup()
    coords, xyz, desc, sizes = synth.collate_pairs(pairs)
    pipe = RegistrationPipeline(model, matcher, subsample_size=3000, num_sample=3000)
    np.random.seed(3)
    out = pipe.run(torch.from_numpy(coords).cuda(), torch.from_numpy(xyz).cuda(), sizes, descriptors=torch.from_numpy(desc).cuda())
    # the same thing written like scripts/test_kitti.py:130-181, pair by pair, same RNG stream
    np.random.seed(3)
    off = 0
    for i, p in enumerate(pairs):
        F = []
        for c in (p['coords0'], p['coords1']):
            cb = torch.from_numpy(synth.collate([c])).cuda()
            F.append(model(SparseTensor(torch.ones(len(cb), 1).cuda(), coordinates=cb)).F)
        n0, n1 = sizes[i]
        assert torch.equal(out['features'][off:off + n0], F[0]) and torch.equal(out['features'][off + n0:off + n0 + n1], F[1])
        off += n0 + n1
        F0, F1 = torch.from_numpy(p['desc0']).cuda(), torch.from_numpy(p['desc1']).cuda()
        find_corr(p['xyz0'], p['xyz1'], F0, F1, subsample_size=3000)
        x0, f0 = random_sample(p['xyz0'], F0, 3000)
        x1, f1 = random_sample(p['xyz1'], F1, 3000)
        T, labels, sc, tc, fit = matcher.estimator(torch.from_numpy(x0)[None].cuda(), torch.from_numpy(x1)[None].cuda(), f0[None], f1[None])
        assert torch.equal(out['src_corr'][i], sc[0]) and torch.equal(out['tgt_corr'][i], tc[0])
        assert torch.equal(out['trans'][i], T[0]) and torch.equal(out['labels'][i], labels[0])
        assert torch.equal(out['fitness'][i], fit[0])
        # and the pose is the planted one
        Tg = p['T_gt']
        assert np.linalg.norm(T[0, :3, 3].cpu().numpy() - Tg[:3, 3]) < 0.3


def test_pipeline_vs_oracle_end_to_end():
    """Whole path vs the oracle on the same seeded inputs: identical correspondence index sets and inlier masks, pose
    within 1e-4 / 1e-3 (BASELINE.json north_star)."""
    from eyoc_b200.scripts.test_kitti import random_sample
    from oracle import matching_oracle as MO, sc2pcr_oracle as O
    pairs, sd, model, matcher = _setup(n_pairs=1, num_node=3000)
    p = pairs[0]
    F0, F1 = torch.from_numpy(p['desc0']), torch.from_numpy(p['desc1'])
    np.random.seed(11)
    x0, f0 = random_sample(p['xyz0'], F0.cuda(), 4000)
    x1, f1 = random_sample(p['xyz1'], F1.cuda(), 4000)
    T, labels, sc, tc, fit = matcher.estimator(torch.from_numpy(x0)[None].cuda(), torch.from_numpy(x1)[None].cuda(), f0[None], f1[None])
    np.random.seed(11)
    ox0, of0 = MO.random_sample(p['xyz0'], F0, 4000)
    ox1, of1 = MO.random_sample(p['xyz1'], F1, 4000)
    det = {}
    T_o, labels_o, sc_o, tc_o, fit_o = O.estimator(torch.from_numpy(ox0)[None], torch.from_numpy(ox1)[None], of0[None], of1[None],
                                                    O.SC2Config(num_node=3000, stable_ties=True), det, dense_weight=False)
    assert torch.equal(matcher._last_match[2][0].cpu(), det['nn_idx'])              # bit-exact correspondence indices
    assert torch.equal(sc.cpu(), sc_o) and torch.equal(tc.cpu(), tc_o)
    assert torch.equal(labels.cpu(), labels_o)                                       # bit-exact inlier mask
    assert float(torch.linalg.norm(T[0, :3, :3].cpu() - T_o[0, :3, :3])) < 1e-4
    assert float(torch.linalg.norm(T[0, :3, 3].cpu() - T_o[0, :3, 3])) < 1e-3


def test_overlapped_runner_equals_sequential_runs():
    """pipeline.OverlappedRunner (stage 3 of block k on a second stream beside stage 1 of block k + 1) returns, block by block,
    the bits of sequential pipe.run calls: same features, correspondences, poses, labels - over several different blocks."""
    from eyoc_b200 import synth
    from eyoc_b200.pipeline import OverlappedRunner, RegistrationPipeline
    from eyoc_b200.model import load_model
    from eyoc_b200.scripts.SC2_PCR.SC2_PCR import Matcher
    from oracle import resunet_oracle as RO
    dev = torch.device('cuda')
    model = load_model('ResUNetBN2C')(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3)
    model.load_state_dict(RO.make_state_dict(1, 32, 5, seed=2))
    m = Matcher(inlier_threshold=0.6, num_node=2000, use_mutual=False, d_thre=0.1, num_iterations=20, ratio=0.2,
                nms_radius=0.6, max_points=8000, k1=30, k2=20)
    pipe = RegistrationPipeline(model.to(dev).eval(), m, subsample_size=2000, num_sample=3000)
    blocks = []
    for b in range(4):
        pairs = [synth.make_pair(100 + 3 * b + i, distance=6.0 + i, az_step_deg=1.0) for i in range(2 + b % 2)]
        coords, xyz, desc, sizes = synth.collate_pairs(pairs)
        rng = np.random.default_rng(b)
        desc = np.concatenate([d for p_ in pairs for d in synth.planted_descriptors(p_['xyz0'], p_['xyz1'], p_['T_gt'], rng, sigma=0.08)[:2]]).astype(np.float32)
        blocks.append(dict(coords=torch.from_numpy(coords).to(dev), xyz=torch.from_numpy(xyz).to(dev), sizes=sizes,
                           descriptors=torch.from_numpy(desc).to(dev)))
    np.random.seed(3)
    plans = [pipe.plan(b['sizes']) for b in blocks]
    seq = [pipe.run(b['coords'], b['xyz'], b['sizes'], plan=p, descriptors=b['descriptors']) for b, p in zip(blocks, plans)]
    torch.cuda.synchronize()
    runner = OverlappedRunner(pipe, dev)
    got, seen = [], []
    for b, p in zip(blocks, plans):
        o = runner.submit(b['coords'], b['xyz'], b['sizes'], plan=p, descriptors=b['descriptors'],
                          after_match=lambda out: seen.append(int(out['labels'].shape[0])))
        if o is not None:
            got.append(o)
    got.append(runner.flush())
    torch.cuda.synchronize()
    assert len(got) == len(seq) == 4 and seen == [len(b['sizes']) for b in blocks]
    for a, b in zip(seq, got):
        for key in ('features', 'trans', 'labels', 'fitness', 'src_corr', 'tgt_corr', 'tgt_corr_idx', 'find_corr_tgt'):
            assert torch.equal(a[key], b[key]), key
