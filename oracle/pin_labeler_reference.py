"""Pin oracle/labeler_oracle.py against the reference's own labeler methods and write tests/golden/labeler_*.npz.

Runs ONLY in the build container (needs /root/reference):

    python -m oracle.pin_labeler_reference

lib/trainer.py is imported unmodified with its unavailable third-party modules stubbed (MinkowskiEngine, open3d,
tensorboardX, easydict, the reference's own ME-based `model` package) and with **pytorch3d** replaced by the restatement in
oracle/labeler_oracle.py (`knn_points`, `Pointclouds`: pytorch3d is absent from this image - those two symbols stay
"parity unpinned", everything around them is the reference's code).  `CorrespondenceExtensionTrainer.match_and_filter_corr`
(:1025-1151) and `.corr_through_registration` (:1153-1224) are then called as plain functions on a stand-in `self` that
carries the reference `Matcher`, and compared bit for bit with the oracle restatement on the same seeded CPU tensors.
"""
import codecs
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')
KITTI = dict(inlier_threshold=0.6, num_node=8000, use_mutual=False, d_thre=0.1, num_iterations=20, ratio=0.2,
             nms_radius=0.6, max_points=8000, k1=30, k2=20)


def labeler_inputs(pair_ids=(40, 41), az_step_deg=1.2, sigma=0.1):
    """Seeded inputs shared by this script and tests/test_labeler_gpu.py: small synthetic pairs with planted descriptors."""
    sys.path.insert(0, ROOT)
    from eyoc_b200 import synth
    C0, C1, F0, F1 = [], [], [], []
    for k, pid in enumerate(pair_ids):
        p = synth.make_pair(pid, distance=7.0 + 4.0 * k, az_step_deg=az_step_deg)
        f0, f1, _ = synth.planted_descriptors(p['xyz0'], p['xyz1'], p['T_gt'], np.random.default_rng(500 + pid), sigma=sigma)
        C0.append(torch.from_numpy(p['xyz0']))
        C1.append(torch.from_numpy(p['xyz1']))
        F0.append(torch.from_numpy(f0))
        F1.append(torch.from_numpy(f1))
    return C0, F0, C1, F1


def import_reference_trainer():
    from oracle import labeler_oracle as LO
    codecs.register(lambda name: codecs.lookup('utf-8') if name in ('future_fstrings', 'future-fstrings') else None)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    for n in ('open3d', 'MinkowskiEngine'):
        sys.modules.setdefault(n, types.ModuleType(n))
    stub('tensorboardX', SummaryWriter=object)
    stub('easydict', EasyDict=dict)
    stub('model', load_model=lambda name: None)                      # the reference's ME network definitions: not on this path
    st = stub('pytorch3d.structures', Pointclouds=LO.Pointclouds)
    stub('pytorch3d.ops.knn', knn_points=LO.knn_points)
    stub('pytorch3d.ops')
    stub('pytorch3d', structures=st)
    sys.path.insert(0, REF)
    import lib.trainer as T
    from scripts.SC2_PCR.SC2_PCR import Matcher
    sys.path.remove(REF)
    sys.modules.pop('model', None)
    return T, Matcher


def main():
    sys.path.insert(0, ROOT)
    from oracle import labeler_oracle as LO, sc2pcr_oracle as O
    T, Matcher = import_reference_trainer()
    cls = T.CorrespondenceExtensionTrainer
    me = types.SimpleNamespace(matcher=Matcher(**KITTI), device='cpu')
    me.calculate_ratio_test = types.MethodType(cls.calculate_ratio_test, me)
    me.get_topk_matches = types.MethodType(cls.get_topk_matches, me)
    torch.set_num_threads(os.cpu_count())
    C0, F0, C1, F1 = labeler_inputs()
    radius = 8.0

    def eq(a, b, what):
        if not torch.equal(torch.as_tensor(a), torch.as_tensor(b)):
            raise SystemExit(f'PIN FAILED: {what}')
        print(f'  pinned  {what}')

    out = {}
    for ff, sf in (('Lowe', 'Spherical'), ('None', 'None')):
        m_ref, u_ref = cls.match_and_filter_corr(me, [c.clone() for c in C0], F0, [c.clone() for c in C1], F1, radius=radius,
                                                 feature_filter=ff, spatial_filter=sf)
        m_or, u_or = LO.match_and_filter_corr(C0, F0, C1, F1, radius=radius, feature_filter=ff, spatial_filter=sf)
        eq(m_ref, m_or, f'match_and_filter_corr[{ff},{sf}] collated matches')
        for i, (a, b) in enumerate(zip(u_ref, u_or)):
            eq(a, b, f'match_and_filter_corr[{ff},{sf}] uncollated pair {i} ({len(a)} matches)')
        out[f'unc_{ff}_{sf}'] = u_ref
    unc = out['unc_Lowe_Spherical']
    ocfg = O.SC2Config(**{k: KITTI[k] for k in ('inlier_threshold', 'num_node', 'd_thre', 'num_iterations', 'ratio', 'nms_radius',
                                                'max_points', 'k1', 'k2')})
    inp = dict(pcd0=C0, pcd1=C1)
    torch.manual_seed(77)
    T_ref, corr_ref, _, fit_ref, ucorr_ref = cls.corr_through_registration(me, inp, unc)
    torch.manual_seed(77)
    T_or, corr_or, _, fit_or, ucorr_or = LO.corr_through_registration(inp, unc, ocfg)
    for i in range(len(unc)):
        eq(T_ref[i], T_or[i], f'corr_through_registration pose {i}')
        eq(fit_ref[i], fit_or[i], f'corr_through_registration fitness {i}')
        eq(ucorr_ref[i], ucorr_or[i], f'corr_through_registration correspondences {i} ({len(ucorr_ref[i])})')
    eq(corr_ref, corr_or, 'corr_through_registration collated correspondences')
    save = dict(radius=radius, T_ransac=np.stack(T_ref), corr=corr_ref.numpy().astype(np.int32))
    for i in range(len(unc)):
        save[f'unc_lowe_sph_{i}'] = out['unc_Lowe_Spherical'][i].numpy().astype(np.int32)
        save[f'unc_none_none_{i}'] = out['unc_None_None'][i].numpy().astype(np.int32)
        save[f'ucorr_{i}'] = ucorr_ref[i].numpy().astype(np.int32)
        save[f'fitness_{i}'] = fit_ref[i].numpy()
    np.savez_compressed(os.path.join(GOLD, 'labeler_2pairs.npz'), **save)
    print('LABELER PINNED', {k: v.shape for k, v in save.items() if hasattr(v, 'shape')})


if __name__ == '__main__':
    main()
