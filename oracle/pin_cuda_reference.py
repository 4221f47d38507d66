"""Pin the tie rule of the CUDA path against the reference's REAL backend: torch-CUDA.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs on a GPU box:

    python -m oracle.pin_cuda_reference [--out gpurun_out/cuda_goldens]

The reference sorts three times without `stable` (scripts/SC2_PCR/SC2_PCR.py:53,84,105); which of several equal scores
comes first is then the sort backend's business.  oracle/pin_against_reference.py pins the torch-CPU behaviour (an
AVX-512 quicksort order).  The reference's production backend is torch-CUDA (`.cuda()` hard-coded at SC2_PCR.py:299), so
this script runs the SAME torch ops on the GPU - the reference's own `Matcher` when /root/reference is mounted (it is not
on the gpurun boxes: logged), else oracle/sc2pcr_oracle.py, whose ops are bit-identical to the reference's on the CPU
(pin_against_reference.py) and device-generic (SVD on the CPU like common.py:36) - with `stable_ties=False`, i.e. plain
`argsort(descending=True)` exactly as the reference issues it, and records seeds / top-k1 / top-k2 / fitness / labels as
`cuda_*` goldens.  It then reports (a) whether torch-CUDA's unstable-flag sort equals the stable rule (descending value,
lowest index first) on these inputs, and (b) how the eyoc_b200 kernels compare, stage by stage with the torch-CUDA
upstream tensors fed through eyoc_sc2_hooks, and end to end.  The goldens are copied to tests/golden/ and committed;
tests/test_sc2pcr_gpu.py::test_stage_parity_vs_torch_cuda then holds the kernels to them.
"""
import argparse
import dataclasses
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
CASES = ['sc2pcr_n1000_s1', 'sc2pcr_n2000_s2', 'sc2pcr_n2000_s3', 'sc2pcr_n25_s4', 'sc2pcr_n8000_s5']


def _reference_matcher():
    if not os.path.isdir(REF):
        return None, f'{REF} is not mounted on this box'
    try:
        for name in ('open3d', 'MinkowskiEngine'):
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.path.insert(0, REF)
        from scripts.SC2_PCR.SC2_PCR import Matcher
        sys.path.remove(REF)
        return Matcher, 'reference Matcher imported'
    except Exception as e:          # noqa: BLE001
        return None, f'reference import failed: {e!r}'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'cuda_goldens'))
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit('pin_cuda_reference: needs a CUDA device')
    sys.path.insert(0, ROOT)
    from oracle import sc2pcr_oracle as O
    from eyoc_b200.scripts.SC2_PCR.SC2_PCR import Matcher
    os.makedirs(args.out, exist_ok=True)
    dev = torch.device('cuda', 0)
    RefMatcher, why = _reference_matcher()
    lines = [f'# torch {torch.__version__} CUDA {torch.version.cuda} on {torch.cuda.get_device_name(0)}; {why}']
    for name in CASES:
        g = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
        cfgd = json.loads(str(g['cfg']))
        cfg = O.SC2Config(**{k: cfgd[k] for k in ('inlier_threshold', 'num_node', 'd_thre', 'num_iterations', 'ratio',
                                                  'nms_radius', 'max_points', 'k1', 'k2')})
        src = torch.from_numpy(g['src'])[None].to(dev)
        tgt = torch.from_numpy(g['tgt'])[None].to(dev)
        # ---- the reference's ops on torch-CUDA (plain argsort, as the reference issues it)
        det = {}
        T_c, fit_c = O.sc2_pcr(src.clone(), tgt.clone(), cfg, det)
        labels_c = (torch.sum((O.se3_transform(src, T_c) - tgt) ** 2, dim=-1) ** 0.5 < cfg.inlier_threshold)
        if RefMatcher is not None:
            m = RefMatcher(inlier_threshold=cfg.inlier_threshold, num_node=cfg.num_node, use_mutual=False, d_thre=cfg.d_thre,
                           num_iterations=cfg.num_iterations, ratio=cfg.ratio, nms_radius=cfg.nms_radius,
                           max_points=cfg.max_points, k1=cfg.k1, k2=cfg.k2)
            T_r, fit_r = m.SC2_PCR(src.clone(), tgt.clone())
            assert torch.equal(T_r, T_c) and torch.equal(fit_r, fit_c), 'oracle-on-CUDA differs from the reference-on-CUDA'
        # ---- the same with stable=True: is torch-CUDA's sort the stable rule on these inputs?
        ds = {}
        T_s, fit_s = O.sc2_pcr(src.clone(), tgt.clone(), dataclasses.replace(cfg, stable_ties=True), ds)
        stable_same = all(torch.equal(det[k], ds[k]) for k in ('seeds', 'topk1', 'topk2')) and torch.equal(fit_c, fit_s)
        # ---- torch-CUDA vs the pinned torch-CPU reference output
        cpu_seeds_same = float((det['seeds'][0].cpu().numpy() == g['seeds']).mean())
        cpu_label_ham = int((labels_c[0].cpu().numpy() != g['labels']).sum())
        # ---- eyoc_b200 kernels: end to end, then stage by stage on the torch-CUDA upstream tensors
        mk = Matcher(inlier_threshold=cfg.inlier_threshold, num_node=cfg.num_node, use_mutual=False, d_thre=cfg.d_thre,
                     num_iterations=cfg.num_iterations, ratio=cfg.ratio, nms_radius=cfg.nms_radius,
                     max_points=cfg.max_points, k1=cfg.k1, k2=cfg.k2)
        d0 = {}
        T_k, fit_k, lab_k = mk._run(src, tgt, want_labels=True, detail=d0)
        e2e = dict(seeds=float((d0['seeds'][0] == det['seeds'][0].int()).float().mean()),
                   topk1_rows=float((d0['topk1'][0] == det['topk1'][0].int()).all(-1).float().mean()),
                   fitness=float((fit_k[0] == fit_c[0]).float().mean()),
                   labels_hamming=int((lab_k[0] != labels_c[0].float()).sum()),
                   dR=float(torch.linalg.norm(T_k[0, :3, :3] - T_c[0, :3, :3])),
                   dt=float(torch.linalg.norm(T_k[0, :3, 3] - T_c[0, :3, 3])))
        d1 = {}
        mk._run(src, tgt, want_labels=True, detail=d1, hooks=dict(confidence=det['confidence']))
        d2 = {}
        _, fit_h, _ = mk._run(src, tgt, want_labels=True, detail=d2, hooks=dict(seeds=det['seeds'].int()))
        stage = dict(seeds_given_conf=bool(torch.equal(d1['seeds'][0], det['seeds'][0].int())),
                     topk1_given_seeds=bool(torch.equal(d2['topk1'][0], det['topk1'][0].int())),
                     topk2_given_seeds=bool(torch.equal(d2['topk2'][0], det['topk2'][0].int())),
                     fitness_given_seeds_equal=float((fit_h[0] == fit_c[0]).float().mean()),
                     fitness_given_seeds_maxdiff=float((fit_h[0] - fit_c[0]).abs().max()))
        lines.append(f'{name}: torch-CUDA plain argsort == stable rule: {stable_same}; vs torch-CPU reference: seeds equal '
                     f'{cpu_seeds_same:.4f}, label hamming {cpu_label_ham}')
        lines.append(f'    kernels end to end vs torch-CUDA: {json.dumps(e2e)}')
        lines.append(f'    kernels stage parity (torch-CUDA upstream through hooks): {json.dumps(stage)}')
        np.savez_compressed(
            os.path.join(args.out, 'cuda_' + name + '.npz'),
            cuda_confidence=det['confidence'][0].cpu().numpy(), cuda_global_iters=det['global_iters'],
            cuda_seeds=det['seeds'][0].cpu().numpy().astype(np.int32),
            cuda_topk1=det['topk1'][0].cpu().numpy().astype(np.int16), cuda_topk2=det['topk2'][0].cpu().numpy().astype(np.int16),
            cuda_local_iters=det['local_iters'], cuda_fitness=fit_c[0].cpu().numpy(),
            cuda_best_seed=int(det['best_seed'][0]), cuda_initial_trans=det['initial_trans'][0].cpu().numpy(),
            cuda_refine_counts=np.array(det['refine_counts']), cuda_final_trans=T_c[0].cpu().numpy(),
            cuda_labels=labels_c[0].cpu().numpy(), cuda_plain_equals_stable=stable_same,
            backend=f'torch {torch.__version__} / {torch.cuda.get_device_name(0)} / ' + ('reference Matcher' if RefMatcher else 'oracle ops'))
    report = '\n'.join(lines)
    print(report)
    with open(os.path.join(args.out, 'pin_cuda_report.txt'), 'w') as f:
        f.write(report + '\n')


if __name__ == '__main__':
    main()
