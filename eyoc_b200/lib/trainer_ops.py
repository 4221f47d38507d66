"""The EYOC labeler's correspondence path on the sm_100a kernels, batched over a training batch.

Drop-ins for the pieces of the reference's ``lib/trainer.py`` that sit between the labeler network and the loss:

    knn_points                 pytorch3d.ops.knn.knn_points as lib/trainer.py:1064-1065,1182 uses it (K = 1 or 2)
    calculate_ratio_test       lib/trainer.py:993-1010
    get_topk_matches           lib/trainer.py:1012-1016
    match_and_filter_corr      lib/trainer.py:1025-1151   mutual K-NN in feature space, Lowe ratio, spherical filter
    corr_through_registration  lib/trainer.py:1153-1224   SC2-PCR per pair, then 3-D nearest neighbours under the pose

The authors' ToDo at lib/trainer.py:1157 ("Fix SC2-PCR so that this loop can be batched and parallelized") is what
``corr_through_registration`` does here: pairs with the same number of putative correspondences go through ONE batched
``Matcher.SC2_PCR`` call (bs > 1), all groups are queued on the stream back to back, and the poses come back with a single
device-to-host copy instead of one blocking ``.cpu()`` per pair.

pytorch3d is not installed in this image (and not vendored by the reference): ``knn_points`` restates its published
semantics - squared L2 distances accumulated channel by channel (one FMA per channel, ascending), the K smallest in
ascending order, ties to the lowest index, zeros in the padded rows of ragged batches.  Parity unpinned against pytorch3d
itself; pinned against oracle/labeler_oracle.py (tests/test_labeler_gpu.py).
"""
from collections import namedtuple

import numpy as np
import torch

from .. import _C
from .eval import knn1

_KNN = namedtuple('KNN', 'dists idx knn')


def _knn_excluding(q, r, exclude):
    """Nearest neighbour of every row of q [B, nq, D] in r [B, nr, D] ignoring column exclude[b, i] (eyoc_knn1_excluding)."""
    B, nq, dim = q.shape
    nr = r.shape[1]
    lib = _C.lib()
    idx = torch.empty((B, nq), dtype=torch.int64, device=q.device)
    dist = torch.empty((B, nq), dtype=torch.float32, device=q.device)
    ws = torch.empty(max(lib.eyoc_knn1_workspace_bytes(_C.c_int(B), _C.c_int64(nq)), 8), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        _C.check(lib.eyoc_knn1_excluding(_C.ptr(q), _C.ptr(r), _C.c_int(B), _C.c_int64(nq), _C.c_int64(nr), _C.c_int(dim),
                                         _C.c_int(0), _C.ptr(exclude.contiguous()), _C.ptr(ws), _C.c_size_t(ws.numel()),
                                         _C.ptr(idx), _C.ptr(dist), _C.stream()))
    return idx, dist


def _knn_k(q, r, K):
    """q [B, nq, D], r [B, nr, D] -> (dists [B, nq, K], idx [B, nq, K]) ascending."""
    i1, d1 = knn1(q, r, form=0, return_distance=True)
    if K == 1:
        return d1[:, :, None], i1[:, :, None]
    i2, d2 = _knn_excluding(q, r, i1)
    return torch.stack([d1, d2], 2), torch.stack([i1, i2], 2)


def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, version=-1, return_nn=False, return_sorted=True):
    """pytorch3d.ops.knn_points for K in {1, 2}: p1 [N, P1, D], p2 [N, P2, D] (padded), lengths [N] ->
    KNN(dists [N, P1, K] squared L2 ascending, idx [N, P1, K] int64, knn None).  Rows beyond lengths1 hold zeros."""
    if K not in (1, 2):
        raise NotImplementedError('knn_points: K = 1 or 2 (all the reference path uses, lib/trainer.py:1062)')
    if return_nn:
        raise NotImplementedError('knn_points: return_nn is not used on the reference path')
    _C.require_cuda(p1, p2)
    p1, p2 = _C.f32c(p1), _C.f32c(p2)
    N, P1, D = p1.shape
    P2 = p2.shape[1]
    l1 = [P1] * N if lengths1 is None else [int(v) for v in lengths1.tolist()]
    l2 = [P2] * N if lengths2 is None else [int(v) for v in lengths2.tolist()]
    if min(l2, default=K) < K:
        raise RuntimeError(f'knn_points: a reference cloud has fewer than K = {K} points')
    dists = torch.zeros((N, P1, K), dtype=torch.float32, device=p1.device)
    idx = torch.zeros((N, P1, K), dtype=torch.int64, device=p1.device)
    if N and all(a == P1 for a in l1) and all(b == P2 for b in l2):
        d, i = _knn_k(p1, p2, K)                                    # one batched launch sequence
        return _KNN(d, i, None)
    for n in range(N):                                              # ragged: per cloud, no host synchronisation in between
        if l1[n] == 0:
            continue
        d, i = _knn_k(p1[n:n + 1, :l1[n]], p2[n:n + 1, :l2[n]], K)
        dists[n, :l1[n]] = d[0]
        idx[n, :l1[n]] = i[0]
    return _KNN(dists, idx, None)


def calculate_ratio_test(dists):
    """lib/trainer.py:993-1010: (N, P, 2) cosine similarity to the two nearest neighbours -> (N, P, 1) weight, higher = more unique."""
    dists = (1 - dists).clamp(min=1e-9)
    ratio = dists[:, :, 0:1] / dists[:, :, 1:2]
    return 1 - ratio


def get_topk_matches(dists, idx, num_corres):
    """lib/trainer.py:1012-1016."""
    num_corres = min(num_corres, dists.shape[1])
    dist, idx_source = torch.topk(dists, k=num_corres, dim=1)
    idx_target = idx.gather(1, idx_source)
    return idx_source, idx_target, dist


def _pad(tensors):
    """pytorch3d.structures.Pointclouds(...).points_padded() / features_padded(): zero padding to the longest cloud."""
    n = max(len(t) for t in tensors)
    out = tensors[0].new_zeros((len(tensors), n, tensors[0].shape[1]))
    for i, t in enumerate(tensors):
        out[i, :len(t)] = t
    lengths = torch.tensor([len(t) for t in tensors], dtype=torch.int64, device=tensors[0].device)
    return out, lengths


def match_and_filter_corr(C_batch_0, F_batch_0, C_batch_1, F_batch_1, radius=20, feature_filter='Lowe',
                          spatial_filter='Spherical', frame_distance=None, dist_sim_map=None, similarity_thresh=0.4):
    """lib/trainer.py:1025-1151.  Lists of per-pair coordinates [n_i, 3] and features [n_i, d] -> (matches [N, 2] int64 CPU,
    collated with the per-cloud row offsets; uncollated_matches: list of [m_i, 2] int64 on the feature device).
    ``spatial_filter='Similarity'`` needs the lookup tables the reference loads from config/dist_sim_plot (a dict
    {0..5: tensor}) passed as ``dist_sim_map``."""
    dev = F_batch_0[0].device
    C_batch_0 = [c.to(dev) for c in C_batch_0]
    C_batch_1 = [c.to(dev) for c in C_batch_1]
    num_corres = 5000
    P1_F, P1_N = _pad(list(F_batch_0))
    P2_F, P2_N = _pad(list(F_batch_1))
    assert feature_filter in ['None', 'Lowe']
    assert spatial_filter in ['Spherical', 'Similarity', 'None']
    K = 1 if feature_filter == 'None' else 2
    dists_1, idx_1, _ = knn_points(P1_F, P2_F, P1_N, P2_N, K=K)
    dists_2, idx_2, _ = knn_points(P2_F, P1_F, P2_N, P1_N, K=K)
    idx_1 = idx_1[:, :, 0:1]
    idx_2 = idx_2[:, :, 0:1]
    if feature_filter == 'Lowe':
        weights_1 = calculate_ratio_test(1 - 0.5 * dists_1)
        weights_2 = calculate_ratio_test(1 - 0.5 * dists_2)
    else:
        weights_1 = dists_1[:, :, 0:1]
        weights_2 = dists_2[:, :, 0:1]
    n_corres_1 = min(num_corres, int(P1_N.min()))
    n_corres_2 = min(num_corres, int(P2_N.min()))
    m12_idx1, m12_idx2, _ = get_topk_matches(weights_1, idx_1, n_corres_1)
    m21_idx2, m21_idx1, _ = get_topk_matches(weights_2, idx_2, n_corres_2)
    matches_idx1 = torch.cat((m12_idx1, m21_idx1), dim=1)
    matches_idx2 = torch.cat((m12_idx2, m21_idx2), dim=1)
    bias_1 = torch.cumsum(torch.tensor([0] + [len(f) for f in F_batch_0][:-1], dtype=torch.int64, device=dev), 0)
    bias_2 = torch.cumsum(torch.tensor([0] + [len(f) for f in F_batch_1][:-1], dtype=torch.int64, device=dev), 0)
    match_1 = (matches_idx1 + bias_1[:, None, None]).reshape(-1, 1)
    match_2 = (matches_idx2 + bias_2[:, None, None]).reshape(-1, 1)
    matches = torch.cat([match_1, match_2], dim=1)
    uncollated_matches = []
    for i in range(len(C_batch_0)):
        i1, i2 = matches_idx1[i].squeeze(1), matches_idx2[i].squeeze(1)
        if spatial_filter == 'None':
            mask = torch.ones_like(i1, dtype=torch.bool)
        elif spatial_filter == 'Spherical':
            mask = (torch.norm(C_batch_0[i][i1], dim=1) > radius) & (torch.norm(C_batch_1[i][i2], dim=1) > radius)
        else:
            if dist_sim_map is None:
                raise RuntimeError("spatial_filter='Similarity' needs dist_sim_map (config/dist_sim_plot/*_distSimPlot.npz tables)")
            d0 = torch.norm(C_batch_0[i][i1], dim=1)
            d1 = torch.norm(C_batch_1[i][i2], dim=1)
            d1_tmp = torch.abs(d0 - d1)
            d0 = torch.min(torch.vstack([d0, d1]), dim=0).values
            d1 = d1_tmp
            frame_index = min(max(0, int(frame_distance[i]) // 5), 5)
            table = dist_sim_map[frame_index].to(dev)
            xlim, ylim = table.shape
            gridsize = [5, {0: 1, 1: 1.5, 2: 2, 3: 2.5, 4: 2.5, 5: 2.5}[frame_index]]
            d0 = (d0 / gridsize[0]).long().clamp(0, ylim - 1)
            d1 = (d1 / gridsize[1]).long().clamp(0, xlim - 1)
            mask = table[d1, d0] > similarity_thresh
        uncollated_matches.append(torch.cat([matches_idx1[i][mask], matches_idx2[i][mask]], dim=1))
    return matches.cpu().detach(), uncollated_matches


def corr_through_registration(input_dict, uncollated_pairs, matcher, device=None):
    """lib/trainer.py:1153-1224 with the SC2-PCR loop batched: -> (T_ransac list of [4,4] float32 numpy,
    correspondences [M, 2] int64 collated, [], fitnesses list of [1, S_i], uncollated_corr list of [m_i, 2]).
    Draws ``torch.randperm`` from the global CPU generator once per pair, in order, like the reference (:1199)."""
    n_pairs = len(uncollated_pairs)
    dev = device if device is not None else uncollated_pairs[0].device
    pcd0 = [p.to(dev).float() for p in input_dict['pcd0'][:n_pairs]]
    pcd1 = [p.to(dev).float() for p in input_dict['pcd1'][:n_pairs]]
    # ---- SC2-PCR: one batched call per group of equal correspondence count (after the max_points cut of SC2_PCR.py:324-327)
    srcs, tgts, groups = [], [], {}
    for i, pairs in enumerate(uncollated_pairs):
        pairs = pairs.to(dev)
        srcs.append(pcd0[i][pairs[:, 0]])
        tgts.append(pcd1[i][pairs[:, 1]])
        groups.setdefault(min(len(pairs), matcher.max_points), []).append(i)
    results, fitnesses = [None] * n_pairs, [None] * n_pairs
    for n, members in groups.items():
        src = torch.stack([srcs[i][:n] for i in members])
        tgt = torch.stack([tgts[i][:n] for i in members])
        T, fit = matcher.SC2_PCR(src, tgt)
        for j, i in enumerate(members):
            results[i], fitnesses[i] = T[j], fit[j:j + 1]
    T_all = torch.stack(results)                                     # [n_pairs, 4, 4] on the device
    T_ransac = list(T_all.cpu().float().numpy())                      # the single device-to-host copy of the poses
    # ---- nearest neighbours in 3-D under the estimated pose (K = 1, per pair: the clouds are ragged)
    C_batch_0 = [pcd0[i] @ T_all[i, :3, :3].T + T_all[i, :3, 3] for i in range(n_pairs)]
    idx_1 = [knn1(C_batch_0[i], pcd1[i], form=0) for i in range(n_pairs)]
    P1_N = [len(p) for p in pcd0]
    P2_N = [len(p) for p in pcd1]
    bias_1 = np.concatenate([[0], np.cumsum(P1_N)])
    bias_2 = np.concatenate([[0], np.cumsum(P2_N)])
    correspondences, uncollated_corr = [], []
    for i in range(n_pairs):
        pos_sel_1 = torch.randperm(P1_N[i])[:min(P1_N[i], 5000)].to(dev)
        src_k = C_batch_0[i][pos_sel_1]
        tgt_k = pcd1[i][idx_1[i][pos_sel_1]]
        within = torch.norm(src_k - tgt_k, dim=1) < 2                 # loose on purpose: tolerates pose error (:1204-1205)
        pos_sel_1 = pos_sel_1[within]
        pos_sel_2 = idx_1[i][pos_sel_1]
        uncollated_corr.append(torch.stack([pos_sel_1, pos_sel_2], 1))
        correspondences.append(torch.stack([pos_sel_1 + int(bias_1[i]), pos_sel_2 + int(bias_2[i])], 1))
    return T_ransac, torch.cat(correspondences, 0), [], fitnesses, uncollated_corr
