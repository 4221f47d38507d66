"""CPU oracle: the EYOC labeler's correspondence path (lib/trainer.py:993-1224 of the reference).

TEST INFRASTRUCTURE (see oracle/__init__.py).  torch-CPU restatement, op for op, of ``calculate_ratio_test`` (:993-1010),
``get_topk_matches`` (:1012-1016), ``match_and_filter_corr`` (:1025-1151) and ``corr_through_registration``
(:1153-1224); oracle/pin_labeler_reference.py runs the reference's own methods (lib/trainer.py imported with its
unavailable third-party modules stubbed) against these and asserts bit-equality.

Third-party dependency holding part of the arithmetic: **pytorch3d** (``pytorch3d.ops.knn.knn_points``,
``pytorch3d.structures.Pointclouds``), un-pinned in the reference (README.md install line, no version) and absent from
this image -> PARITY UNPINNED for those two symbols.  They are restated here from pytorch3d's published behaviour:
``knn_points`` = brute-force squared L2 distance accumulated one fused multiply-add per channel (channels ascending,
pytorch3d/csrc/knn/knn.cu ``dist += diff * diff``), the K smallest in ascending order with ties to the lower index,
zeros in the rows beyond ``lengths1``; ``Pointclouds`` = zero padding to the longest cloud.
"""
from collections import namedtuple

import numpy as np
import torch

from . import sc2pcr_oracle as O
from .matching_oracle import _fma32

_KNN = namedtuple('KNN', 'dists idx knn')


def _knn_numpy(a, b, K):
    """K smallest squared distances per row of a against b, kernel order (sequential fp32 FMA), ties to the lower index."""
    idx = np.empty((len(a), K), np.int64)
    dist = np.empty((len(a), K), np.float32)
    for s in range(0, len(a), 512):
        q = a[s:s + 512]
        acc = np.zeros((len(q), len(b)), np.float32)
        for c in range(a.shape[1]):
            d = q[:, c:c + 1] - b[None, :, c]
            acc = _fma32(d, d, acc)
        rows = np.arange(len(q))
        for k in range(K):                                   # K argmin passes: ascending, lowest index among equals
            idx[s:s + len(q), k] = acc.argmin(1)
            dist[s:s + len(q), k] = acc[rows, idx[s:s + len(q), k]]
            acc[rows, idx[s:s + len(q), k]] = np.inf
    return idx, dist


_C_LIB = False


def _knn(a, b, K, force_numpy=False):
    """The same through the C helper (oracle/csrc/knn_oracle.c: fmaf loops, OpenMP) when it can be built."""
    global _C_LIB
    if _C_LIB is False:
        from .build_c import load
        _C_LIB = load()
    if _C_LIB is None or force_numpy or K > 8:
        return _knn_numpy(a, b, K)
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    idx = np.empty((len(a), K), np.int64)
    dist = np.empty((len(a), K), np.float32)
    _C_LIB.knn_sq_seq(a.ctypes.data, len(a), b.ctypes.data, len(b), a.shape[1], K, idx.ctypes.data, dist.ctypes.data)
    return idx, dist


def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, **_):
    """pytorch3d.ops.knn_points restated (see the module docstring)."""
    N, P1, D = p1.shape
    dists = torch.zeros((N, P1, K), dtype=torch.float32)
    idx = torch.zeros((N, P1, K), dtype=torch.int64)
    for n in range(N):
        l1 = int(lengths1[n]) if lengths1 is not None else P1
        l2 = int(lengths2[n]) if lengths2 is not None else p2.shape[1]
        if l1 == 0:
            continue
        i, d = _knn(p1[n, :l1].numpy().astype(np.float32), p2[n, :l2].numpy().astype(np.float32), K)
        idx[n, :l1] = torch.from_numpy(i)
        dists[n, :l1] = torch.from_numpy(d)
    return _KNN(dists, idx, None)


class Pointclouds:
    """pytorch3d.structures.Pointclouds as the labeler uses it: padded points / features and the cloud sizes."""

    def __init__(self, points, features=None):
        self._points, self._features = list(points), (list(features) if features is not None else None)

    @staticmethod
    def _pad(ts):
        n = max(len(t) for t in ts)
        out = ts[0].new_zeros((len(ts), n, ts[0].shape[1]))
        for i, t in enumerate(ts):
            out[i, :len(t)] = t
        return out

    def points_padded(self):
        return self._pad(self._points)

    def features_padded(self):
        return self._pad(self._features)

    def num_points_per_cloud(self):
        return torch.tensor([len(t) for t in self._points], dtype=torch.int64)


def calculate_ratio_test(dists):
    """lib/trainer.py:993-1010."""
    dists = (1 - dists).clamp(min=1e-9)
    ratio = dists[:, :, 0:1] / dists[:, :, 1:2]
    return 1 - ratio


def get_topk_matches(dists, idx, num_corres):
    """lib/trainer.py:1012-1016."""
    num_corres = min(num_corres, dists.shape[1])
    dist, idx_source = torch.topk(dists, k=num_corres, dim=1)
    return idx_source, idx.gather(1, idx_source), dist


def match_and_filter_corr(C_batch_0, F_batch_0, C_batch_1, F_batch_1, radius=20, feature_filter='Lowe', spatial_filter='Spherical'):
    """lib/trainer.py:1025-1151 ('None' / 'Spherical' spatial filters) -> (matches [N, 2], uncollated_matches)."""
    num_corres = 5000
    P1, P2 = Pointclouds(C_batch_0, features=F_batch_0), Pointclouds(C_batch_1, features=F_batch_1)
    P1_F, P2_F = P1.features_padded(), P2.features_padded()
    P1_N, P2_N = P1.num_points_per_cloud(), P2.num_points_per_cloud()
    K = 1 if feature_filter == 'None' else 2
    dists_1, idx_1, _ = knn_points(P1_F, P2_F, P1_N, P2_N, K=K)
    dists_2, idx_2, _ = knn_points(P2_F, P1_F, P2_N, P1_N, K=K)
    idx_1, idx_2 = idx_1[:, :, 0:1], idx_2[:, :, 0:1]
    if feature_filter == 'Lowe':
        weights_1 = calculate_ratio_test(1 - 0.5 * dists_1)
        weights_2 = calculate_ratio_test(1 - 0.5 * dists_2)
    else:
        weights_1, weights_2 = dists_1[:, :, 0:1], dists_2[:, :, 0:1]
    n_corres_1, n_corres_2 = min(num_corres, P1_N.min()), min(num_corres, P2_N.min())
    m12_idx1, m12_idx2, _ = get_topk_matches(weights_1, idx_1, n_corres_1)
    m21_idx2, m21_idx1, _ = get_topk_matches(weights_2, idx_2, n_corres_2)
    matches_idx1 = torch.cat((m12_idx1, m21_idx1), dim=1)
    matches_idx2 = torch.cat((m12_idx2, m21_idx2), dim=1)
    bias_1 = torch.cumsum(torch.Tensor([0] + [len(f) for f in F_batch_0][:-1]), 0)
    bias_2 = torch.cumsum(torch.Tensor([0] + [len(f) for f in F_batch_1][:-1]), 0)
    match_1 = torch.cat([m + b for m, b in zip(matches_idx1, bias_1)], dim=0)
    match_2 = torch.cat([m + b for m, b in zip(matches_idx2, bias_2)], dim=0)
    matches = torch.cat([match_1, match_2], dim=1)
    uncollated = []
    for i in range(len(C_batch_0)):
        if spatial_filter == 'None':
            mask = torch.ones_like(torch.norm(C_batch_0[i][matches_idx1[i].squeeze(1)], dim=1)).bool()
        else:
            mask = (torch.norm(C_batch_0[i][matches_idx1[i].squeeze(1)], dim=1) > radius) & \
                   (torch.norm(C_batch_1[i][matches_idx2[i].squeeze(1)], dim=1) > radius)
        uncollated.append(torch.cat([matches_idx1[i][mask], matches_idx2[i][mask]], dim=1))
    return matches, uncollated


def corr_through_registration(input_dict, uncollated_pairs, cfg):
    """lib/trainer.py:1153-1224 (mutual = False branch): SC2-PCR pair by pair, then 3-D nearest neighbours under the pose.
    ``cfg``: oracle SC2Config standing in for ``self.matcher``.  Draws torch.randperm from the global generator per pair."""
    T_ransac, C_batch_0, C_batch_1, fitnesses = [], [], [], []
    for i in range(len(uncollated_pairs)):
        src = input_dict['pcd0'][i][uncollated_pairs[i][:, 0]][None, :, :]
        tgt = input_dict['pcd1'][i][uncollated_pairs[i][:, 1]][None, :, :]
        result, fitness = O.sc2_pcr(src, tgt, cfg)
        result = result[0]
        fitnesses.append(fitness)
        T_ransac.append(result.cpu().float().numpy())
        C_batch_0.append(input_dict['pcd0'][i] @ result[:3, :3].T + result[:3, 3].T)
        C_batch_1.append(input_dict['pcd1'][i])
    P1, P2 = Pointclouds(C_batch_0), Pointclouds(C_batch_1)
    P1_N, P2_N = P1.num_points_per_cloud(), P2.num_points_per_cloud()
    _, idx_1, _ = knn_points(P1.points_padded(), P2.points_padded(), P1_N, P2_N, K=1)
    idx_1 = idx_1[:, :, 0]
    bias_1 = torch.cumsum(torch.Tensor([0] + P1_N.tolist()), 0).long()
    bias_2 = torch.cumsum(torch.Tensor([0] + P2_N.tolist()), 0).long()
    correspondences, uncollated_corr = [], []
    for i, (l1, l2) in enumerate(zip(P1_N, P2_N)):
        pos_sel_1 = torch.randperm(l1)[:min(l1, 5000)]
        pose = torch.Tensor(T_ransac[i])
        src_k = input_dict['pcd0'][i][pos_sel_1.long()] @ pose[:3, :3].T + pose[:3, 3].T
        tgt_k = input_dict['pcd1'][i][idx_1[i][pos_sel_1].long()]
        within = torch.norm(src_k - tgt_k, dim=1) < 2
        pos_sel_1 = pos_sel_1[within]
        pos_sel_2 = idx_1[i][pos_sel_1]
        uncollated_corr.append(torch.cat([pos_sel_1.unsqueeze(1), pos_sel_2.unsqueeze(1)], dim=1))
        pos_sel_1 = pos_sel_1 + bias_1[i]
        pos_sel_2 = pos_sel_2 + bias_2[i]
        correspondences.append(torch.cat([pos_sel_1.unsqueeze(1), pos_sel_2.unsqueeze(1)], dim=1))
    return T_ransac, torch.cat(correspondences, dim=0), [], fitnesses, uncollated_corr
