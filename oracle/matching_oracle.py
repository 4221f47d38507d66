"""CPU oracle: nearest-neighbour feature matching, find_corr, random_sample, IRLS pose.

TEST INFRASTRUCTURE (see oracle/__init__.py).  torch CPU fp32 restatements; citations are
relative to /root/reference.  ``knn_*_seq`` are the *kernel-order* restatements (sequential
fp32 FMA over the descriptor dimension, the accumulation order the CUDA kernel defines);
they exist so that an index disagreement between the CUDA path and the torch-order oracle
can be classified as an fp32 near-tie (SURVEY.md §7 hard part 2).
"""
import numpy as np
import torch


def pdist_sq(A, B):
    """lib/metrics.py:26-27 ('SquareL2')."""
    return torch.sum((A.unsqueeze(1) - B.unsqueeze(0)).pow(2), 2)


def pdist_l2(A, B):
    """lib/metrics.py:23-25 ('L2')."""
    return torch.sqrt(pdist_sq(A, B) + 1e-7)


def find_nn(F0, F1, nn_max_n=-1, return_distance=False):
    """lib/eval.py:18-48.  Chunking (:20-29) does not change the result; kept for fidelity."""
    if nn_max_n > 1:
        n = len(F0)
        chunks = int(np.ceil(n / nn_max_n))
        dists, inds = [], []
        for i in range(chunks):
            d, ind = pdist_sq(F0[i * nn_max_n:(i + 1) * nn_max_n], F1).min(dim=1)
            dists.append(d.unsqueeze(1))
            inds.append(ind)
        dists, inds = torch.cat(dists), torch.cat(inds)
    else:
        d, inds = pdist_sq(F0, F1).min(dim=1)
        dists = d.unsqueeze(1)
    return (inds, dists) if return_distance else inds


def find_corr(xyz0, xyz1, F0, F1, subsample_size=-1):
    """scripts/test_kitti.py:28-42.  Global numpy RNG, draw order: choice(len F0), choice(len F1)."""
    subsample = len(F0) > subsample_size
    if subsample_size > 0 and subsample:
        n0, n1 = min(len(F0), subsample_size), min(len(F1), subsample_size)
        inds0 = np.random.choice(len(F0), n0, replace=False)
        inds1 = np.random.choice(len(F1), n1, replace=False)
        F0, F1 = F0[inds0], F1[inds1]
    nn_inds = find_nn(F0, F1, nn_max_n=500)
    if subsample_size > 0 and subsample:
        return xyz0[inds0], xyz1[inds1[nn_inds]]
    return xyz0, xyz1[nn_inds]


def random_sample(pcd, feats, n):
    """scripts/test_kitti.py:54-73."""
    n1 = pcd.shape[0]
    if n1 == n:
        return pcd, feats
    choice = np.random.permutation(n1)[:n] if n1 > n else np.random.choice(n1, n)
    return pcd[choice], feats[choice]


def match_argmin(src_desc, tgt_desc):
    """scripts/SC2_PCR/SC2_PCR.py:296-298: argmin_j sqrt(2 - 2 s.t + 1e-6)."""
    return torch.argmin(torch.sqrt(2 - 2 * (src_desc @ tgt_desc.T) + 1e-6), dim=1)


# ------------------------------------------------------------ kernel-order forms
def _fma32(a, b, c):
    """fp32 fused multiply-add emulated in fp64 (24x24-bit product is exact in fp64)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def knn_sq_seq(F0, F1, block=512):
    """argmin_j sum_c (a_c - b_c)^2 with d = a-b rounded, then sequential fma(d, d, acc), c ascending."""
    F0, F1 = np.asarray(F0, np.float32), np.asarray(F1, np.float32)
    idx = np.empty(len(F0), np.int64)
    best = np.empty(len(F0), np.float32)
    for s in range(0, len(F0), block):
        a = F0[s:s + block]
        acc = np.zeros((len(a), len(F1)), np.float32)
        for c in range(F0.shape[1]):
            d = a[:, c:c + 1] - F1[None, :, c]
            acc = _fma32(d, d, acc)
        idx[s:s + block] = acc.argmin(1)
        best[s:s + block] = acc.min(1)
    return idx, best


def knn_cos_seq(S, T, block=512):
    """argmin_j sqrt((2 - 2*dot) + 1e-6) with dot = sequential fma over c ascending."""
    S, T = np.asarray(S, np.float32), np.asarray(T, np.float32)
    idx = np.empty(len(S), np.int64)
    best = np.empty(len(S), np.float32)
    for s in range(0, len(S), block):
        a = S[s:s + block]
        acc = np.zeros((len(a), len(T)), np.float32)
        for c in range(S.shape[1]):
            acc = _fma32(np.broadcast_to(a[:, c:c + 1], acc.shape), np.broadcast_to(T[None, :, c], acc.shape), acc)
        d = np.sqrt((np.float32(2) - np.float32(2) * acc) + np.float32(1e-6))
        idx[s:s + block] = d.argmin(1)
        best[s:s + block] = d.min(1)
    return idx, best


# ------------------------------------------------------------------- IRLS pose
def _rot(axis, x):
    c, s = torch.cos(x), torch.sin(x)
    out = torch.zeros((3, 3))
    i, j = [(1, 2), (2, 0), (0, 1)][axis]
    out[axis, axis] = 1
    out[i, i] = c
    out[j, j] = c
    out[i, j] = -s
    out[j, i] = s
    return out


def irls_pose(pts0, pts1, weight=None):
    """util/transform_estimation.py:89-116 (est_quad_linear_robust) with its helpers :5-86."""
    cur = pts0
    trans = torch.eye(4)
    par = 1.0
    if weight is None:
        weight = torch.ones(pts0.size()[0], 1)
    for i in range(20):
        if i > 0 and i % 5 == 0:
            par /= 2.0
        n = cur.shape[0]
        A0, A1, A2 = torch.zeros((n, 6)), torch.zeros((n, 6)), torch.zeros((n, 6))
        A0[:, 1], A0[:, 2], A0[:, 3] = cur[:, 2], -cur[:, 1], 1
        A1[:, 0], A1[:, 2], A1[:, 4] = -cur[:, 2], cur[:, 0], 1
        A2[:, 0], A2[:, 1], A2[:, 5] = cur[:, 1], -cur[:, 0], 1
        A = weight.repeat(3, 6) * torch.cat((A0, A1, A2), 0)
        b = weight.repeat(3, 1) * torch.cat(
            (pts1[:, 0] - cur[:, 0], pts1[:, 1] - cur[:, 1], pts1[:, 2] - cur[:, 2]), 0).unsqueeze(1)
        x = torch.inverse(A.t().mm(A)).mm(A.t()).mm(b)
        step = torch.eye(4)
        step[:3, :3] = _rot(2, x[2]).mm(_rot(1, x[1])).mm(_rot(0, x[0]))
        step[:3, 3] = x[3:, 0]
        cur = torch.t(step[:3, :3] @ torch.t(cur)) + step[:3, 3]
        weight = par / (torch.norm(cur - pts1, dim=1).unsqueeze(1) + par)
        trans = step.mm(trans)
    return trans
