"""Hot-path helpers of the reference's inference driver scripts/test_kitti.py, same names and semantics:
``find_corr`` (:28-42), ``apply_transform`` (:44-47), ``evaluate_nn_dist`` (:49-52), ``random_sample`` (:54-73)
and the RTE / RRE / success formulas of ``main`` (:188-210).  The draw order on the global numpy RNG is the
reference's.  The dataset / argparse driver around them is a caller, not part of the hot path.
"""
import numpy as np
import torch

from ..lib.eval import find_nn_gpu


def find_corr(xyz0, xyz1, F0, F1, subsample_size=-1):
    """scripts/test_kitti.py:28-42."""
    subsample = len(F0) > subsample_size
    if subsample_size > 0 and subsample:
        N0 = min(len(F0), subsample_size)
        N1 = min(len(F1), subsample_size)
        inds0 = np.random.choice(len(F0), N0, replace=False)
        inds1 = np.random.choice(len(F1), N1, replace=False)
        F0, F1 = F0[torch.from_numpy(inds0).to(F0.device)], F1[torch.from_numpy(inds1).to(F1.device)]
    nn_inds = find_nn_gpu(F0, F1, nn_max_n=500)
    if subsample_size > 0 and subsample:
        return xyz0[inds0], xyz1[inds1[nn_inds]]
    return xyz0, xyz1[nn_inds]


def apply_transform(pts, trans):
    """scripts/test_kitti.py:44-47."""
    R = trans[:3, :3]
    T = trans[:3, 3]
    return pts @ R.t() + T


def evaluate_nn_dist(xyz0, xyz1, T_gth):
    """scripts/test_kitti.py:49-52."""
    xyz0 = apply_transform(xyz0, T_gth)
    dist = np.sqrt(((xyz0 - xyz1) ** 2).sum(1) + 1e-6)
    return dist.tolist()


def random_sample(pcd, feats, N):
    """scripts/test_kitti.py:54-73: exactly N points (permutation if more, with replacement if fewer)."""
    n1 = pcd.size(0) if isinstance(pcd, torch.Tensor) else pcd.shape[0]
    if n1 == N:
        return pcd, feats
    choice = np.random.permutation(n1)[:N] if n1 > N else np.random.choice(n1, N)
    fidx = torch.from_numpy(choice).to(feats.device) if isinstance(feats, torch.Tensor) else choice
    return pcd[choice], feats[fidx]


def rte_rre(T_est, T_gth):
    """scripts/test_kitti.py:188-191 (CPU torch [4,4] in, python floats out), including the diag clamp."""
    rte = np.linalg.norm(T_est[:3, 3] - T_gth[:3, 3])
    trace_matrix = T_est[:3, :3].t() @ T_gth[:3, :3]
    trace_matrix[[0, 1, 2], [0, 1, 2]] = torch.min(torch.ones(3), trace_matrix[[0, 1, 2], [0, 1, 2]])
    rre = np.arccos((np.trace(trace_matrix) - 1) / 2)
    return float(rte), float(rre)


def is_success(rte, rre, rte_thresh=2.0, rre_thresh=5.0):
    """scripts/test_kitti.py:206 (defaults :253-255)."""
    return bool(rte < rte_thresh and not np.isnan(rre) and rre < np.pi / 180 * rre_thresh)
