"""Drop-in for the reference's lib/metrics.py (pdist, corr_dist) — /root/reference/lib/metrics.py:13-29.

``pdist`` materialises an [n, m] matrix exactly like the reference; it is kept for API
compatibility (lib/trainer.py:461-462).  The hot path (find_nn_gpu) never calls it: the fused
kNN kernel in csrc/knn.cu does not materialise distances.
"""
import torch


def pdist(A, B, dist_type='L2'):
    """lib/metrics.py:22-29."""
    if dist_type == 'L2':
        D2 = torch.sum((A.unsqueeze(1) - B.unsqueeze(0)).pow(2), 2)
        return torch.sqrt(D2 + 1e-7)
    elif dist_type == 'SquareL2':
        return torch.sum((A.unsqueeze(1) - B.unsqueeze(0)).pow(2), 2)
    else:
        raise NotImplementedError('Not implemented')


def corr_dist(est, gth, xyz0, xyz1, weight=None, max_dist=1):
    """lib/metrics.py:13-19."""
    xyz0_est = xyz0 @ est[:3, :3].t() + est[:3, 3]
    xyz0_gth = xyz0 @ gth[:3, :3].t() + gth[:3, 3]
    dists = torch.clamp(torch.sqrt(((xyz0_est - xyz0_gth).pow(2)).sum(1)), max=max_dist)
    if weight is not None:
        dists = weight * dists
    return dists.mean()
