"""SE(3) helpers with the reference's scripts/SC2_PCR/utils/SE3.py interface (transform :43-57,
integrate_trans :73-96, decompose_trans :59-71).  Host-side glue on [bs,4,4] tensors; the fused
estimator kernels apply transforms on device themselves (csrc/sc2pcr.cu apply_T)."""
import numpy as np
import torch


def transform(pts, trans):
    """R @ p + t for pts [bs,n,3] / [n,3] and trans [bs,4,4] / [4,4] (SE3.py:43-57)."""
    if len(pts.shape) == 3:
        out = trans[:, :3, :3] @ pts.permute(0, 2, 1) + trans[:, :3, 3:4]
        return out.permute(0, 2, 1)
    out = trans[:3, :3] @ pts.T + trans[:3, 3:4]
    return out.T


def decompose_trans(trans):
    """SE3.py:59-71."""
    if len(trans.shape) == 3:
        return trans[:, :3, :3], trans[:, :3, 3:4]
    return trans[:3, :3], trans[:3, 3:4]


def integrate_trans(R, t):
    """Pack R [bs,3,3] (or [3,3]) and t into 4x4 (SE3.py:73-96)."""
    if len(R.shape) == 3:
        if isinstance(R, torch.Tensor):
            trans = torch.eye(4)[None].repeat(R.shape[0], 1, 1).to(R.device)
        else:
            trans = np.eye(4)[None]
        trans[:, :3, :3] = R
        trans[:, :3, 3:4] = t.view([-1, 3, 1])
    else:
        trans = torch.eye(4).to(R.device) if isinstance(R, torch.Tensor) else np.eye(4)
        trans[:3, :3] = R
        trans[:3, 3:4] = t
    return trans
