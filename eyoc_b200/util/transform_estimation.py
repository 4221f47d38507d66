"""Drop-in for the reference's util/transform_estimation.py on the device: ``est_quad_linear_robust`` (:89-116, the
20-round robust small-angle 6-DoF solve used by the trainers' validation, lib/trainer.py:360,1783) and ``pose_estimation``
(:119-144).  The reference evaluates the solve on CPU tensors; here one CUDA kernel runs all rounds (csrc/irls.cu,
eyoc_irls_pose).  Inputs may be CPU or CUDA tensors - they are moved to the current CUDA device (there is no CPU
implementation: without a GPU this raises) - and the 4 x 4 result comes back on the device of ``pts0`` like the reference's.
"""
import torch

from .. import _C
from ..lib.eval import knn1
from ..sparse import SparseTensor


def est_quad_linear_robust(pts0, pts1, weight=None, iterations=20):
    """util/transform_estimation.py:89-116.  pts0, pts1 [n, 3]; weight [n, 1] or None -> trans [4, 4] fp32."""
    if not torch.cuda.is_available():
        raise RuntimeError('eyoc_b200: est_quad_linear_robust needs a CUDA device (no CPU fallback)')
    home = pts0.device
    dev = home if home.type == 'cuda' else torch.device('cuda', torch.cuda.current_device())
    p0 = pts0.detach().to(dev, torch.float32).contiguous()
    p1 = pts1.detach().to(dev, torch.float32).contiguous()
    if p0.dim() != 2 or p0.shape[1] != 3 or p1.shape != p0.shape or p0.shape[0] == 0:
        raise RuntimeError(f'est_quad_linear_robust: pts0 / pts1 must be matching [n, 3] tensors, got {tuple(pts0.shape)} / {tuple(pts1.shape)}')
    w = None
    if weight is not None:
        w = weight.detach().to(dev, torch.float32).reshape(-1).contiguous()
        if w.numel() != p0.shape[0]:
            raise RuntimeError('est_quad_linear_robust: one weight per point')
    n = p0.shape[0]
    lib = _C.lib()
    lib.eyoc_irls_workspace_bytes.restype = _C.c_size_t
    ws = torch.empty(max(lib.eyoc_irls_workspace_bytes(_C.c_int64(n)), 8), dtype=torch.uint8, device=dev)
    out = torch.empty((4, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _C.check(lib.eyoc_irls_pose(_C.ptr(p0), _C.ptr(p1), _C.ptr(w), _C.c_int64(n), _C.c_int(int(iterations)), _C.ptr(out),
                                    _C.ptr(ws), _C.c_size_t(ws.numel()), _C.stream()))
    return out.to(home)


def pose_estimation(model, device, xyz0, xyz1, coord0, coord1, feats0, feats1, return_corr=False):
    """util/transform_estimation.py:119-144: features of both clouds, per-point best match by inner product (the reference
    materialises the full N0 x N1 product and takes ``max(dim=1)``; for the L2-normalised descriptors of the model the arg-max
    of the inner product is the arg-min of sqrt(2 - 2 s + 1e-6), which the fused 1-NN kernel returns without the matrix),
    then the robust pose with the inner products as weights.  ``return_corr`` returns the matched indices in place of the
    dense matrix."""
    F0 = model(SparseTensor(feats0.to(device), coordinates=coord0.to(device))).F
    F1 = model(SparseTensor(feats1.to(device), coordinates=coord1.to(device))).F
    inds = knn1(F0, F1, form=1)
    weight = (F0 * F1[inds]).sum(1, keepdim=True)
    xyz1_corr = xyz1.to(device)[inds, :]
    trans = est_quad_linear_robust(xyz0.to(device), xyz1_corr, weight)
    if return_corr:
        return trans, weight.cpu(), inds
    return trans, weight.cpu()
