"""Drop-in for scripts/SC2_PCR/common.py:7-45 ``rigid_transform_3d`` (weighted Kabsch).

The reference builds a dense [bs,n,n] diag matrix and round-trips H through the CPU for LAPACK's SVD
(common.py:33-37).  Here one CTA per batch item accumulates the weighted moments in fp64 and solves
the 3x3 SVD on device by one-sided Jacobi (csrc/sc2pcr.cu kabsch_kernel).  Like the reference
(common.py:20) negative weights are zeroed IN PLACE on the caller's tensor.
"""
import torch

from ... import _C


def rigid_transform_3d(A, B, weights=None, weight_threshold=0):
    _C.require_cuda(A, B, weights)
    bs, n = A.shape[0], A.shape[1]
    a, b = _C.f32c(A), _C.f32c(B)
    w = None
    if weights is not None:
        w = weights if (weights.dtype == torch.float32 and weights.is_contiguous()) else _C.f32c(weights)
    T = torch.empty((bs, 4, 4), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        _C.check(_C.lib().eyoc_kabsch_batched(_C.ptr(a), _C.ptr(b), _C.ptr(w), _C.c_int(bs), _C.c_int(n),
                                              _C.c_float(weight_threshold), _C.ptr(T), _C.stream()))
    if weights is not None and w is not weights:
        weights.copy_(w)
    return T
