"""Drop-in for the reference's lib/eval.py:18-48 ``find_nn_gpu`` on the fused sm_100a kNN kernel.

Same signature, same return convention (CPU int64 indices [N0], optional CPU fp32 distances
[N0, 1]; ties -> lowest index).  ``nn_max_n`` only controlled the reference's row chunking
(lib/eval.py:20-35), which does not change results; the kernel never materialises the
[N0, N1] matrix so the argument is accepted and ignored.  ``find_nn_cpu`` (cKDTree) is a CPU
fallback in the reference and is deliberately not provided.
"""
import torch

from .. import _C


# 'tc'  : 32-channel descriptors go through the tensor-core pre-filter + exact fp32 re-scoring (eyoc_knn1_tc): the same
#         indices and values as the fp32-FMA kernel, bit for bit
# 'fp32': always the fp32-FMA kernel (eyoc_knn1)
KNN_MODE = 'tc'


def knn1(F0, F1, form=0, return_distance=False):
    """Device-side 1-NN.  F0 [B?, N0, D], F1 [B?, N1, D] CUDA fp32 -> int64 [B?, N0] (and fp32 distances)."""
    _C.require_cuda(F0, F1)
    batched = F0.dim() == 3
    q = _C.f32c(F0 if batched else F0[None])
    r = _C.f32c(F1 if batched else F1[None])
    if q.shape[0] != r.shape[0] or q.shape[2] != r.shape[2]:
        raise RuntimeError(f'knn1: shape mismatch {tuple(F0.shape)} vs {tuple(F1.shape)}')
    B, nq, dim = q.shape
    nr = r.shape[1]
    lib = _C.lib()
    idx = torch.empty((B, nq), dtype=torch.int64, device=q.device)
    if nq == 0:                                  # torch: argmin over dim=1 of a [0, N1] matrix is an empty index vector
        dist0 = torch.empty((B, 0), dtype=torch.float32, device=q.device)
        if not batched:
            idx, dist0 = idx[0], dist0[0]
        return (idx, dist0) if return_distance else idx
    dist = torch.empty((B, nq), dtype=torch.float32, device=q.device) if return_distance else None
    use_tc = KNN_MODE == 'tc' and nr > 0 and bool(lib.eyoc_knn1_tc_supported(_C.c_int(dim)))
    if use_tc:
        ws_bytes = lib.eyoc_knn1_tc_workspace_bytes(_C.c_int(B), _C.c_int64(nq), _C.c_int64(nr))
    else:
        ws_bytes = lib.eyoc_knn1_workspace_bytes(_C.c_int(B), _C.c_int64(nq))
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=q.device)
    fn = lib.eyoc_knn1_tc if use_tc else lib.eyoc_knn1
    with torch.cuda.device(q.device):
        _C.check(fn(_C.ptr(q), _C.ptr(r), _C.c_int(B), _C.c_int64(nq), _C.c_int64(nr), _C.c_int(dim),
                    _C.c_int(form), _C.ptr(ws), _C.c_size_t(ws.numel()), _C.ptr(idx), _C.ptr(dist),
                    _C.stream()))
    if not batched:
        idx = idx[0]
        dist = dist[0] if dist is not None else None
    return (idx, dist) if return_distance else idx


def find_nn_gpu(F0, F1, nn_max_n=-1, return_distance=False, dist_type='SquareL2'):
    """lib/eval.py:18-48.  Returns CPU tensors like the reference (:28-29,37-48)."""
    if dist_type != 'SquareL2':
        raise NotImplementedError('find_nn_gpu: only SquareL2 is on the hot path (lib/eval.py:18 default)')
    out = knn1(F0, F1, form=0, return_distance=return_distance)
    if return_distance:
        inds, dists = out
        return inds.cpu(), dists.cpu().unsqueeze(1)
    return out.cpu()
