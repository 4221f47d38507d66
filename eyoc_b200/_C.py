"""ctypes binding of libeyoc_b200.so (the C-ABI boundary, include/eyoc_b200.h).

There is NO fallback: if the shared library is missing or a CUDA device is not present the
product path raises.  The oracle under ``oracle/`` is never imported from here.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libeyoc_b200.so')

c_void_p, c_int, c_int64, c_size_t, c_float = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                                ctypes.c_size_t, ctypes.c_float)


class SC2Cfg(ctypes.Structure):
    """struct eyoc_sc2_cfg (include/eyoc_b200.h).  c_float rounds Python doubles to fp32 exactly like torch
    rounds a Python scalar that meets an fp32 tensor."""
    _fields_ = [('inlier_threshold', c_float), ('d_thre', c_float), ('d_thre_half', c_float), ('d_thre_sq', c_float),
                ('nms_radius', c_float), ('refine_threshold', c_float), ('num_iterations', c_int), ('k1', c_int),
                ('k2', c_int), ('refine_iterations', c_int)]


class SC2Hooks(ctypes.Structure):
    """struct eyoc_sc2_hooks."""
    _fields_ = [('confidence', c_void_p), ('seeds', c_void_p), ('initial_trans', c_void_p), ('sc2_dense', c_void_p)]


class SC2Layout(ctypes.Structure):
    """struct eyoc_sc2_layout."""
    _fields_ = [(k, c_size_t) for k in ('points', 'hard_bits', 'tight_bits', 'vbuf', 'u', 'confidence', 'scores', 'seeds',
                                        'topk1', 'topk2', 'local_v', 'seed_weights', 'seed_trans', 'counters',
                                        'global_iters', 'local_notclose', 'best_seed', 'refine_counts', 'total',
                                        'csr_rowptr', 'csr_cols', 'csr_vals', 'csr_capacity', 'sort_keys', 'sort_idx',
                                        'sort_offsets', 'sort_temp', 'sort_temp_bytes', 'near_bits', 'status', 'big')] + \
               [(k, c_int) for k in ('words_per_row', 'k1', 'k2', 'num_seeds')]


_lib = None


def lib():
    """Load (once) and return the shared library; fail loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not built: run `python -m eyoc_b200.csrc.build` (nvcc, sm_100a). '
                'eyoc_b200 has no CPU or PyTorch fallback.')
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.eyoc_last_error.restype = ctypes.c_char_p
        _lib.eyoc_version.restype = c_int
        _lib.eyoc_launch_count.restype = ctypes.c_ulonglong
        for name in ('eyoc_knn1_workspace_bytes', 'eyoc_sc2pcr_workspace_bytes', 'eyoc_downsample_workspace_bytes',
                     'eyoc_tile_order_workspace_bytes', 'eyoc_conv_weight_image_floats', 'eyoc_voxelize_workspace_bytes',
                     'eyoc_convh_weight_image_halves', 'eyoc_knn1_tc_workspace_bytes', 'eyoc_pick_seeds_workspace_bytes',
                     'eyoc_power_iteration_workspace_bytes', 'eyoc_irls_workspace_bytes', 'eyoc_knn2_workspace_bytes', 'eyoc_stem_conv_workspace_bytes', 'eyoc_instance_norm_workspace_bytes'):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = c_size_t
    return _lib


def check(code):
    if code != 0:
        raise RuntimeError(f'eyoc_b200 C-ABI error {code}: {lib().eyoc_last_error().decode()}')


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or NULL for None)."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError('eyoc_b200 kernels need CUDA tensors (no CPU fallback)')
    if not t.is_contiguous():
        raise RuntimeError('eyoc_b200 kernels need contiguous tensors')
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('eyoc_b200: tensor is on %s; the B200 path has no CPU fallback' % t.device)


def f32c(t):
    return t.detach().to(torch.float32).contiguous()
