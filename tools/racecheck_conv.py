"""Small driver for `compute-sanitizer --tool racecheck` / `--tool memcheck` on the persistent loop of sparse_conv_h_kernel:
both instantiations (C_out <= 64 and the 128-channel one), a tiled table with precomputed masks, a residual, and a grid cap so
that every CTA walks a dozen tile pairs (ring phases and barrier parities carried across pairs).  Checks the result against
the fp64 gather-GEMM like tests/test_conv_big_gpu.py does, at a size the sanitizer finishes in minutes.

    compute-sanitizer --tool racecheck python tools/racecheck_conv.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eyoc_b200 import _C, nn as enn  # noqa: E402
from eyoc_b200.sparse import xh_pack, xh_unpack  # noqa: E402
from tests.test_conv_tc_gpu import _ref  # noqa: E402


def main():
    lib = _C.lib()
    for c0, cout, n_in, n_out, cap in ((64, 64, 2000, 6144, 2), (128, 256, 1000, 3072, 1)):
        g = torch.Generator().manual_seed(c0 + cout)
        K = 27
        x = torch.randn(n_in, c0, generator=g).cuda()
        W = (torch.randn(K, c0, cout, generator=g) / np.sqrt(c0 * K)).cuda()
        nbr = torch.randint(0, n_in, (K, n_out), generator=g, dtype=torch.int32)
        nbr[torch.rand(K, n_out, generator=g) < 0.6] = -1
        nbr = nbr.cuda()
        res = torch.randn(n_out, cout, generator=g).cuda()
        perm = torch.randperm(n_out, generator=g).to(torch.int32).cuda()
        tiled = nbr[:, perm.long()].contiguous()
        masks = torch.empty((n_out + 255) // 256, dtype=torch.int32, device='cuda')
        _C.check(lib.eyoc_tile_masks(_C.ptr(tiled), _C.c_int(K), _C.c_int64(n_out), _C.ptr(masks), _C.stream()))
        _C.check(lib.eyoc_debug_convh_grid_cap(_C.c_int(cap)))
        out = torch.zeros((n_out, 2 * cout), dtype=torch.float16, device='cuda')
        enn.sparse_conv_h_raw(xh_pack(x), None, tiled, W, None, None, xh_pack(res), True, False, out, row_perm=perm, nbr_tiled=True,
                              tile_masks=masks)
        torch.cuda.synchronize()
        want = _ref(x, None, nbr, W, None, None, res, True, False)
        err = float((xh_unpack(out).double() - want).abs().max()) / float(want.abs().max())
        print(f'cin {c0} cout {cout} n_out {n_out} grid cap {cap}: max rel err {err:.2e}')
        assert err < 4e-5
    _C.check(lib.eyoc_debug_convh_grid_cap(_C.c_int(0)))
    print('racecheck driver finished')


if __name__ == '__main__':
    main()
