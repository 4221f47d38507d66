"""Run csrc/gather4_probe.cu on a B200: does cp.async.bulk.tensor.2d tile::gather4 deliver the SWIZZLE_128B operand stage the
tcgen05 descriptors of the sparse convolution expect (row p at p * 128 + ((c ^ (p & 7)) << 4), zeros for row index -1), and at
what rate (cycles per 256-row stage, one issuing warp)?

    python tools/gather4_probe.py
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eyoc_b200 import _C  # noqa: E402


def main():
    lib = _C.lib()
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(0)
    n_rows = int(os.environ.get('G4_ROWS', 3_000_000))
    X = torch.randn(n_rows, 64, generator=g).half().to(dev)
    n_idx = 256 * 4096
    idx = torch.randint(0, n_rows, (n_idx,), generator=g, dtype=torch.int32)
    idx[torch.rand(n_idx, generator=g) < 0.2] = -1
    idx_d = idx.to(dev)
    # configuration = issuing warps x stages in flight (1: 1x1, 102: 1x6, 2: 2x6, 4: 4x6, 8: 8x6, 16: 16x6, 116: 16x1)
    for box_rows in (1, 102, 2, 4, 8, 16, 116):
        for ctas, reps in ((148, 240),):
            out = torch.zeros(256 * 64, dtype=torch.float16, device=dev)
            cyc = torch.zeros(ctas, dtype=torch.int64, device=dev)
            rc = lib.eyoc_debug_gather4_probe(_C.ptr(X), ctypes.c_int64(n_rows), _C.ptr(idx_d), ctypes.c_int(n_idx), ctypes.c_int(ctas),
                                              ctypes.c_int(reps), ctypes.c_int(box_rows), _C.ptr(out), _C.ptr(cyc), _C.stream())
            if rc:
                print(f'box_rows {box_rows}: error {lib.eyoc_last_error().decode()}')
                break
            try:
                torch.cuda.synchronize()
            except RuntimeError as e:
                print(f'config {box_rows} ctas {ctas}: kernel failed: {e}')
                return
            last = ((0 * reps + reps - 1) % (n_idx // 256)) * 256
            rows = idx[last:last + 256].long()
            want = torch.zeros(256, 64, dtype=torch.float16)
            ok = rows >= 0
            want[ok] = X.cpu()[rows[ok]]
            stage = out.cpu().view(256, 8, 8)                      # [row][16-byte chunk slot][8 halves]
            p = torch.arange(256)
            unsw = torch.stack([stage[p, (c ^ (p & 7))] for c in range(8)], 1).reshape(256, 64)     # chunk c sits at slot c ^ (p & 7)
            good = bool(torch.equal(unsw, want))
            plain = bool(torch.equal(stage.reshape(256, 64), want))
            cy = cyc.cpu().numpy()
            print(f'config {box_rows} ctas {ctas} reps {reps}: swizzled layout correct {good} (unswizzled {plain}); '
                  f'{float(np.median(cy)) / reps:.0f} cycles per 256-row stage (median over CTAs)')


if __name__ == '__main__':
    main()
