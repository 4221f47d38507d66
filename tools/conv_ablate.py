"""Time one tensor-core sparse convolution of the benchmark network with parts of the kernel switched off
(eyoc_debug_conv_ablate): which of gather / x_lo pass, weight slabs, MMAs bounds the launch.

    python tools/conv_ablate.py [--pairs 32]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eyoc_b200 import _C, nn as enn, synth  # noqa: E402
from eyoc_b200.sparse import CoordinateManager  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=32)
    ap.add_argument('--mode', default='f16x3', choices=['f16x3', 'tf32x3'])
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    pairs = synth.make_pairs(list(range(args.pairs)))
    coords_np, _, _, _ = synth.collate_pairs(pairs)
    mgr = CoordinateManager(torch.from_numpy(coords_np).to(dev))
    lib = _C.lib()
    for (ts, cin, cout) in ((1, 64, 64), (1, 32, 32), (4, 128, 128)):
        nbr, perm = mgr.tiled_map(ts, ts, 3)
        n = nbr.shape[1]
        g = torch.Generator().manual_seed(0)
        x = torch.randn(n, cin, generator=g).to(dev)
        W = (torch.randn(27, cin, cout, generator=g) / 40).to(dev)
        h = args.mode == 'f16x3'
        enn.CONV_MODE = args.mode
        img = enn.split_weights_h(W) if h else enn.split_weights(W)
        out = torch.empty(n, cout, device=dev)
        masks = mgr.tile_masks(ts, ts, 3) if h else None
        if h:
            from eyoc_b200.sparse import xh_pack
            x = xh_pack(x)
        set_ablate = lib.eyoc_debug_convh_ablate if h else lib.eyoc_debug_conv_ablate
        get_times = lib.eyoc_debug_convh_times if h else lib.eyoc_debug_conv_times

        def run():
            if h:
                enn.sparse_conv_h_raw(x, None, nbr, W, None, None, None, True, False, out, row_perm=perm, nbr_tiled=True,
                                      tile_masks=masks, h_img=img)
            else:
                enn.sparse_conv_raw(x, None, nbr, W, None, None, None, True, False, out, row_perm=perm, nbr_tiled=True, wt_img=img)
        M = int((nbr >= 0).sum())
        variants = ((0, 'full'), (6, 'MMA only'), (5, 'gather only'), (7, 'skeleton'))
        if h:
            variants += ((32768, 'full (DBG build)'), (32768 | 256, 'DBG, no probes'), (32768 | 512, 'DBG, no proxy fence'), (32768 | 768, 'DBG, neither'))
        for flags, name in variants:
            _C.check(set_ablate(flags))
            ts_ = []
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run()
                e1.record()
                torch.cuda.synchronize()
                ts_.append(e0.elapsed_time(e1))
            print(f'ts={ts} {cin:3d}->{cout:3d} n={n:8d} M={M:9d} {name:16s} {min(ts_[1:]):8.3f} ms', flush=True)
        # per-CTA phase times of the full kernel
        import ctypes
        import numpy as np
        _C.check(set_ablate(8))
        run()
        torch.cuda.synchronize()
        buf = np.zeros((1024, 6), np.int64)
        _C.check(get_times(buf.ctypes.data_as(ctypes.c_void_p)))
        ncta = min(1024, (n + 511) // 512)
        b = buf[:ncta]
        d = np.diff(b[:, :5], axis=1)
        print('   CTA phases (cycles, median over %d CTAs): prologue %d  main loop %d  drain %d  epilogue %d  items %d  -> %.0f cycles/item'
              % (ncta, *np.median(d, axis=0), np.median(b[:, 5]), np.median(d[:, 1] / np.maximum(b[:, 5], 1))), flush=True)
        _C.check(set_ablate(0))


if __name__ == '__main__':
    main()
