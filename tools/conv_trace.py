"""Per-item timeline of the split-half tensor-core convolution (eyoc_debug_convh_trace): where a CTA's producer warp 0 and
its MMA thread spend their cycles, item by item.   python tools/conv_trace.py [--flags 0|5|6|7]"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eyoc_b200 import _C, nn as enn, synth  # noqa: E402
from eyoc_b200.sparse import CoordinateManager, xh_pack  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=16)
    ap.add_argument('--flags', type=int, nargs='*', default=[0, 7])
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    pairs = synth.make_pairs(list(range(args.pairs)))
    coords_np, _, _, _ = synth.collate_pairs(pairs)
    mgr = CoordinateManager(torch.from_numpy(coords_np).to(dev))
    lib = _C.lib()
    ts, cin, cout = 1, 64, 64
    nbr, perm = mgr.tiled_map(ts, ts, 3)
    masks = mgr.tile_masks(ts, ts, 3)
    n = nbr.shape[1]
    g = torch.Generator().manual_seed(0)
    x = xh_pack(torch.randn(n, cin, generator=g).to(dev))
    W = (torch.randn(27, cin, cout, generator=g) / 40).to(dev)
    img = enn.split_weights_h(W)
    out = torch.empty(n, cout, device=dev)

    def run():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        enn.sparse_conv_h_raw(x, None, nbr, W, None, None, None, True, False, out, row_perm=perm, nbr_tiled=True,
                              tile_masks=masks, h_img=img)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for flags in args.flags:
        _C.check(lib.eyoc_debug_convh_ablate(flags | 16 | 8))
        t = [run() for _ in range(4)]
        buf = np.zeros((4, 96, 6), np.int64)
        _C.check(lib.eyoc_debug_convh_trace(buf.ctypes.data_as(ctypes.c_void_p)))
        tb = np.zeros((1024, 6), np.int64)
        _C.check(lib.eyoc_debug_convh_times(tb.ctypes.data_as(ctypes.c_void_p)))
        print(f'flags={flags} launches ms: ' + ' '.join(f'{v:.3f}' for v in t))
        for cta in range(2):
            b = buf[cta]
            nit = int(tb[200 + cta, 5])
            t0 = tb[200 + cta, 0]
            print(f' CTA {200 + cta}: items {nit}; phases (prologue, main, drain, epilogue) = {np.diff(tb[200 + cta, :5])}')
            print('  item  P.wait_start  P.wait_cycles  P.issued   M.wait_start  M.wait_cycles  M.issue_cycles')
            for i in range(min(nit, 40)):
                r = b[i]
                print(f'  {i:4d}  {r[0] - t0:12d}  {r[1] - r[0]:13d}  {r[2] - t0:9d}  {r[3] - t0:12d}  {r[4] - r[3]:13d}  {r[5] - r[4]:14d}')
        _C.check(lib.eyoc_debug_convh_ablate(0))


if __name__ == '__main__':
    main()
