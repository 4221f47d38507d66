"""Time the coordinate-manager steps of one benchmark block (64 pairs) with CUDA events."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eyoc_b200 import synth  # noqa: E402
from eyoc_b200.sparse import CoordinateManager  # noqa: E402


def timed(name, fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f'{name:40s} {e0.elapsed_time(e1):8.3f} ms', flush=True)
    return out


def main():
    dev = torch.device('cuda', 0)
    pairs = synth.make_pairs(list(range(64)))
    coords_np, _, _, _ = synth.collate_pairs(pairs)
    coords = torch.from_numpy(coords_np).to(dev)
    for rep in range(2):
        print(f'--- pass {rep}')
        mgr = timed('hash build (level 1)', lambda: CoordinateManager(coords))
        timed('levels 2, 4, 8', lambda: mgr.ensure_levels(8))
        w = torch.randn(125, 32, device=dev)
        ones = torch.ones(coords.shape[0], device=dev)
        for packed in (False, True):
            mgr._maps.pop((1, 1, 3, False), None)
            timed(f'stem conv k5 (+ k3 table) packed={packed}', lambda: mgr.stem_conv(ones, w, None, None, True, 5, packed))
        mgr._maps.pop((1, 1, 3, False), None)
        timed('stem conv k5 (+ k3 table) all-ones input', lambda: mgr.stem_conv(None, w, None, None, True, 5, True))
        mgr._maps.pop((1, 1, 3, False), None)
        timed('map k5 s1 (self)', lambda: mgr.kernel_map(1, 1, 5))
        mgr._maps.pop((1, 1, 5, False), None)
        for ts in (1, 2, 4, 8):
            timed(f'map k3 ts{ts} (self)', lambda: mgr.kernel_map(ts, ts, 3))
        for ts in (1, 2, 4):
            timed(f'map k3 ts{ts}->ts{2 * ts} (strided)', lambda: mgr.kernel_map(ts, 2 * ts, 3))
            timed(f'map k3 ts{2 * ts}->ts{ts} (transposed)', lambda: mgr.kernel_map(2 * ts, ts, 3, True))
        for ts in (1, 2, 4, 8):
            timed(f'tile order + masks k3 ts{ts}', lambda: (mgr.tiled_map(ts, ts, 3), mgr.tile_masks(ts, ts, 3)))
        for ts in (1, 2, 4):
            timed(f'tile order strided ts{ts}', lambda: (mgr.tiled_map(ts, 2 * ts, 3), mgr.tile_masks(ts, 2 * ts, 3)))
            timed(f'tile order transposed ts{2 * ts}', lambda: (mgr.tiled_map(2 * ts, ts, 3, True), mgr.tile_masks(2 * ts, ts, 3, True)))


if __name__ == '__main__':
    main()
