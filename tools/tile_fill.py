"""Fill factor of the tile order (valid map entries / entries of the (tile, offset) items the convolution executes) for
alternative sort keys, on the level-1 and level-2 k = 3 maps of a benchmark block."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eyoc_b200 import synth  # noqa: E402
from eyoc_b200.sparse import CoordinateManager  # noqa: E402


def fill(mask, order, T=256):
    m = mask[order]
    n = m.numel()
    pad = (-n) % T
    m = torch.cat([m, m.new_zeros(pad)]).view(-1, T)
    orr = m[:, 0].clone()
    for j in range(1, T):
        orr |= m[:, j]
    pc = lambda x: sum(((x >> b) & 1) for b in range(27))
    return float(pc(mask).sum()) / float(pc(orr).sum() * T), int(pc(orr).sum())


def main():
    dev = torch.device('cuda', 0)
    pairs = synth.make_pairs(list(range(32)))
    coords_np, _, _, _ = synth.collate_pairs(pairs)
    mgr = CoordinateManager(torch.from_numpy(coords_np).to(dev))
    for ts in (1, 2):
        nbr = mgr.kernel_map(ts, ts, 3)
        mask = ((nbr >= 0).long() << torch.arange(27, device=dev)[:, None]).sum(0)
        n = mask.numel()
        print(f'ts={ts} n={n} distinct masks {int(torch.unique(mask).numel())}')
        keys = {
            'natural order': torch.arange(n, device=dev),
            'mask (current, one group)': torch.sort(mask, stable=True)[1],
        }
        pc = sum(((mask >> b) & 1) for b in range(27))
        keys['popcount, then mask'] = torch.sort((pc << 27) | mask, stable=True)[1]
        # offsets ordered by how often they are set: rarest offsets in the most significant bits
        freq = torch.stack([((mask >> b) & 1).sum() for b in range(27)])
        for name, order in (('rare offsets most significant', torch.argsort(freq)), ('frequent offsets most significant', torch.argsort(freq, descending=True))):
            k = torch.zeros_like(mask)
            for rank, b in enumerate(order.tolist()):
                k |= ((mask >> b) & 1) << (26 - rank)
            keys[name] = torch.sort(k, stable=True)[1]
        # z-plane blocks: the three 9-offset planes as digits (plane patterns cluster)
        gray = mask ^ (mask >> 1)
        keys['gray code of mask'] = torch.sort(gray, stable=True)[1]
        for name, order in keys.items():
            f, items = fill(mask, order)
            print(f'   {name:36s} fill {f:.3f}  (tile, offset) items {items}')


if __name__ == '__main__':
    main()
