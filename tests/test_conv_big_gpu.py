"""GPU: the PERSISTENT loop of sparse_conv_h_kernel at the sizes bench.py runs it.

A CTA of the persistent grid walks tile pair after tile pair; its stage ring, weight-slab ring, barrier phases and the
`done` barrier's parity carry over (`ibase` / `wbase` / `iter` in csrc/sparse_conv_h.cu).  Small inputs never take a second
trip through that loop, so these tests (a) run single launches with >= 300 k output rows (4-20 tile pairs per CTA; the
debug grid cap raises that to dozens) against the fp64 gather-GEMM restatement, for the C_out <= 64 and the 128-channel
(WIDE) instantiations, with a residual and a tiled table + precomputed tile masks, and (b) push 14 KITTI-sized clouds
(~380 k voxels) through the whole ResUNetBN2C forward in one batch and require every cloud's rows to equal that cloud's
single-cloud forward BIT FOR BIT, two of them within 1e-5 of the oracle (model/resunet.py:142-193 semantics), and the
grid-capped run (every launch of every level loops) to be bit-identical too."""
import numpy as np
import pytest
import torch

from tests.test_conv_tc_gpu import _ref

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _uncap():
    from eyoc_b200 import _C
    yield
    _C.lib().eyoc_debug_convh_grid_cap(_C.c_int(0))


@pytest.mark.parametrize('c0,c1,cout,n_in,n_out,cap', [
    (64, 0, 64, 150_000, 520_000, 0),        # 1016 tile pairs on 148 CTAs: ~7 trips per CTA
    (64, 0, 64, 150_000, 520_000, 24),       # ~42 trips per CTA
    (32, 0, 32, 100_000, 400_000, 19),
    (128, 0, 128, 60_000, 300_000, 0),       # WIDE: 586 pairs on 148 CTAs
    (128, 0, 256, 60_000, 300_000, 0),       # WIDE, two 128-channel parts: 74 CTAs per part, ~8 trips
    (256, 0, 256, 40_000, 160_000, 9),       # WIDE, 8 chunks, ~35 trips
    (64, 64, 64, 100_000, 310_000, 0),       # fused concat
])
def test_h_conv_persistent_loop_matches_fp64(c0, c1, cout, n_in, n_out, cap):
    from eyoc_b200 import _C, nn as enn
    from eyoc_b200.sparse import xh_pack, xh_unpack
    K = 27
    g = torch.Generator().manual_seed(c0 + 3 * cout + n_out + cap)
    dev = 'cuda'
    in0 = torch.randn(n_in, c0, generator=g).to(dev)
    in1 = torch.randn(n_in, c1, generator=g).to(dev) if c1 else None
    cin = c0 + c1
    W = (torch.randn((K, cin, cout), generator=g) / np.sqrt(cin * K)).to(dev)
    # neighbour tables with structure: the presence of an offset varies from tile to tile (so work lists differ between the
    # tile pairs one CTA walks), some tiles have no neighbour at all for an offset, one whole stretch of rows has none
    nbr = torch.randint(0, n_in, (K, n_out), generator=g, dtype=torch.int32)
    tile = torch.arange(n_out) // 256
    for k in range(K):
        dens = 0.15 + 0.7 * ((tile * (k + 3)) % 7 == 0).float() + 0.3 * ((tile + k) % 3 == 0).float()
        drop = torch.rand(n_out, generator=g) >= dens.clamp(max=0.95)
        nbr[k, drop] = -1
        nbr[k, ((tile + 2 * k) % 5 == 0)] = -1
    nbr[:, 70_000:71_500] = -1
    nbr = nbr.to(dev)
    scale = (torch.rand(cout, generator=g) + 0.5).to(dev)
    shift = (torch.randn(cout, generator=g) * 0.1).to(dev)
    residual = torch.randn(n_out, cout, generator=g).to(dev)
    perm = torch.randperm(n_out, generator=g).to(torch.int32).to(dev)
    want = _ref(in0, in1, nbr, W, scale, shift, residual, True, False)
    ref_scale = float(want.abs().max())
    h0, h1 = xh_pack(in0), (xh_pack(in1) if in1 is not None else None)
    res_h = xh_pack(residual)
    lib = _C.lib()
    # tiled table in `perm` order + the per-tile masks the kernel is normally handed
    tiled = nbr[:, perm.long()].contiguous()
    masks = torch.empty((n_out + 255) // 256, dtype=torch.int32, device=dev)
    _C.check(lib.eyoc_tile_masks(_C.ptr(tiled), _C.c_int(K), _C.c_int64(n_out), _C.ptr(masks), _C.stream()))
    _C.check(lib.eyoc_debug_convh_grid_cap(_C.c_int(cap)))
    outs = []
    for kw in (dict(row_perm=perm, nbr_tiled=True, tile_masks=masks), dict(row_perm=perm, nbr_tiled=True), dict()):
        out = torch.full((n_out, 2 * cout), float('nan'), dtype=torch.float16, device=dev)
        enn.sparse_conv_h_raw(h0, h1, tiled if kw else nbr, W, scale, shift, res_h, True, False, out, **kw)
        outs.append(xh_unpack(out))
    torch.cuda.synchronize()
    e = float((outs[0].double() - want).abs().max()) / ref_scale
    assert e < 4e-5, e
    # natural order, tiled order, tiled order with each CTA deriving its masks: the same bits
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    # a second launch over the same memory (fresh counters, warm rings): the same bits again
    out2 = torch.full((n_out, 2 * cout), float('nan'), dtype=torch.float16, device=dev)
    enn.sparse_conv_h_raw(h0, h1, tiled, W, scale, shift, res_h, True, False, out2, row_perm=perm, nbr_tiled=True, tile_masks=masks)
    assert torch.equal(xh_unpack(out2), outs[0])


def _model(sd):
    from eyoc_b200.model import load_model
    model = load_model('ResUNetBN2C')(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3)
    model.load_state_dict(sd)
    return model.cuda().eval()


def test_forward_bench_scale_batched_equals_single_and_oracle():
    """14 KITTI-sized clouds in one batch (>= 350 k level-1 rows: every level-1 launch makes >= 4 trips per CTA) ==
    14 single-cloud forwards, bit for bit; clouds 0 and 9 within 1e-5 of the oracle."""
    from eyoc_b200 import _C, nn as enn, synth
    from eyoc_b200.sparse import SparseTensor
    from oracle import resunet_oracle as RO
    assert enn.CONV_MODE == 'f16x3' and enn.TILE_ORDER
    pairs = synth.make_pairs(list(range(7)))
    clouds = [c for p in pairs for c in (p['coords0'], p['coords1'])]
    coords = synth.collate(clouds)
    n = len(coords)
    assert n >= 350_000, n
    offs = np.concatenate([[0], np.cumsum([len(c) for c in clouds])])
    sd = RO.make_state_dict(1, 32, 5, seed=3)
    model = _model(sd)
    dev = torch.device('cuda')

    def fwd(c):
        return model(SparseTensor(torch.ones(len(c), 1, device=dev), coordinates=torch.from_numpy(c).to(dev))).F

    F = fwd(coords)
    assert bool(torch.isfinite(F).all())
    for b, c in enumerate(clouds):
        Fb = fwd(synth.collate([c]))
        assert torch.equal(F[offs[b]:offs[b + 1]], Fb), f'cloud {b}: batched rows differ from the single-cloud forward'
    for b in (0, 9):
        cb = synth.collate([clouds[b]])
        want = RO.resunet_forward(cb, torch.ones(len(cb), 1), sd, True, 5)
        err = float((F[offs[b]:offs[b + 1]].cpu() - want).abs().max())
        assert err <= 1e-5, (b, err)
    # every launch of every level (the 128 / 256-channel ones too) walks dozens of tile pairs per CTA: same bits
    _C.check(_C.lib().eyoc_debug_convh_grid_cap(_C.c_int(11)))
    F_cap = fwd(coords)
    _C.check(_C.lib().eyoc_debug_convh_grid_cap(_C.c_int(0)))
    assert torch.equal(F, F_cap)
    # and the run is reproducible launch to launch
    assert torch.equal(F, fwd(coords))
