"""GPU: tensor-core sparse convolution (csrc/sparse_conv_tc.cu, tcgen05 TF32x3) vs the fp32 FMA kernel and an fp64
gather-GEMM restatement, shape by shape, then the whole ResUNetBN2C forward vs the oracle in tf32x3 mode."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore_conv_mode():
    """Tests here flip eyoc_b200.nn.CONV_MODE; put the shipped default back so later test files exercise it."""
    from eyoc_b200 import nn as enn
    saved = enn.CONV_MODE
    yield
    enn.CONV_MODE = saved


def _ref(in0, in1, nbr, W, scale, shift, residual, relu, l2norm):
    x = in0 if in1 is None else torch.cat([in0, in1], 1)
    x = x.double()
    W3 = (W if W.dim() == 3 else W[None]).double()
    n_out = nbr.shape[1] if nbr is not None else x.shape[0]
    out = torch.zeros((n_out, W3.shape[-1]), dtype=torch.float64, device=x.device)
    for k in range(W3.shape[0]):
        if nbr is None:
            out += x @ W3[k]
        else:
            idx = nbr[k].long()
            m = idx >= 0
            out[m] += x[idx[m]] @ W3[k]
    if scale is not None:
        out = out * scale.double() + shift.double()
    elif shift is not None:
        out = out + shift.double()
    if residual is not None:
        out = out + residual.double()
    if relu:
        out = out.clamp(min=0)
    if l2norm:
        out = out / out.norm(dim=1, keepdim=True)
    return out


@pytest.mark.parametrize('c0,c1,cout,K,n_in,n_out,opts', [
    (32, 0, 32, 27, 700, 700, dict(bn=True, relu=True)),
    (32, 0, 64, 27, 900, 300, dict(bn=True)),
    (64, 0, 64, 27, 1500, 1500, dict(bn=True, residual=True, relu=True)),
    (128, 0, 128, 27, 400, 400, dict(bn=True, relu=True)),
    (128, 0, 256, 27, 500, 150, dict(bn=True)),
    (256, 0, 256, 27, 130, 130, dict(bn=True, residual=True, relu=True)),
    (256, 0, 128, 27, 130, 600, dict(bn=True, perm=True)),
    (128, 128, 64, 27, 300, 1000, dict(bn=True, perm=True)),
    (64, 32, 64, 1, 1000, 1000, dict(relu=True)),
    (64, 0, 32, 1, 1000, 1000, dict(bias=True, l2norm=True)),
    (64, 0, 64, 27, 5000, 40000, dict(bn=True, relu=True)),          # many tiles: nacc > 1 accumulators per CTA
])
def test_tc_conv_matches_fp32_and_fp64(c0, c1, cout, K, n_in, n_out, opts):
    from eyoc_b200 import nn as enn
    g = torch.Generator().manual_seed(c0 * 7 + cout + K + n_out)
    dev = 'cuda'
    in0 = torch.randn(n_in, c0, generator=g).to(dev)
    in1 = torch.randn(n_in, c1, generator=g).to(dev) if c1 else None
    cin = c0 + c1
    W = (torch.randn((K, cin, cout) if K > 1 else (cin, cout), generator=g) / np.sqrt(cin * K)).to(dev)
    nbr = None
    if K > 1:
        nbr = torch.randint(0, n_in, (K, n_out), generator=g, dtype=torch.int32)
        nbr[torch.rand(K, n_out, generator=g) < 0.6] = -1
        nbr[13] = torch.randint(0, n_in, (n_out,), generator=g, dtype=torch.int32)      # centre offset always present
        nbr[5] = -1                                                                       # an offset nobody has
        nbr = nbr.to(dev)
    scale = shift = residual = perm = None
    if opts.get('bn'):
        scale = (torch.rand(cout, generator=g) + 0.5).to(dev)
        shift = (torch.randn(cout, generator=g) * 0.1).to(dev)
    if opts.get('bias'):
        shift = (torch.randn(cout, generator=g) * 0.1).to(dev)
    if opts.get('residual'):
        residual = torch.randn(n_out, cout, generator=g).to(dev)
    if opts.get('perm'):
        perm = torch.randperm(n_out, generator=g).to(torch.int32).to(dev)
    relu, l2 = bool(opts.get('relu')), bool(opts.get('l2norm'))
    want = _ref(in0, in1, nbr, W, scale, shift, residual, relu, l2)
    outs = {}
    for mode in ('fp32', 'tf32x3'):
        enn.CONV_MODE = mode
        out = torch.full((n_out, cout), float('nan'), device=dev)
        enn.sparse_conv_raw(in0, in1, nbr, W, scale, shift, residual, relu, l2, out, row_perm=perm)
        torch.cuda.synchronize()
        outs[mode] = out
    scale_ref = float(want.abs().max())
    e32 = float((outs['fp32'].double() - want).abs().max()) / scale_ref
    etc = float((outs['tf32x3'].double() - want).abs().max()) / scale_ref
    assert e32 < 5e-6, e32
    # 3xTF32: products carry ~2^-21 relative error and the TMEM accumulation is not round-to-nearest, so the error
    # grows with the K * cin = up to 6912 accumulated terms; a few times the fp32 FMA kernel.
    assert etc < 5e-5, (etc, e32)


def test_tile_order_and_tiled_conv():
    """eyoc_tile_order: row_perm is a permutation sorted by (cloud group, neighbour mask), nbr_tiled = nbr[:, perm];
    the convolution gives the same rows whether it is handed the natural or the tiled table."""
    from eyoc_b200 import nn as enn
    from eyoc_b200.sparse import CoordinateManager
    from tests.test_resunet_gpu import _cloud
    from eyoc_b200 import sparse as sp
    coords = torch.from_numpy(_cloud(4000, 4, batch=6)).cuda()
    saved_budget = sp.TILE_GROUP_BYTES
    try:
        sp.TILE_GROUP_BYTES = 2.5e6                  # small budget: several cloud groups in this 6-cloud block
        mgr = CoordinateManager(coords)
        nbr = mgr.kernel_map(1, 1, 3)
        tiled, perm = mgr.tiled_map(1, 1, 3)
        n = nbr.shape[1]
        rows_per_cloud = max(1, n // (mgr.max_batch + 1))
        group = max(1, min(mgr.max_batch + 1, int(sp.TILE_GROUP_BYTES // (rows_per_cloud * 4 * 64))))
    finally:
        sp.TILE_GROUP_BYTES = saved_budget
    assert mgr.max_batch == 5 and 1 <= group < 6
    assert torch.equal(torch.sort(perm.long())[0], torch.arange(n, device='cuda'))
    assert torch.equal(tiled, nbr[:, perm.long()])
    # the per-tile offset masks that come out of the same pass == eyoc_tile_masks on the tiled table == a torch restatement
    from eyoc_b200 import _C
    m_pass = mgr.tile_masks(1, 1, 3)
    m_ref = torch.empty_like(m_pass)
    _C.check(_C.lib().eyoc_tile_masks(_C.ptr(tiled), _C.c_int(27), _C.c_int64(n), _C.ptr(m_ref), _C.stream()))
    pad = (-n) % 256
    present = torch.nn.functional.pad((tiled >= 0), (0, pad)).view(27, -1, 256).any(2).long()          # [27, tiles]
    m_torch = sum(present[k] << k for k in range(27)).to(torch.int32)
    assert torch.equal(m_pass, m_ref) and torch.equal(m_pass, m_torch)
    # sort key: (cloud group, neighbour mask with the offsets ordered by frequency - the rarest offset in the top bit,
    # ties by offset index)
    bits = (nbr >= 0).long()                                       # [27, n]
    cnt = bits.sum(1).tolist()
    order = sorted(range(27), key=lambda k: (cnt[k], k))
    pos = {k: 26 - r for r, k in enumerate(order)}
    pmask = sum(bits[k] << pos[k] for k in range(27))
    key = ((coords[:, 0].long() // group) << 27) | pmask
    ks = key[perm.long()]
    assert bool((ks[1:] >= ks[:-1]).all())
    same = ks[1:] == ks[:-1]
    assert bool((perm[1:][same] > perm[:-1][same]).all())          # stable
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, 64, generator=g).cuda()
    W = (torch.randn(27, 64, 64, generator=g) / 40).cuda()
    outs = []
    try:
        enn.CONV_MODE = 'tf32x3'
        for tb, pm, flag in ((nbr, None, False), (nbr, perm, False), (tiled, perm, True)):
            out = torch.full((n, 64), float('nan'), device='cuda')
            enn.sparse_conv_raw(x, None, tb, W, None, None, None, True, False, out, row_perm=pm, nbr_tiled=flag)
            outs.append(out)
    finally:
        pass
    want = _ref(x, None, nbr, W, None, None, None, True, False)
    for o in outs:
        assert float((o.double() - want).abs().max()) < 2e-5 * float(want.abs().max())
    # the accumulation order inside a row does not depend on which tile the row sits in
    assert torch.equal(outs[1], outs[2])


def test_forward_tf32x3_vs_oracle():
    from eyoc_b200 import nn as enn
    from eyoc_b200.model import load_model
    from eyoc_b200.sparse import SparseTensor
    from oracle import resunet_oracle as RO
    from tests.test_resunet_gpu import _cloud
    coords = _cloud(3000, 4, batch=2)
    feats = torch.ones((len(coords), 1), dtype=torch.float32)
    sd = RO.make_state_dict(1, 32, 5, seed=4)
    model = load_model('ResUNetBN2C')(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    want = RO.resunet_forward(coords, feats, sd, True, 5)
    enn.CONV_MODE = 'tf32x3'
    got = model(SparseTensor(feats.cuda(), coordinates=torch.from_numpy(coords).cuda())).F.cpu()
    assert float((got - want).abs().max()) <= 1e-5, float((got - want).abs().max())
