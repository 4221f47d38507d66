"""GPU: fp16 hi/lo split tensor-core sparse convolution (csrc/sparse_conv_h.cu, 'f16x3') - the split-half activation
format, the kernel shape by shape against an fp64 gather-GEMM restatement (and the fp32 FMA kernel beside it), tiled
tables with precomputed tile masks, and the whole ResUNetBN2C forward against the oracle."""
import numpy as np
import pytest
import torch

from tests.test_conv_tc_gpu import _ref

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _restore_conv_mode():
    from eyoc_b200 import nn as enn
    saved = enn.CONV_MODE
    yield
    enn.CONV_MODE = saved


def test_split_half_format_round_trip():
    """x = hi + lo' 2^-11 with hi = fp16(x), lo' = fp16((x - hi) 2^11): layout [n, c/32, (hi | lo'), 32] halves,
    22 significant bits, exact for fp16-representable values and for zeros."""
    from eyoc_b200.sparse import xh_pack, xh_unpack
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(1000, 96, generator=g) * torch.logspace(-4, 3, 96)[None]).cuda()
    x[0] = 0.0
    x[1] = torch.randn(96, generator=g).half().float().cuda()
    xh = xh_pack(x)
    assert xh.shape == (1000, 192) and xh.dtype == torch.float16
    v = xh.view(1000, 3, 2, 32).float()
    hi, lo = v[:, :, 0].reshape(1000, 96), v[:, :, 1].reshape(1000, 96)
    assert torch.equal(hi, x.half().float())
    assert torch.equal(lo, ((x - hi) * 2048).half().float())
    back = xh_unpack(xh)
    assert torch.equal(back[:2], x[:2])
    err = (back - x).abs()[2:]
    # 22 significant bits while lo' is a normal fp16 number, an absolute floor of 2^-36 (half a subnormal step / 2^11) below
    assert bool((err <= x.abs()[2:] * 2.0 ** -22 + 2.0 ** -35).all()), float((err / x.abs()[2:].clamp(min=1e-30)).max())


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
    (64, 0, 64, 27, 5000, 40000, dict(bn=True, relu=True)),          # many CTAs
    (128, 256, 128, 27, 300, 700, dict(bn=True, relu=True)),         # 12 chunks (ResUNetFatBN's widest concat)
])
def test_h_conv_matches_fp64(c0, c1, cout, K, n_in, n_out, opts):
    from eyoc_b200 import nn as enn
    from eyoc_b200.sparse import xh_pack, xh_unpack
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
        nbr[13] = torch.randint(0, n_in, (n_out,), generator=g, dtype=torch.int32)
        nbr[5] = -1
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
    scale_ref = float(want.abs().max())
    h0, h1 = xh_pack(in0), (xh_pack(in1) if in1 is not None else None)
    # fp32 output rows, fp32 residual
    out = torch.full((n_out, cout), float('nan'), device=dev)
    enn.sparse_conv_h_raw(h0, h1, nbr, W, scale, shift, residual, relu, l2, out, row_perm=perm)
    torch.cuda.synchronize()
    e = float((out.double() - want).abs().max()) / scale_ref
    # operands carry 22 bits and the TMEM accumulation is not round-to-nearest: the error grows with the K * cin (up to
    # 10 368) accumulated terms, as for the tf32x3 kernel (tests/test_conv_tc_gpu.py allows 5e-5)
    tol = 1e-5 if K * cin <= 2000 else 4e-5
    assert e < tol, e
    if not l2:
        # split-half output rows, split-half residual
        outh = torch.zeros((n_out, 2 * cout), dtype=torch.float16, device=dev)
        enn.sparse_conv_h_raw(h0, h1, nbr, W, scale, shift, xh_pack(residual) if residual is not None else None, relu, l2,
                              outh, row_perm=perm)
        torch.cuda.synchronize()
        eh = float((xh_unpack(outh).double() - want).abs().max()) / scale_ref
        assert eh < tol, eh


def test_h_conv_tiled_masks_and_order_independence():
    """The same rows come out whether the kernel is handed the natural table, the tiled one (each CTA deriving its tile
    masks), or the tiled one with eyoc_tile_masks - bit for bit: the accumulation order inside a row does not depend on
    the tile it sits in."""
    from eyoc_b200 import nn as enn
    from eyoc_b200.sparse import CoordinateManager, xh_pack
    from tests.test_resunet_gpu import _cloud
    coords = torch.from_numpy(_cloud(4000, 4, batch=6)).cuda()
    mgr = CoordinateManager(coords)
    nbr = mgr.kernel_map(1, 1, 3)
    tiled, perm = mgr.tiled_map(1, 1, 3)
    masks = mgr.tile_masks(1, 1, 3)
    n = nbr.shape[1]
    bits = ((tiled >= 0).long() << torch.arange(27, device='cuda')[:, None]).sum(0)
    pad = (-n) % 256
    want_masks = torch.cat([bits, bits.new_zeros(pad)]).view(-1, 256)
    acc = want_masks[:, 0].clone()
    for j in range(1, 256):
        acc |= want_masks[:, j]
    assert torch.equal(masks.long(), acc)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, 64, generator=g).cuda()
    W = (torch.randn(27, 64, 64, generator=g) / 40).cuda()
    xh = xh_pack(x)
    outs = []
    for tb, pm, flag, mk in ((nbr, None, False, None), (nbr, perm, False, None), (tiled, perm, True, None), (tiled, perm, True, masks)):
        out = torch.full((n, 64), float('nan'), device='cuda')
        enn.sparse_conv_h_raw(xh, None, tb, W, None, None, None, True, False, out, row_perm=pm, nbr_tiled=flag, tile_masks=mk)
        outs.append(out)
    want = _ref(x, None, nbr, W, None, None, None, True, False)
    for o in outs:
        assert float((o.double() - want).abs().max()) < 5e-6 * float(want.abs().max())
    assert torch.equal(outs[1], outs[2]) and torch.equal(outs[2], outs[3])


def test_forward_f16x3_vs_oracle():
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
    assert enn.CONV_MODE == 'f16x3'                              # the shipped default
    got = model(SparseTensor(feats.cuda(), coordinates=torch.from_numpy(coords).cuda())).F.cpu()
    assert float((got - want).abs().max()) <= 1e-5, float((got - want).abs().max())
    # un-normalised output: the last convolution then writes split-half rows and .F converts them
    model.normalize_feature = False
    raw = model(SparseTensor(feats.cuda(), coordinates=torch.from_numpy(coords).cuda()))
    assert raw._F is None and raw._Fh is not None
    want_raw = RO.resunet_forward(coords, feats, sd, False, 5)
    assert float((raw.F.cpu() - want_raw).abs().max()) <= 1e-5 * max(1.0, float(want_raw.abs().max()))


def test_forward_fat_variant_vs_oracle():
    """ResUNetFatBN (model/resunet.py:224-227, the model of scripts/test_waymo.sh:13): TR_CHANNELS [_, 128, 128, 128, 256],
    i.e. a 12-chunk 384-channel concat and 256-channel transposed convolutions - same kernels, different channel table."""
    from eyoc_b200.model import load_model
    from eyoc_b200.sparse import SparseTensor
    from oracle import resunet_oracle as RO
    from tests.test_resunet_gpu import _cloud
    coords = _cloud(2500, 4, batch=2)
    feats = torch.ones((len(coords), 1), dtype=torch.float32)
    torch.manual_seed(9)
    model = load_model('ResUNetFatBN')(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3)
    g = torch.Generator().manual_seed(10)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    assert sd['conv3_tr.kernel'].shape == (27, 384, 128) and sd['conv4_tr.kernel'].shape == (27, 256, 256)
    want = RO.resunet_forward(coords, feats, sd, True, 5)
    model = model.cuda().eval()
    got = model(SparseTensor(feats.cuda(), coordinates=torch.from_numpy(coords).cuda())).F.cpu()
    assert float((got - want).abs().max()) <= 1e-5, float((got - want).abs().max())


def test_range_guard_falls_back_to_fp32_activations():
    """A checkpoint whose activations leave the fp16 hi/lo range (|x| >= 65504): the kernels flag it on the device and
    forward() re-runs through the fp32-activation path - finite features within tolerance of the fp32 oracle, where the
    split-half path alone would have produced Inf / NaN.  An in-range checkpoint leaves the flag clear."""
    import warnings
    from eyoc_b200 import nn as enn
    from eyoc_b200.model import load_model
    from eyoc_b200.sparse import SparseTensor
    from oracle import resunet_oracle as RO
    from tests.test_resunet_gpu import _cloud
    coords = _cloud(1500, 6, batch=2)
    feats = torch.ones((len(coords), 1), dtype=torch.float32)
    sd = RO.make_state_dict(1, 32, 5, seed=8)
    sd['conv1.kernel'] = sd['conv1.kernel'] * 3.0e6
    model = load_model('ResUNetBN2C')(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    want = RO.resunet_forward(coords, feats, sd, True, 5)
    assert bool(torch.isfinite(want).all())
    x = SparseTensor(feats.cuda(), coordinates=torch.from_numpy(coords).cuda())
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        got = model(x).F.cpu()
    assert any('fp16 hi/lo range' in str(m.message) for m in w), [str(m.message) for m in w]
    assert enn.CONV_MODE == 'f16x3'                                   # restored
    assert bool(torch.isfinite(got).all()), 'fallback output not finite'
    assert float((got - want).abs().max()) <= 5e-5, ('fallback vs oracle', float((got - want).abs().max()))
    # without the guard the same forward is SILENTLY wrong (Inf halves, NaNs scrubbed to 0 by the ReLUs): what the flag is for
    model.RANGE_CHECK = False
    raw = model(SparseTensor(feats.cuda(), coordinates=torch.from_numpy(coords).cuda())).F.cpu()
    assert not (float((raw - want).abs().max()) <= 1e-2), 'the unguarded split-half forward was expected to be wrong'
    # in-range weights: no flag, no warning
    sd2 = RO.make_state_dict(1, 32, 5, seed=8)
    model.load_state_dict(sd2)
    model.RANGE_CHECK = True
    x2 = SparseTensor(feats.cuda(), coordinates=torch.from_numpy(coords).cuda())
    with warnings.catch_warnings(record=True) as w2:
        warnings.simplefilter('always')
        model(x2)
    assert not w2, [str(m.message) for m in w2]
    assert int(x2.coordinate_manager.range_status.item()) == 0


def test_forward_expanded_variant_vs_oracle():
    """ResUNetExpBN2C (model/resunet.py:254-492): two residual blocks per level with a stand-alone eval BatchNorm between
    them (eyoc_xh_affine on split-half rows) - same state-dict keys as the reference, features within 1e-5 of the oracle."""
    from eyoc_b200.model import load_model
    from eyoc_b200.sparse import SparseTensor
    from oracle import resunet_oracle as RO
    from tests.test_resunet_gpu import _cloud
    coords = _cloud(2500, 12, batch=2)
    feats = torch.ones((len(coords), 1), dtype=torch.float32)
    sd = RO.make_state_dict_expanded(1, 32, 5, seed=6)
    model = load_model('ResUNetExpBN2C')(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3)
    model.load_state_dict(sd)                                    # strict: every reference key present, none extra
    model = model.cuda().eval()
    want = RO.resunet_expanded_forward(coords, feats, sd, True, 5)
    got = model(SparseTensor(feats.cuda(), coordinates=torch.from_numpy(coords).cuda())).F.cpu()
    assert float((got - want).abs().max()) <= 1e-5, float((got - want).abs().max())


@pytest.mark.parametrize('mode', ['f16x3', 'tf32x3'])
def test_forward_instance_norm_variant_vs_oracle(mode):
    """ResUNetIN2C (model/resunet.py:239-241): batch norm on the trunk, ME.MinkowskiInstanceNorm inside every residual block
    (BasicBlockIN) - the convolution runs alone and eyoc_instance_norm fuses the block's residual add and ReLU.  Three clouds of
    different sizes in one batch: the statistics are per cloud."""
    from eyoc_b200 import nn as enn
    from eyoc_b200.model import load_model
    from eyoc_b200.sparse import SparseTensor
    from oracle import resunet_oracle as RO
    from tests.test_resunet_gpu import _cloud
    coords = np.concatenate([_cloud(1800, 4, batch=2), _cloud(700, 9, batch=1) + np.array([2, 0, 0, 0])]).astype(np.int32)
    feats = torch.ones((len(coords), 1), dtype=torch.float32)
    torch.manual_seed(4)
    model = load_model('ResUNetIN2C')(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3)
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            if isinstance(m, enn.MinkowskiInstanceNorm):
                m.weight.copy_(torch.rand(1, m.num_features, generator=g) + 0.5)
                m.bias.copy_(torch.randn(1, m.num_features, generator=g) * 0.1)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    assert sd['block1.norm1.weight'].shape == (1, 32) and 'block1.norm1.bn.weight' not in sd and 'norm1.bn.weight' in sd
    want = RO.resunet_forward(coords, feats, sd, True, 5)
    model = model.cuda().eval()
    old = enn.CONV_MODE
    try:
        enn.CONV_MODE = mode
        got = model(SparseTensor(feats.cuda(), coordinates=torch.from_numpy(coords).cuda())).F.cpu()
    finally:
        enn.CONV_MODE = old
    assert float((got - want).abs().max()) <= 2e-5, float((got - want).abs().max())


@pytest.mark.parametrize('c', [32, 256])
@pytest.mark.parametrize('packed', [False, True])
def test_instance_norm_kernel_vs_fp64(c, packed):
    """eyoc_instance_norm alone: rows of five clouds in shuffled order (a CTA's 256-row chunk then holds many clouds: the
    shared-memory slots and the direct global path both run), one cloud with a single row, residual + ReLU fused."""
    from eyoc_b200 import _C, nn as enn
    from eyoc_b200.sparse import xh_pack, xh_unpack
    rng = np.random.default_rng(c)
    n = 5000
    batch = rng.choice(5, n, p=[0.5, 0.3, 0.15, 0.0498, 0.0002]).astype(np.int32)
    batch[17] = 4
    coords = np.zeros((n, 4), np.int32)
    coords[:, 0] = batch
    x = (rng.normal(size=(n, c)) * rng.uniform(0.1, 30, c) + rng.normal(size=c) * 5).astype(np.float32)
    res = rng.normal(size=(n, c)).astype(np.float32)
    w, b = rng.uniform(0.5, 1.5, c).astype(np.float32), rng.normal(size=c).astype(np.float32)
    xd, rd = torch.from_numpy(x).cuda(), torch.from_numpy(res).cuda()
    if packed:
        xd, rd = xh_pack(xd), xh_pack(rd)
        x, res = xh_unpack(xd).cpu().numpy(), xh_unpack(rd).cpu().numpy()          # what the kernel really reads
    out = torch.empty_like(xd)
    lib = _C.lib()
    ws = torch.empty(lib.eyoc_instance_norm_workspace_bytes(5, c), dtype=torch.uint8, device='cuda')
    st = torch.zeros(1, dtype=torch.int32, device='cuda')
    cd, wd, bd = torch.from_numpy(coords).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda()   # alive across the call
    _C.check(lib.eyoc_instance_norm(_C.ptr(xd), int(packed), _C.ptr(cd), _C.c_int64(n), c, 5,
                                    _C.ptr(wd), _C.ptr(bd), _C.c_float(1e-8),
                                    _C.ptr(rd), int(packed), 1, _C.ptr(out), int(packed), _C.ptr(st), _C.ptr(ws),
                                    _C.c_size_t(ws.numel()), _C.stream()))
    got = (xh_unpack(out) if packed else out).cpu().numpy().astype(np.float64)
    want = np.empty((n, c))
    for k in range(5):
        sel = batch == k
        xb = x[sel].astype(np.float64)
        mu = xb.mean(0)
        var = ((xb - mu) ** 2).mean(0)
        want[sel] = (xb - mu) / np.sqrt(var + 1e-8) * w + b
    want = np.maximum(want + res, 0.0)
    tol = 3e-5 if not packed else 2e-4
    big = np.abs(want) < 1e3                     # the single-row cloud divides by sqrt(1e-8): compare it apart
    assert np.abs(got - want)[big].max() <= tol * max(1.0, np.abs(want[big]).max()), np.abs(got - want)[big].max()
    assert int(st.item()) == 0
