"""CPU oracle: MinkowskiEngine-semantics ResUNetBN2C (FCGF feature extractor).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED against MinkowskiEngine
itself (un-vendored, un-pinned dependency: README.md:27,61 "v0.5 or higher", absent from
/root/reference and from this image).  This file restates ME 0.5.x's *published* sparse
convolution algorithm: per-kernel-offset gather -> dense GEMM ->
scatter-add, k ascending, generalized sparse convolution on integer coordinate maps.  It is
pinned by ``tests/test_oracle_resunet.py`` against dense ``F.conv3d`` / ``F.conv_transpose3d``
on a densified grid.  Topology follows model/resunet.py:18-193,206-209, model/residual_block.py:13-53
and model/common.py:4-10 (citations relative to /root/reference).
"""
import numpy as np
import torch
import torch.nn.functional as F

CHANNELS = [None, 32, 64, 128, 256]      # model/resunet.py:208
TR_CHANNELS = [None, 64, 64, 64, 128]    # model/resunet.py:209
BN_EPS = 1e-5                            # torch.nn.BatchNorm1d default, via ME.MinkowskiBatchNorm


def kernel_offsets(ksize):
    """Offsets of a hyper-cubic kernel, first spatial axis fastest: k = ix + K*(iy + K*iz)
    (SURVEY.md Appendix B.2)."""
    r = (ksize - 1) // 2
    ax = np.arange(-r, r + 1)
    oz, oy, ox = np.meshgrid(ax, ax, ax, indexing='ij')
    return np.stack([ox.ravel(), oy.ravel(), oz.ravel()], 1).astype(np.int64)   # [K,3]


def _pack(c):
    """[N,4] int (b,x,y,z) -> int64 key; 16 bits per field, spatial biased by 2^15."""
    c = c.astype(np.int64)
    return (c[:, 0] << 48) | ((c[:, 1] + 32768) << 32) | ((c[:, 2] + 32768) << 16) | (c[:, 3] + 32768)


class CoordMap:
    """One coordinate set at a tensor stride, with O(log n) lookup."""

    def __init__(self, coords, tensor_stride):
        self.coords = np.ascontiguousarray(coords, np.int64)
        self.ts = tensor_stride
        keys = _pack(self.coords)
        self.order = np.argsort(keys, kind='stable')
        self.sorted_keys = keys[self.order]
        assert len(np.unique(keys)) == len(keys), 'duplicate coordinates'

    def lookup(self, coords):
        keys = _pack(coords)
        pos = np.searchsorted(self.sorted_keys, keys)
        pos = np.minimum(pos, len(self.sorted_keys) - 1)
        hit = self.sorted_keys[pos] == keys
        return np.where(hit, self.order[pos], -1)

    def downsample(self):
        """Stride-2 coordinate set: unique(floor(c / 2ts) * 2ts), first-occurrence order
        (Appendix B.4; the order is unobservable)."""
        ts2 = self.ts * 2
        c = self.coords.copy()
        c[:, 1:] = np.floor_divide(c[:, 1:], ts2) * ts2
        keys = _pack(c)
        _, first = np.unique(keys, return_index=True)
        return CoordMap(c[np.sort(first)], ts2)


def kernel_map(in_map, out_map, ksize, transposed=False):
    """Neighbour table nbr[K, N_out] (int64, -1 = absent).

    forward   (Appendix B.3/B.4): in row at  c_out + off_k * ts_in
    transposed (Appendix B.5):    in (coarse) row at c_out - off_k * ts_out, same k (no flip)
    """
    offs = kernel_offsets(ksize)
    ts = out_map.ts if transposed else in_map.ts
    sign = -1 if transposed else 1
    nbr = np.empty((len(offs), len(out_map.coords)), np.int64)
    for k, o in enumerate(offs):
        q = out_map.coords.copy()
        q[:, 1:] += sign * o[None, :] * ts
        nbr[k] = in_map.lookup(q)
    return nbr


def sparse_conv(x, W, nbr):
    """out[o] = sum_k x[nbr[k,o]] @ W[k]  -- offset by offset, k ascending (Appendix B.3/B.9)."""
    out = torch.zeros((nbr.shape[1], W.shape[-1]), dtype=x.dtype)
    for k in range(nbr.shape[0]):
        o = np.nonzero(nbr[k] >= 0)[0]
        if len(o) == 0:
            continue
        out.index_add_(0, torch.from_numpy(o), x[torch.from_numpy(nbr[k, o])] @ W[k])
    return out


def make_state_dict(in_channels=1, out_channels=32, conv1_kernel_size=5, seed=0, dtype=torch.float32):
    """Random-init weights with MinkowskiEngine's state_dict key names/shapes
    (SURVEY.md §5 checkpoint row, §8d 'Weights')."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, K, cin, cout, bias=False):
        bound = 1.0 / np.sqrt(cin * K)
        shape = (cin, cout) if K == 1 else (K, cin, cout)
        sd[name + '.kernel'] = ((torch.rand(shape, generator=g) * 2 - 1) * bound).to(dtype)
        if bias:
            sd[name + '.bias'] = (torch.randn((1, cout), generator=g) * 0.1).to(dtype)

    def bn(name, c):
        sd[name + '.bn.weight'] = (torch.rand(c, generator=g) + 0.5).to(dtype)
        sd[name + '.bn.bias'] = (torch.randn(c, generator=g) * 0.1).to(dtype)
        sd[name + '.bn.running_mean'] = (torch.randn(c, generator=g) * 0.1).to(dtype)
        sd[name + '.bn.running_var'] = (torch.rand(c, generator=g) + 0.5).to(dtype)
        sd[name + '.bn.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)

    def block(name, c):
        conv(name + '.conv1', 27, c, c); bn(name + '.norm1', c)
        conv(name + '.conv2', 27, c, c); bn(name + '.norm2', c)

    C, T = CHANNELS, TR_CHANNELS
    conv('conv1', conv1_kernel_size ** 3, in_channels, C[1]); bn('norm1', C[1]); block('block1', C[1])
    conv('conv2', 27, C[1], C[2]); bn('norm2', C[2]); block('block2', C[2])
    conv('conv3', 27, C[2], C[3]); bn('norm3', C[3]); block('block3', C[3])
    conv('conv4', 27, C[3], C[4]); bn('norm4', C[4]); block('block4', C[4])
    conv('conv4_tr', 27, C[4], T[4]); bn('norm4_tr', T[4]); block('block4_tr', T[4])
    conv('conv3_tr', 27, C[3] + T[4], T[3]); bn('norm3_tr', T[3]); block('block3_tr', T[3])
    conv('conv2_tr', 27, C[2] + T[3], T[2]); bn('norm2_tr', T[2]); block('block2_tr', T[2])
    conv('conv1_tr', 1, C[1] + T[2], T[1])
    conv('final', 1, T[1], out_channels, bias=True)
    return sd


def _bn(x, sd, name):
    """ME.MinkowskiBatchNorm in eval mode = nn.BatchNorm1d on .F (Appendix B.7)."""
    return F.batch_norm(x, sd[name + '.bn.running_mean'], sd[name + '.bn.running_var'],
                        sd[name + '.bn.weight'], sd[name + '.bn.bias'], False, 0.0, BN_EPS)


IN_EPS = 1e-8


def _in(x, sd, name, batch):
    """ME.MinkowskiInstanceNorm (model/common.py:7-8) restated from MinkowskiEngine 0.5's MinkowskiNormalization.py
    [ME-upstream, not runnable here]: per cloud, mean = global average pool of x, var = global average pool of (x - mean)^2
    (biased), out = (x - mean) * (1 / sqrt(var + 1e-8)) * weight + bias with [1, C] parameters."""
    out = torch.empty_like(x)
    batch = torch.as_tensor(np.asarray(batch))
    for b in torch.unique(batch).tolist():
        sel = batch == b
        xb = x[sel]
        mean = xb.double().mean(0).to(x.dtype)
        centred = xb - mean
        var = (centred.double() ** 2).mean(0).to(x.dtype)
        out[sel] = centred * (1.0 / torch.sqrt(var + IN_EPS))
    return out * sd[name + '.weight'].reshape(1, -1) + sd[name + '.bias'].reshape(1, -1)


def _norm(x, sd, name, batch):
    return _bn(x, sd, name) if (name + '.bn.weight') in sd else _in(x, sd, name, batch)


def _block(x, sd, name, nbr, batch=None):
    """model/residual_block.py:37-53 (BasicBlockBN / BasicBlockIN by the keys the state dict holds)."""
    out = F.relu(_norm(sparse_conv(x, sd[name + '.conv1.kernel'], nbr), sd, name + '.norm1', batch))
    out = _norm(sparse_conv(out, sd[name + '.conv2.kernel'], nbr), sd, name + '.norm2', batch)
    return F.relu(out + x)


def build_maps(coords, conv1_kernel_size=5):
    """The 8 kernel maps of one (batched) cloud: {k5 s1}, {k3 s1}x4 levels, {k3 s2}x3 (Appendix A notes)."""
    m1 = CoordMap(np.asarray(coords), 1)
    m2 = m1.downsample(); m4 = m2.downsample(); m8 = m4.downsample()
    maps = {'levels': [m1, m2, m4, m8],
            'k5': kernel_map(m1, m1, conv1_kernel_size),
            's1': [kernel_map(m, m, 3) for m in (m1, m2, m4, m8)],
            'down': [kernel_map(a, b, 3) for a, b in ((m1, m2), (m2, m4), (m4, m8))],
            'up': [kernel_map(b, a, 3, transposed=True) for a, b in ((m1, m2), (m2, m4), (m4, m8))]}
    return maps


def resunet_forward(coords, feats, sd, normalize_feature=True, conv1_kernel_size=5, detail=None, maps=None):
    """model/resunet.py:142-193.  coords [N,4] int (b,x,y,z), feats [N,Cin] -> F [N,32]."""
    maps = maps or build_maps(coords, conv1_kernel_size)
    s1, dn, up = maps['s1'], maps['down'], maps['up']
    bt = [m.coords[:, 0] for m in maps['levels']]              # batch column per level (instance-norm blocks)
    x = torch.as_tensor(feats)
    o1 = _bn(sparse_conv(x, sd['conv1.kernel'], maps['k5']), sd, 'norm1')
    o1 = _block(o1, sd, 'block1', s1[0], bt[0]); out = F.relu(o1)
    o2 = _bn(sparse_conv(out, sd['conv2.kernel'], dn[0]), sd, 'norm2')
    o2 = _block(o2, sd, 'block2', s1[1], bt[1]); out = F.relu(o2)
    o4 = _bn(sparse_conv(out, sd['conv3.kernel'], dn[1]), sd, 'norm3')
    o4 = _block(o4, sd, 'block3', s1[2], bt[2]); out = F.relu(o4)
    o8 = _bn(sparse_conv(out, sd['conv4.kernel'], dn[2]), sd, 'norm4')
    o8 = _block(o8, sd, 'block4', s1[3], bt[3]); out = F.relu(o8)

    out = _bn(sparse_conv(out, sd['conv4_tr.kernel'], up[2]), sd, 'norm4_tr')
    o4t = F.relu(_block(out, sd, 'block4_tr', s1[2], bt[2]))
    out = torch.cat([o4t, o4], 1)
    out = _bn(sparse_conv(out, sd['conv3_tr.kernel'], up[1]), sd, 'norm3_tr')
    o2t = F.relu(_block(out, sd, 'block3_tr', s1[1], bt[1]))
    out = torch.cat([o2t, o2], 1)
    out = _bn(sparse_conv(out, sd['conv2_tr.kernel'], up[0]), sd, 'norm2_tr')
    o1t = F.relu(_block(out, sd, 'block2_tr', s1[0], bt[0]))
    out = torch.cat([o1t, o1], 1)
    out = F.relu(out @ sd['conv1_tr.kernel'])                      # 1x1 conv (Appendix B.6)
    out = out @ sd['final.kernel'] + sd['final.bias']
    if detail is not None:
        detail.update(out_s1=o1, out_s2=o2, out_s4=o4, out_s8=o8, out_s4_tr=o4t, out_s2_tr=o2t,
                      out_s1_tr=o1t, pre_norm=out, maps=maps)
    if normalize_feature:
        out = out / torch.norm(out, p=2, dim=1, keepdim=True)      # no epsilon (resunet.py:189)
    return out


# ------------------------------------------------------------- dense cross-check
def dense_conv_reference(coords_in, x, coords_out, W, ksize, ts_in, stride, transposed):
    """The same generalized sparse convolution evaluated by densifying onto a regular grid and
    calling F.conv3d / F.conv_transpose3d (single batch index).  Used to pin sparse_conv+kernel_map."""
    ci = np.asarray(coords_in)[:, 1:] // ts_in
    ts_out = ts_in // stride if transposed else ts_in * stride
    co = np.asarray(coords_out)[:, 1:] // ts_out
    K = ksize
    Wd = W.reshape(K, K, K, W.shape[-2], W.shape[-1])              # [iz,iy,ix,cin,cout]
    Wd = Wd.permute(4, 3, 2, 1, 0).contiguous()                    # [cout,cin,ix,iy,iz]
    pad = 2 * K
    lo = ci.min(0) - pad
    lo -= lo % 2                                                   # keep parity of coarse grid
    size = ci.max(0) - lo + 1 + pad
    g = torch.zeros((1, x.shape[1], *size), dtype=x.dtype)
    idx = ci - lo
    g[0, :, idx[:, 0], idx[:, 1], idx[:, 2]] = x.t()
    r = (K - 1) // 2
    if not transposed:
        y = F.conv3d(g, Wd, stride=stride, padding=r)
        oi = (co * stride - lo) // stride if stride > 1 else co - lo
        if stride > 1:
            oi = co - lo // stride
    else:
        y = F.conv_transpose3d(g, Wd.permute(1, 0, 2, 3, 4).contiguous(), stride=stride, padding=r)
        oi = co - lo * stride
    return y[0, :, oi[:, 0], oi[:, 1], oi[:, 2]].t()


# ------------------------------------------------------------- ResUNetExpanded (model/resunet.py:254-492)
def make_state_dict_expanded(in_channels=1, out_channels=32, conv1_kernel_size=5, seed=0, channels=None, tr_channels=None,
                             dtype=torch.float32):
    """State dict of ResUNetExpBN2C: ResUNetBN2C's keys plus ``norm<L>_2`` / ``block<L>_2`` for every level."""
    C = channels or CHANNELS
    T = tr_channels or TR_CHANNELS
    sd = make_state_dict(in_channels, out_channels, conv1_kernel_size, seed=seed, dtype=dtype)
    g = torch.Generator().manual_seed(seed + 1000)
    for name, c in (('1', C[1]), ('2', C[2]), ('3', C[3]), ('4', C[4]), ('4_tr', T[4]), ('3_tr', T[3]), ('2_tr', T[2])):
        pre = f'norm{name}_2'
        sd[pre + '.bn.weight'] = (torch.rand(c, generator=g) + 0.5).to(dtype)
        sd[pre + '.bn.bias'] = (torch.randn(c, generator=g) * 0.1).to(dtype)
        sd[pre + '.bn.running_mean'] = (torch.randn(c, generator=g) * 0.1).to(dtype)
        sd[pre + '.bn.running_var'] = (torch.rand(c, generator=g) + 0.5).to(dtype)
        sd[pre + '.bn.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)
        for cv in ('conv1', 'conv2'):
            bound = 1.0 / np.sqrt(c * 27)
            sd[f'block{name}_2.{cv}.kernel'] = ((torch.rand((27, c, c), generator=g) * 2 - 1) * bound).to(dtype)
        for nm in ('norm1', 'norm2'):
            p2 = f'block{name}_2.{nm}'
            sd[p2 + '.bn.weight'] = (torch.rand(c, generator=g) + 0.5).to(dtype)
            sd[p2 + '.bn.bias'] = (torch.randn(c, generator=g) * 0.1).to(dtype)
            sd[p2 + '.bn.running_mean'] = (torch.randn(c, generator=g) * 0.1).to(dtype)
            sd[p2 + '.bn.running_var'] = (torch.rand(c, generator=g) + 0.5).to(dtype)
            sd[p2 + '.bn.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)
    return sd


def resunet_expanded_forward(coords, feats, sd, normalize_feature=True, conv1_kernel_size=5, maps=None):
    """model/resunet.py:396-486 (ResUNetExpanded.forward)."""
    maps = maps or build_maps(coords, conv1_kernel_size)
    s1, dn, up = maps['s1'], maps['down'], maps['up']

    def level(x, name, nbr):
        x = F.relu(_block(x, sd, f'block{name}', nbr))
        x = _bn(x, sd, f'norm{name}_2')
        return F.relu(_block(x, sd, f'block{name}_2', nbr))

    x = torch.as_tensor(feats)
    o1 = level(_bn(sparse_conv(x, sd['conv1.kernel'], maps['k5']), sd, 'norm1'), '1', s1[0])
    o2 = level(_bn(sparse_conv(o1, sd['conv2.kernel'], dn[0]), sd, 'norm2'), '2', s1[1])
    o4 = level(_bn(sparse_conv(o2, sd['conv3.kernel'], dn[1]), sd, 'norm3'), '3', s1[2])
    o8 = level(_bn(sparse_conv(o4, sd['conv4.kernel'], dn[2]), sd, 'norm4'), '4', s1[3])
    o4t = level(_bn(sparse_conv(o8, sd['conv4_tr.kernel'], up[2]), sd, 'norm4_tr'), '4_tr', s1[2])
    o2t = level(_bn(sparse_conv(torch.cat([o4t, o4], 1), sd['conv3_tr.kernel'], up[1]), sd, 'norm3_tr'), '3_tr', s1[1])
    o1t = level(_bn(sparse_conv(torch.cat([o2t, o2], 1), sd['conv2_tr.kernel'], up[0]), sd, 'norm2_tr'), '2_tr', s1[0])
    out = F.relu(torch.cat([o1t, o1], 1) @ sd['conv1_tr.kernel'])
    out = out @ sd['final.kernel'] + sd['final.bias']
    if normalize_feature:
        out = out / torch.norm(out, p=2, dim=1, keepdim=True)
    return out
