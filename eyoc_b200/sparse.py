"""Sparse-tensor plumbing with the MinkowskiEngine surface the reference touches.

``SparseTensor(features, coordinates=...)`` (scripts/test_kitti.py:143-147), ``.F`` / ``.C`` /
``coordinate_map_key`` / ``coordinate_manager`` / ``decomposed_coordinates_and_features``
(model/resunet.py:188-191, lib/trainer.py:1289), plus the coordinate manager that owns the per-level
coordinate hashes and the cached kernel maps (8 maps per batch, each reused by 1-5
convolutions).  All device work goes through the C-ABI (csrc/coordmap.cu); there is no CPU path.
"""
import numpy as np
import torch

from . import _C


def _pow2_at_least(x):
    c = 2
    while c < x:
        c *= 2
    return c


class CoordinateMapKey:
    """(tensor_stride,) identifier of one coordinate set inside a manager."""

    def __init__(self, tensor_stride):
        self.tensor_stride = int(tensor_stride)

    def get_tensor_stride(self):
        return [self.tensor_stride] * 3

    def __eq__(self, other):
        return isinstance(other, CoordinateMapKey) and other.tensor_stride == self.tensor_stride

    def __hash__(self):
        return hash(self.tensor_stride)

    def __repr__(self):
        return f'CoordinateMapKey(tensor_stride={self.tensor_stride})'


# Budget (bytes of input features) of one cloud group of the tile order.  Measured on B200 (bench.py --tile-group-mb): the
# gather is not L2 / DRAM bound (L2 throughput 21 %, DRAM 11 %), so denser tiles beat a smaller L2 working set - 24 MB:
# 54 % of the HBM roofline, 48 MB: 57 %, 96 MB: 60 %, 200 MB: 63 %, one group for the whole block: 66 %.
TILE_GROUP_BYTES = 4e9


class _Level:
    __slots__ = ('coords', 'n', 'keys', 'vals', 'cap', 'ts', 'n_read', 'coords_buf')


_PINNED_POOL = []
_SIDE_STREAMS = {}


class _AsyncRead:
    """A few int32 values on their way to the host, copied on a side stream right behind the kernel that produced them.
    ``.item()`` / ``.tolist()`` append their copy to the END of the launching stream and so wait for everything queued since;
    this read completes as soon as its producer has, however much work the launching stream has been given meanwhile - the
    host can keep the GPU's queue full across the data-dependent sizes of the coordinate levels."""

    def __init__(self, src):
        dev = src.device
        key = (dev.type, dev.index)
        if key not in _SIDE_STREAMS:
            _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
        side = _SIDE_STREAMS[key]
        self.n = src.numel()
        self.host = _PINNED_POOL.pop() if _PINNED_POOL else torch.empty(8, dtype=torch.int32).pin_memory()
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(side):
            side.wait_event(ready)
            self.host[:self.n].copy_(src.reshape(-1), non_blocking=True)
            self.done = torch.cuda.Event()
            self.done.record(side)
        src.record_stream(side)
        self.values = None

    def get(self):
        if self.values is None:
            self.done.synchronize()
            self.values = self.host[:self.n].tolist()
            _PINNED_POOL.append(self.host)
            self.host = None
        return self.values


class CoordinateManager:
    """Per-batch coordinate sets (tensor stride 1, 2, 4, ...) and kernel maps, all resident on one GPU."""

    def __init__(self, coordinates):
        if not coordinates.is_cuda:
            raise RuntimeError('eyoc_b200.SparseTensor needs CUDA coordinates (no CPU fallback)')
        if coordinates.dim() != 2 or coordinates.shape[1] != 4:
            raise RuntimeError('coordinates must be [N, 4] (batch, x, y, z)')
        self.device = coordinates.device
        self.levels = {}
        self._maps = {}
        self._tiled = {}
        self._tile_masks = {}
        self._perms = {}
        self.max_batch = 0
        self._status = torch.zeros(3, dtype=torch.int32, device=self.device)
        # [2]: bit 0 set by the split-half kernels when a value does not fit the fp16 hi/lo format (nn.CONV_MODE 'f16x3')
        self.range_status = self._status[2:3]
        c = coordinates.to(torch.int32).contiguous()
        lv = _Level()
        lv.coords, lv.n, lv.ts = c, c.shape[0], 1
        lv.cap = _pow2_at_least(2 * max(lv.n, 1))
        lv.keys = torch.empty(lv.cap, dtype=torch.int64, device=self.device)
        lv.vals = torch.empty(lv.cap, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _C.check(_C.lib().eyoc_hash_build(_C.ptr(c), _C.c_int64(lv.n), _C.ptr(lv.keys), _C.ptr(lv.vals),
                                              _C.c_int64(lv.cap), _C.ptr(self._status), _C.stream()))
        self.levels[1] = lv
        self._checked = False
        self._pending = {}
        self._status_read = _AsyncRead(self._status[:2])
        if lv.n > 0:
            self._start_level(2)            # queued right behind the hash build: its size is on the host long before it is needed

    # ------------------------------------------------------------------ levels
    def _check_status(self):
        if not self._checked:
            st, self.max_batch = (int(v) for v in self._status_read.get())
            if st & 1:
                raise RuntimeError('coordinates outside the packed 16-bit range (batch 0..65535, xyz -32768..32767)')
            if st & 2:
                raise RuntimeError('duplicate coordinates: quantise first (sparse_quantize)')
            self._checked = True

    def _start_level(self, ts2):
        """Queue the kernels of the stride-``ts2`` coordinate set (needs the finer level's row count on the host); its own
        row count travels back through an _AsyncRead."""
        fine = self.levels[ts2 // 2]
        lib = _C.lib()
        lv = _Level()
        lv.ts = ts2
        lv.cap = _pow2_at_least(2 * max(fine.n, 1))
        lv.keys = torch.empty(lv.cap, dtype=torch.int64, device=self.device)
        lv.vals = torch.empty(lv.cap, dtype=torch.int32, device=self.device)
        lv.coords_buf = torch.empty((fine.n, 4), dtype=torch.int32, device=self.device)
        n_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        ws = torch.empty(max(lib.eyoc_downsample_workspace_bytes(_C.c_int64(fine.n)), 8), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _C.check(lib.eyoc_coords_downsample(_C.ptr(fine.coords), _C.c_int64(fine.n), _C.c_int(ts2), _C.ptr(lv.keys),
                                                _C.ptr(lv.vals), _C.c_int64(lv.cap), _C.ptr(lv.coords_buf), _C.ptr(n_dev), _C.ptr(ws),
                                                _C.c_size_t(ws.numel()), _C.stream()))
            lv.n_read = _AsyncRead(n_dev)
        self._pending[ts2] = lv

    def _finish_level(self, ts2):
        lv = self._pending.pop(ts2)
        lv.n = int(lv.n_read.get()[0])
        lv.coords = lv.coords_buf[:lv.n]
        lv.n_read = lv.coords_buf = None
        self.levels[ts2] = lv
        if ts2 < self.EAGER_MAX_STRIDE and lv.n > 0:
            self._start_level(ts2 * 2)      # the next level is queued as soon as this one's size is known

    EAGER_MAX_STRIDE = 8                    # the ResUNets go down to tensor stride 8 (model/resunet.py:70)

    def ensure_levels(self, max_stride):
        """Build the stride-2 coordinate sets up to ``max_stride``.  Every level is QUEUED one level ahead of its first use
        and its row count (which sizes the next level's buffers, as in MinkowskiEngine's host-side coordinate manager) comes
        back through a side-stream read, so the host does not drain the GPU's queue to learn it."""
        ts2 = 2
        while ts2 <= max_stride:
            if ts2 not in self.levels:
                if ts2 not in self._pending:
                    self._start_level(ts2)
                self._finish_level(ts2)
            ts2 *= 2
        self._check_status()

    def num_rows(self, ts):
        return self.levels[ts].n

    # ------------------------------------------------------------------ kernel maps
    def kernel_map(self, ts_in, ts_out, ksize, transposed=False):
        """nbr [K, N_out] int32.  Forward: rows of the ts_in map at c_out + off*ts_in.  Transposed (ts_out < ts_in):
        rows of the coarse ts_in map at c_out - off*ts_out, same k (MinkowskiEngine convention, no flip)."""
        key = (ts_in, ts_out, ksize, transposed)
        if key not in self._maps:
            self.ensure_levels(max(ts_in, ts_out))
            lin, lout = self.levels[ts_in], self.levels[ts_out]
            step = -ts_out if transposed else ts_in
            K = ksize ** 3
            nbr = torch.empty((K, lout.n), dtype=torch.int32, device=self.device)
            lib = _C.lib()
            with torch.cuda.device(self.device):
                if transposed and ts_in == 2 * ts_out:
                    # mirror of the forward strided map (built once by the encoder, or here)
                    down = self.kernel_map(ts_out, ts_in, ksize, False)
                    _C.check(lib.eyoc_kernel_map_transpose(_C.ptr(down), _C.c_int64(lin.n), _C.c_int64(lout.n), _C.c_int(K),
                                                           _C.ptr(nbr), _C.stream()))
                elif not transposed and ts_in == ts_out:
                    _C.check(lib.eyoc_kernel_map_self(_C.ptr(lout.coords), _C.c_int64(lout.n), _C.ptr(lin.keys), _C.ptr(lin.vals),
                                                      _C.c_int64(lin.cap), _C.c_int(ksize), _C.c_int(step), _C.ptr(nbr),
                                                      _C.stream()))
                else:
                    _C.check(lib.eyoc_kernel_map(_C.ptr(lout.coords), _C.c_int64(lout.n), _C.ptr(lin.keys), _C.ptr(lin.vals),
                                                 _C.c_int64(lin.cap), _C.c_int(ksize), _C.c_int(step), _C.ptr(nbr),
                                                 _C.stream()))
            self._maps[key] = nbr
        return self._maps[key]

    def stem_conv(self, feats, weight, scale, shift, relu, ksize, packed):
        """The network's first convolution (1 input channel -> 32, 3^3 or 5^3 offsets, stride 1 on the stride-1 map) fused with
        its own neighbour search (eyoc_stem_conv: block-occupancy pre-filter, no ksize^3-column table).  Returns the output rows
        (split-half fp16 when ``packed``) and leaves the level's 3^3 neighbour table in the map cache as a by-product."""
        lv = self.levels[1]
        lib = _C.lib()
        out = torch.empty((lv.n, 64) if packed else (lv.n, 32), dtype=torch.float16 if packed else torch.float32, device=self.device)
        key3 = (1, 1, 3, False)
        nbr3 = torch.empty((27, lv.n), dtype=torch.int32, device=self.device) if key3 not in self._maps else None
        ws = torch.empty(lib.eyoc_stem_conv_workspace_bytes(_C.c_int64(lv.cap)), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _C.check(lib.eyoc_stem_conv(_C.ptr(lv.coords), _C.c_int64(lv.n), _C.ptr(lv.keys), _C.ptr(lv.vals), _C.c_int64(lv.cap),
                                        _C.c_int(ksize), _C.ptr(feats), _C.ptr(weight), _C.ptr(scale), _C.ptr(shift),
                                        _C.c_int(int(relu)), _C.ptr(out), _C.c_int(int(packed)), _C.ptr(self.range_status),
                                        _C.ptr(nbr3), _C.ptr(ws), _C.c_size_t(ws.numel()), _C.stream()))
        if nbr3 is not None:
            self._maps[key3] = nbr3
        return out

    def tiled_map(self, ts_in, ts_out, ksize, transposed=False):
        """(nbr_tiled [K, N_out], row_perm [N_out]) for the tensor-core convolution: output rows sorted by
        (cloud group, neighbour-pattern bit mask) so that 128-row tiles are dense or skipped per kernel offset
        (csrc/coordmap.cu: eyoc_tile_order).  The cloud group bounds the gather's L2 working set."""
        key = (ts_in, ts_out, ksize, transposed)
        if key not in self._tiled:
            nbr = self.kernel_map(ts_in, ts_out, ksize, transposed)
            self._check_status()
            lin, lout = self.levels[ts_in], self.levels[ts_out]
            K, n_out = nbr.shape
            rows_per_cloud = max(1, lin.n // (self.max_batch + 1))
            cin_hint = min(256, 64 * ts_in)          # widest input the maps of this level feed (ResUNet channel table)
            group = max(1, min(self.max_batch + 1, int(TILE_GROUP_BYTES // (rows_per_cloud * 4 * cin_hint))))
            perm = torch.empty(n_out, dtype=torch.int32, device=self.device)
            tiled = torch.empty_like(nbr)
            masks = torch.empty(max(1, (n_out + 255) // 256), dtype=torch.int32, device=self.device)
            lib = _C.lib()
            ws = torch.empty(max(lib.eyoc_tile_order_workspace_bytes(_C.c_int64(n_out)), 8), dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                _C.check(lib.eyoc_tile_order(_C.ptr(nbr), _C.c_int(K), _C.c_int64(n_out), _C.ptr(lout.coords), _C.c_int(group),
                                             _C.c_int(self.max_batch), _C.ptr(perm), _C.ptr(tiled), _C.ptr(masks), _C.ptr(ws),
                                             _C.c_size_t(ws.numel()), _C.stream()))
            self._tiled[key] = (tiled, perm)
            if n_out > 0:
                self._tile_masks[key] = masks         # per 256-row tile, out of the same pass (else eyoc_tile_masks)
        return self._tiled[key]

    def tile_masks(self, ts_in, ts_out, ksize, transposed=False):
        """Per 256-row tile of ``tiled_map`` the bit mask of kernel offsets that have a neighbour (eyoc_tile_masks):
        computed once per map instead of by every CTA of every convolution that uses it."""
        key = (ts_in, ts_out, ksize, transposed)
        if key not in self._tile_masks:
            tiled, _ = self.tiled_map(ts_in, ts_out, ksize, transposed)
            K, n_out = tiled.shape
            masks = torch.empty(max(1, (n_out + 255) // 256), dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                _C.check(_C.lib().eyoc_tile_masks(_C.ptr(tiled), _C.c_int(K), _C.c_int64(n_out), _C.ptr(masks), _C.stream()))
            self._tile_masks[key] = masks
        return self._tile_masks[key]

    def parity_perm(self, ts):
        """Row order of level ``ts`` grouped by parity class (CTA-uniform offsets for transposed convs)."""
        if ts not in self._perms:
            lv = self.levels[ts]
            cls = torch.empty(lv.n, dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                _C.check(_C.lib().eyoc_parity_class(_C.ptr(lv.coords), _C.c_int64(lv.n), _C.c_int(ts), _C.ptr(cls),
                                                    _C.stream()))
            self._perms[ts] = torch.sort(cls, stable=True)[1].to(torch.int32)
        return self._perms[ts]


def xh_pack(x, range_status=None):
    """fp32 [n, c] (c % 32 == 0) -> split-half rows [n, 2 c] fp16 (eyoc_xh_pack)."""
    _C.require_cuda(x)
    x = x.to(torch.float32).contiguous()
    n, c = x.shape
    out = torch.empty((n, 2 * c), dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device):
        _C.check(_C.lib().eyoc_xh_pack(_C.ptr(x), _C.c_int64(n), _C.c_int(c), _C.ptr(out), _C.ptr(range_status), _C.stream()))
    return out


def xh_unpack(xh):
    """split-half rows [n, 2 c] fp16 -> fp32 [n, c] (eyoc_xh_unpack)."""
    _C.require_cuda(xh)
    n, c = xh.shape[0], xh.shape[1] // 2
    out = torch.empty((n, c), dtype=torch.float32, device=xh.device)
    with torch.cuda.device(xh.device):
        _C.check(_C.lib().eyoc_xh_unpack(_C.ptr(xh.contiguous()), _C.c_int64(n), _C.c_int(c), _C.ptr(out), _C.stream()))
    return out


def ones_features(n, device):
    """The reference's network input - ``torch.ones((N, 1))`` occupancy features (lib/data_loaders.py:971-972) - tagged so that
    the first convolution knows every value is 1.0 and does not gather it (eyoc_stem_conv with ``in`` = NULL).  Do not write
    to the returned tensor."""
    t = torch.ones((n, 1), dtype=torch.float32, device=device)
    t._eyoc_all_ones = True
    return t


class SparseTensor:
    """Features [N, C] fp32 on a coordinate set; row order == input order (the reference relies on it).

    Inside the network the features may exist only in the SPLIT-HALF format of the fp16 tensor-core convolutions
    (``features_xh`` [N, 2 C] fp16: per 32-channel chunk 32 hi | 32 lo' halves, include/eyoc_b200.h); ``.F`` then
    converts once, on first access, and caches the fp32 tensor."""

    def __init__(self, features=None, coordinates=None, coordinate_map_key=None, coordinate_manager=None, device=None,
                 tensor_stride=1, features_xh=None):
        if features is None and features_xh is None:
            raise RuntimeError('SparseTensor needs features')
        if device is not None and features is not None:
            features = features.to(device)
            coordinates = coordinates.to(device) if coordinates is not None else None
        _C.require_cuda(features, features_xh)
        dev = features.device if features is not None else features_xh.device
        if coordinate_manager is None:
            if coordinates is None:
                raise RuntimeError('SparseTensor needs coordinates or a coordinate_manager')
            coordinate_manager = CoordinateManager(coordinates.to(dev))
            coordinate_map_key = CoordinateMapKey(tensor_stride)
        elif coordinate_map_key is None:
            coordinate_map_key = CoordinateMapKey(tensor_stride)
        self._F = features
        self._Fh = features_xh
        self.all_ones = bool(getattr(features, '_eyoc_all_ones', False))       # see ones_features()
        self.coordinate_manager = coordinate_manager
        self.coordinate_map_key = coordinate_map_key
        n = coordinate_manager.levels[coordinate_map_key.tensor_stride].coords.shape[0]
        rows = features.shape[0] if features is not None else features_xh.shape[0]
        if rows != n:
            raise RuntimeError(f'features have {rows} rows, coordinates {n}')

    @property
    def F(self):
        if self._F is None:
            self._F = xh_unpack(self._Fh)
        return self._F

    @property
    def Fh(self):
        """Split-half image of the features (packed on first access; needs C % 32 == 0)."""
        if self._Fh is None:
            self._Fh = xh_pack(self._F, self.coordinate_manager.range_status)
        return self._Fh

    @property
    def num_channels(self):
        return self._F.shape[1] if self._F is not None else self._Fh.shape[1] // 2

    @property
    def C(self):
        return self.coordinate_manager.levels[self.coordinate_map_key.tensor_stride].coords

    @property
    def coordinates(self):
        return self.C

    @property
    def features(self):
        return self.F

    @property
    def device(self):
        return self._F.device if self._F is not None else self._Fh.device

    @property
    def tensor_stride(self):
        return self.coordinate_map_key.get_tensor_stride()

    def __len__(self):
        return self._F.shape[0] if self._F is not None else self._Fh.shape[0]

    @property
    def shape(self):
        return torch.Size((len(self), self.num_channels))

    @property
    def decomposed_coordinates_and_features(self):
        """Per-batch-index lists (lib/trainer.py:1289)."""
        C, F = self.C, self.F
        nb = int(C[:, 0].max().item()) + 1 if len(C) else 0
        coords, feats = [], []
        for b in range(nb):
            m = C[:, 0] == b
            coords.append(C[m, 1:])
            feats.append(F[m])
        return coords, feats

    @property
    def decomposed_features(self):
        return self.decomposed_coordinates_and_features[1]

    @property
    def decomposed_coordinates(self):
        return self.decomposed_coordinates_and_features[0]


# ---------------------------------------------------------------------------------- ME.utils (host-side prep)
def sparse_quantize(coordinates, features=None, return_index=False, quantization_size=None):
    """ME.utils.sparse_quantize as the reference uses it (lib/data_loaders.py:940): floor, unique rows, FIRST
    occurrence kept, index returned in ascending order."""
    is_torch = isinstance(coordinates, torch.Tensor)
    c = coordinates.cpu().numpy() if is_torch else np.asarray(coordinates)
    if quantization_size is not None:
        c = c / quantization_size
    q = np.floor(c).astype(np.int32)
    _, first = np.unique(q, axis=0, return_index=True)
    sel = np.sort(first)
    out = torch.from_numpy(q[sel]) if is_torch else q[sel]
    if return_index:
        idx = torch.from_numpy(sel) if is_torch else sel
        return (out, idx) if features is None else (out, features[sel], idx)
    return out if features is None else (out, features[sel])


def voxelize_gpu(xyz, voxel_size, cloud=None):
    """Device-side voxelisation + collate of raw points (the step in front of the hot path, lib/data_loaders.py:936-979
    + :31-85): returns (coords int32 [m, 4] (batch, x, y, z), sel int64 [m]) with ``coords = floor(xyz[sel] / voxel_size)``
    and ``sel`` = first occurrence of every occupied voxel per cloud, ascending - what
    ``ME.utils.sparse_quantize(xyz / voxel_size, return_index=True)`` selects.  xyz [n, 3] CUDA fp32; cloud [n] int32
    batch index per point (None = one cloud)."""
    _C.require_cuda(xyz, cloud)
    xyz = xyz.to(torch.float32).contiguous()
    n = xyz.shape[0]
    dev = xyz.device
    if cloud is not None:
        cloud = cloud.to(torch.int32).contiguous()
    cap = _pow2_at_least(2 * max(n, 1))
    keys = torch.empty(cap, dtype=torch.int64, device=dev)
    vals = torch.empty(cap, dtype=torch.int32, device=dev)
    coords = torch.empty((n, 4), dtype=torch.int32, device=dev)
    sel = torch.empty(n, dtype=torch.int32, device=dev)
    n_out = torch.zeros(1, dtype=torch.int32, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    lib = _C.lib()
    ws = torch.empty(max(lib.eyoc_voxelize_workspace_bytes(_C.c_int64(n)), 8), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _C.check(lib.eyoc_voxelize(_C.ptr(xyz), _C.ptr(cloud), _C.c_int64(n), _C.c_float(float(voxel_size)), _C.ptr(keys),
                                   _C.ptr(vals), _C.c_int64(cap), _C.ptr(coords), _C.ptr(sel), _C.ptr(n_out), _C.ptr(status),
                                   _C.ptr(ws), _C.c_size_t(ws.numel()), _C.stream()))
    m, st = int(n_out.item()), int(status[0].item())
    if st & 1:
        raise RuntimeError('voxel coordinates outside the packed 16-bit range (batch 0..65535, xyz -32768..32767)')
    return coords[:m], sel[:m].long()


def batched_coordinates(coords, dtype=torch.int32, device=None):
    """ME.utils.batched_coordinates: prepend the batch index -> int32 [sum N, 4]."""
    out = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(c).to(dtype)
        out.append(torch.cat([torch.full((len(c), 1), b, dtype=dtype), c], 1))
    out = torch.cat(out, 0) if out else torch.zeros((0, 4), dtype=dtype)
    return out.to(device) if device is not None else out


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    """ME.utils.sparse_collate (lib/data_loaders.py:65-66): batched coordinates + concatenated features."""
    C = batched_coordinates(coords, dtype=dtype, device=device)
    F = torch.cat([torch.as_tensor(f) for f in feats], 0)
    if device is not None:
        F = F.to(device)
    if labels is not None:
        return C, F, torch.cat([torch.as_tensor(l) for l in labels], 0)
    return C, F
