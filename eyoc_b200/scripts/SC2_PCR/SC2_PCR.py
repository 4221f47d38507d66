"""Drop-in for the reference's scripts/SC2_PCR/SC2_PCR.py ``Matcher`` on the fused sm_100a estimator.

Same constructor and the same public entry points on the inference path:
    Matcher.match_pair   (SC2_PCR.py:280-305)
    Matcher.SC2_PCR      (SC2_PCR.py:307-384)  -> (final_trans [bs,4,4], seedwise_fitness [bs,S])
    Matcher.estimator    (SC2_PCR.py:386-413)  -> 5-tuple
Extension over the reference: bs >= 1 (the reference asserts bs == 1, SC2_PCR.py:44,249); for
bs > 1 each batch item is an independent pair (match_pair matches per item).
Tie rule of the three descending sorts: value, then lowest index (the reference leaves it to the
sort backend).  Host RNG draws happen in the reference's order (:288-289).
"""
import numpy as np
import torch

from ... import _C
from ...lib.eval import knn1
from ...lib.timer import Timer  # noqa: F401  (the reference module exposes it too, SC2_PCR.py:5)
from .common import rigid_transform_3d  # noqa: F401
from .utils.SE3 import transform  # noqa: F401

_DETAIL_INT = ('seeds', 'topk1', 'topk2')


class Matcher():
    def __init__(self,
                 inlier_threshold=0.10,
                 num_node='all',
                 use_mutual=True,
                 d_thre=0.1,
                 num_iterations=10,
                 ratio=0.2,
                 nms_radius=0.1,
                 max_points=8000,
                 k1=30,
                 k2=20,
                 heatmap=False,
                 ):
        self.inlier_threshold = inlier_threshold
        self.num_node = num_node
        self.use_mutual = use_mutual          # stored, never read - like the reference (SC2_PCR.py:23)
        self.d_thre = d_thre
        self.num_iterations = num_iterations
        self.ratio = ratio
        self.max_points = max_points
        self.nms_radius = nms_radius
        self.k1 = k1
        self.k2 = k2
        self.heatmap = heatmap
        self._ws = None

    # ------------------------------------------------------------------ config / workspace
    def _cfg(self):
        refine_thr = 0.10 if self.inlier_threshold == 0.10 else 1.2          # SC2_PCR.py:254-257
        return _C.SC2Cfg(inlier_threshold=self.inlier_threshold, d_thre=self.d_thre, d_thre_half=self.d_thre / 2,
                         d_thre_sq=self.d_thre ** 2, nms_radius=self.nms_radius, refine_threshold=refine_thr,
                         num_iterations=int(self.num_iterations), k1=int(self.k1), k2=int(self.k2),
                         refine_iterations=20)

    def _workspace(self, nbytes, device):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._ws

    # ------------------------------------------------------------------ matching
    def match_pair(self, src_keypts, tgt_keypts, src_features, tgt_features):
        """SC2_PCR.py:280-305.  Draw order on the global numpy RNG: src then tgt (:288-289)."""
        N_src, N_tgt = src_features.shape[1], tgt_features.shape[1]
        if self.num_node == 'all':
            src_sel_ind, tgt_sel_ind = np.arange(N_src), np.arange(N_tgt)
        else:
            src_sel_ind = np.random.choice(N_src, self.num_node)
            tgt_sel_ind = np.random.choice(N_tgt, self.num_node)
        dev = src_features.device
        si = torch.from_numpy(src_sel_ind).to(dev)
        ti = torch.from_numpy(tgt_sel_ind).to(dev)
        src_desc, tgt_desc = src_features[:, si, :], tgt_features[:, ti, :]
        src_keypts, tgt_keypts = src_keypts[:, si, :], tgt_keypts[:, ti, :]
        source_idx = knn1(src_desc, tgt_desc, form=1)                      # [bs, n] int64, per batch item
        self._last_match = (src_sel_ind, tgt_sel_ind, source_idx)
        tgt_keypts_corr = torch.gather(tgt_keypts, 1, source_idx[:, :, None].expand(-1, -1, 3))
        return src_keypts, tgt_keypts_corr

    # ------------------------------------------------------------------ estimator core
    def _run(self, src_keypts, tgt_keypts, want_labels, detail=None, hooks=None):
        _C.require_cuda(src_keypts, tgt_keypts)
        bs, num_corr = src_keypts.shape[0], tgt_keypts.shape[1]
        if num_corr > self.max_points:                                     # SC2_PCR.py:324-327
            src_keypts = src_keypts[:, :self.max_points, :]
            tgt_keypts = tgt_keypts[:, :self.max_points, :]
            num_corr = self.max_points
        src, tgt = _C.f32c(src_keypts), _C.f32c(tgt_keypts)
        dev = src.device
        S = int(num_corr * self.ratio)                                     # SC2_PCR.py:350
        cfg = self._cfg()
        lib = _C.lib()
        nbytes = lib.eyoc_sc2pcr_workspace_bytes(_C.c_int(bs), _C.c_int(num_corr), _C.c_int(max(S, 1)),
                                                 _C.ctypes.byref(cfg))
        ws = self._workspace(max(nbytes, 256), dev)
        trans = torch.empty((bs, 4, 4), dtype=torch.float32, device=dev)
        fitness = torch.empty((bs, max(S, 0)), dtype=torch.float32, device=dev)
        labels = torch.empty((bs, num_corr), dtype=torch.float32, device=dev) if want_labels else None
        hk = None
        keep = []
        if hooks:
            hk = _C.SC2Hooks()
            for name, dt in (('confidence', torch.float32), ('seeds', torch.int32), ('initial_trans', torch.float32)):
                if hooks.get(name) is not None:
                    t = hooks[name].to(device=dev, dtype=dt).contiguous()
                    keep.append(t)
                    setattr(hk, name, t.data_ptr())
        with torch.cuda.device(dev):
            _C.check(lib.eyoc_sc2pcr(_C.ptr(src), _C.ptr(tgt), _C.c_int(bs), _C.c_int(num_corr), _C.c_int(S),
                                     _C.ctypes.byref(cfg), _C.ctypes.byref(hk) if hk is not None else None,
                                     _C.ptr(ws), _C.c_size_t(ws.numel()), _C.ptr(trans), _C.ptr(fitness),
                                     _C.ptr(labels), _C.stream()))
        if detail is not None:
            detail.update(self._read_detail(ws, bs, num_corr, S, cfg))
        return trans, fitness, labels

    def _read_detail(self, ws, bs, n, S, cfg):
        """Views of the intermediate buffers inside the workspace (tests / diagnostics only)."""
        L = _C.SC2Layout()
        _C.check(_C.lib().eyoc_sc2pcr_layout(_C.c_int(bs), _C.c_int(n), _C.c_int(S), _C.ctypes.byref(cfg),
                                             _C.ctypes.byref(L)))
        I, W, k1, k2 = cfg.num_iterations, L.words_per_row, L.k1, L.k2

        def view(off, dtype, shape):
            count = int(np.prod(shape))
            size = torch.empty((), dtype=dtype).element_size()
            return ws[off:off + count * size].view(dtype).view(shape).clone()

        d = dict(
            hard_bits=view(L.hard_bits, torch.int32, (bs, n, W)), tight_bits=view(L.tight_bits, torch.int32, (bs, n, W)),
            confidence=view(L.confidence, torch.float32, (bs, n)), scores=view(L.scores, torch.float32, (bs, n)),
            seeds=view(L.seeds, torch.int32, (bs, S)), topk1=view(L.topk1, torch.int32, (bs, S, k1)),
            topk2=view(L.topk2, torch.int32, (bs, S, k2)),
            seed_weights=view(L.seed_weights, torch.float32, (bs, S, 32))[:, :, :k2],
            seed_trans=view(L.seed_trans, torch.float32, (bs, S, 4, 4)),
            global_iters=view(L.global_iters, torch.int32, (bs,)),
            local_iters=view(L.local_notclose + 4 * bs * (I + 1), torch.int32, (bs,)),
            best_seed=view(L.best_seed, torch.int32, (bs,)),
            refine_counts=view(L.refine_counts, torch.int32, (bs, cfg.refine_iterations + 1)),
            initial_trans=view(L.refine_counts + 4 * bs * (cfg.refine_iterations + 1), torch.float32, (bs, 4, 4)))
        return d

    def SC2_PCR(self, src_keypts, tgt_keypts):
        """SC2_PCR.py:307-384 -> (final_trans [bs,4,4], seedwise_fitness [bs,S])."""
        trans, fitness, _ = self._run(src_keypts, tgt_keypts, want_labels=False)
        return trans, fitness

    def estimator(self, src_keypts, tgt_keypts, src_features, tgt_features):
        """SC2_PCR.py:386-413 -> (pred_trans, pred_labels, src_corr, tgt_corr, seedwise_fitness)."""
        src_keypts_corr, tgt_keypts_corr = self.match_pair(src_keypts, tgt_keypts, src_features, tgt_features)
        pred_trans, seedwise_fitness, pred_labels = self._run(src_keypts_corr, tgt_keypts_corr, want_labels=True)
        n = pred_labels.shape[1]
        return pred_trans, pred_labels, src_keypts_corr[:, :n], tgt_keypts_corr[:, :n], seedwise_fitness
