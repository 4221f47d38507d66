"""Drop-in for the reference's scripts/SC2_PCR/SC2_PCR.py ``Matcher`` on the fused sm_100a estimator.

Same constructor and the same public entry points on the inference path:
    Matcher.match_pair   (SC2_PCR.py:280-305)
    Matcher.SC2_PCR      (SC2_PCR.py:307-384)  -> (final_trans [bs,4,4], seedwise_fitness [bs,S])
    Matcher.estimator    (SC2_PCR.py:386-413)  -> 5-tuple
and the public stage methods, on the tensors the reference hands them:
    Matcher.pick_seeds              (SC2_PCR.py:33-59)    dense dists [bs,n,n] + scores -> seeds [bs,max_num] int64
    Matcher.cal_seed_trans          (SC2_PCR.py:61-168)   seeds + dense SC2 [bs,S,n] -> (final_trans, seedwise_fitness)
    Matcher.cal_leading_eigenvector (SC2_PCR.py:170-196)  dense M [bs,n,n] -> [bs,n]   (method='power')
    Matcher.post_refinement         (SC2_PCR.py:238-278)  initial_trans -> refined trans
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
    def _cfg(self, refine_iterations=20):
        refine_thr = 0.10 if self.inlier_threshold == 0.10 else 1.2          # SC2_PCR.py:254-257
        return _C.SC2Cfg(inlier_threshold=self.inlier_threshold, d_thre=self.d_thre, d_thre_half=self.d_thre / 2,
                         d_thre_sq=self.d_thre ** 2, nms_radius=self.nms_radius, refine_threshold=refine_thr,
                         num_iterations=int(self.num_iterations), k1=int(self.k1), k2=int(self.k2),
                         refine_iterations=int(refine_iterations))

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
    def _run(self, src_keypts, tgt_keypts, want_labels, detail=None, hooks=None, refine_iterations=20, num_seeds=None):
        _C.require_cuda(src_keypts, tgt_keypts)
        bs, num_corr = src_keypts.shape[0], tgt_keypts.shape[1]
        if num_corr > self.max_points:                                     # SC2_PCR.py:324-327
            src_keypts = src_keypts[:, :self.max_points, :]
            tgt_keypts = tgt_keypts[:, :self.max_points, :]
            num_corr = self.max_points
        src, tgt = _C.f32c(src_keypts), _C.f32c(tgt_keypts)
        dev = src.device
        S = int(num_corr * self.ratio) if num_seeds is None else int(num_seeds)      # SC2_PCR.py:350
        cfg = self._cfg(refine_iterations)
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
            for name, dt in (('confidence', torch.float32), ('seeds', torch.int32), ('initial_trans', torch.float32),
                             ('sc2_dense', torch.float32)):
                if hooks.get(name) is not None:
                    t = hooks[name].to(device=dev, dtype=dt).contiguous()
                    keep.append(t)
                    setattr(hk, name, t.data_ptr())
        with torch.cuda.device(dev):
            _C.check(lib.eyoc_sc2pcr(_C.ptr(src), _C.ptr(tgt), _C.c_int(bs), _C.c_int(num_corr), _C.c_int(S),
                                     _C.ctypes.byref(cfg), _C.ctypes.byref(hk) if hk is not None else None,
                                     _C.ptr(ws), _C.c_size_t(ws.numel()), _C.ptr(trans), _C.ptr(fitness),
                                     _C.ptr(labels), _C.stream()))
        if hooks and hooks.get('sc2_dense') is not None:
            L = _C.SC2Layout()
            _C.check(lib.eyoc_sc2pcr_layout(_C.c_int(bs), _C.c_int(num_corr), _C.c_int(S), _C.ctypes.byref(cfg), _C.ctypes.byref(L)))
            if int(ws[L.status:L.status + 4].view(torch.int32).item()) & 1:
                raise RuntimeError('cal_seed_trans: SC2_measure must hold the integer-valued second-order counts '
                                   '(SC2_PCR.py:363: a product of 0/1 matrices), each in [0, 65535]')
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
            near_bits=view(L.near_bits, torch.int32, (bs, n, W)),
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

    # ------------------------------------------------------------------ public stage methods (dense tensors in and out)
    def pick_seeds(self, dists, scores, R, max_num):
        """SC2_PCR.py:33-59: NMS on the confidence within radius R of the dense distance matrix, then the max_num best
        in descending order (ties: lowest index) -> [bs, max_num] int64.  (The reference asserts bs == 1.)"""
        _C.require_cuda(dists, scores)
        dists, scores = _C.f32c(dists), _C.f32c(scores)
        bs, n = scores.shape
        if dists.shape != (bs, n, n):
            raise RuntimeError(f'pick_seeds: dists {tuple(dists.shape)} does not match scores {tuple(scores.shape)}')
        max_num = min(int(max_num), n)
        lib = _C.lib()
        ws = self._workspace(max(lib.eyoc_pick_seeds_workspace_bytes(_C.c_int(bs), _C.c_int(n)), 256), scores.device)
        seeds = torch.empty((bs, max_num), dtype=torch.int64, device=scores.device)
        with torch.cuda.device(scores.device):
            _C.check(lib.eyoc_pick_seeds_dense(_C.ptr(dists), _C.ptr(scores), _C.c_int(bs), _C.c_int(n), _C.c_float(R),
                                               _C.c_int(max_num), _C.ptr(seeds), _C.ptr(ws), _C.c_size_t(ws.numel()), _C.stream()))
        return seeds

    def cal_leading_eigenvector(self, M, method='power'):
        """SC2_PCR.py:170-196: power iteration from ones, at most num_iterations rounds, one allclose over the whole batch."""
        if method != 'power':
            # the reference's 'eig' branch calls torch.symeig, which no longer exists in the torch this image ships
            raise NotImplementedError("cal_leading_eigenvector: only method='power' (the one the reference path uses)")
        _C.require_cuda(M)
        M = _C.f32c(M)
        bs, n = M.shape[0], M.shape[1]
        if M.dim() != 3 or M.shape[2] != n:
            raise RuntimeError(f'cal_leading_eigenvector: M must be [bs, n, n], got {tuple(M.shape)}')
        lib = _C.lib()
        I = int(self.num_iterations)
        ws = self._workspace(max(lib.eyoc_power_iteration_workspace_bytes(_C.c_int(bs), _C.c_int(n), _C.c_int(I)), 256), M.device)
        v = torch.empty((bs, n), dtype=torch.float32, device=M.device)
        with torch.cuda.device(M.device):
            _C.check(lib.eyoc_power_iteration_dense(_C.ptr(M), _C.c_int(bs), _C.c_int(n), _C.c_int(I), _C.ptr(v), None,
                                                    _C.ptr(ws), _C.c_size_t(ws.numel()), _C.stream()))
        return v

    def cal_seed_trans(self, seeds, SC2_measure, src_keypts, tgt_keypts):
        """SC2_PCR.py:61-168: two-stage consensus around every seed on the caller's dense second-order measure, weighted
        Kabsch per seed, inlier counts -> (final_trans [bs,4,4] of the best seed, seedwise_fitness [bs,S])."""
        _C.require_cuda(seeds, SC2_measure)
        bs, S, n = SC2_measure.shape
        if seeds.shape != (bs, S) or src_keypts.shape[1] != n:
            raise RuntimeError('cal_seed_trans: seeds / SC2_measure / keypoints shapes do not match')
        trans, fitness, _ = self._run(src_keypts, tgt_keypts, want_labels=False, refine_iterations=0, num_seeds=S,
                                      hooks=dict(seeds=seeds, sc2_dense=SC2_measure))
        return trans, fitness

    def post_refinement(self, initial_trans, src_keypts, tgt_keypts, it_num, weights=None):
        """SC2_PCR.py:238-278 (``weights`` is accepted and ignored, as in the reference)."""
        trans, _, _ = self._run(src_keypts, tgt_keypts, want_labels=False, refine_iterations=int(it_num), num_seeds=1,
                                hooks=dict(initial_trans=initial_trans))
        return trans

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
