"""Batched registration pipeline: the per-pair loop of the reference's scripts/test_kitti.py:130-181 executed for
a whole block of independent pairs at once, and sharded over GPUs by pair.

Per pair the reference does: 2 x ResUNetBN2C forward -> find_corr (5000-subsample NN, diagnostic) ->
random_sample(5000) x2 -> Matcher.estimator (8000 with-replacement resample, NN matching, SC2-PCR, labels).
Here the 2P clouds of P pairs go through ONE batched forward (batch column in the coordinates), the six host
RNG draws of every pair are made up front in the reference's order (so a sequential reference loop seeded the
same way sees the same indices), the index compositions are applied by a single gather, and the matching and
SC2-PCR kernels run batched over P.
"""
import numpy as np
import torch

from . import _C
from .lib.eval import knn1
from .sparse import SparseTensor, ones_features

RECORD_FLOATS = 24     # T (16) | n_inliers | best_seed | pair_id | status | nn_hit_ratio | 3 pad


def shard_range(num_pairs, world_size, rank):
    """Contiguous block of pair ids owned by ``rank``; blocks differ by at most one pair."""
    base, rem = divmod(num_pairs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def draw_indices(n0, n1, subsample_size, num_sample, num_node):
    """The host RNG draws of one pair, in the reference's order:
    find_corr (test_kitti.py:33-34), random_sample x2 (:159-160 -> :69-71), match_pair (SC2_PCR.py:288-289)."""
    d = {}
    if subsample_size > 0 and n0 > subsample_size:
        d['fc0'] = np.random.choice(n0, min(n0, subsample_size), replace=False)
        d['fc1'] = np.random.choice(n1, min(n1, subsample_size), replace=False)
    else:
        d['fc0'], d['fc1'] = np.arange(n0), np.arange(n1)
    for key, n in (('rs0', n0), ('rs1', n1)):
        if n == num_sample:
            d[key] = np.arange(n)
        else:
            d[key] = np.random.permutation(n)[:num_sample] if n > num_sample else np.random.choice(n, num_sample)
    if num_node == 'all':
        d['mp0'], d['mp1'] = np.arange(num_sample), np.arange(num_sample)
    else:
        d['mp0'] = np.random.choice(num_sample, num_node)
        d['mp1'] = np.random.choice(num_sample, num_node)
    return d


def plan_to_device(plan, device):
    """Upload the index arrays of a plan once (benchmark 'inputs resident in HBM' mode)."""
    out = dict(plan)
    if plan['fc_uniform']:
        out['fc0'] = torch.from_numpy(np.ascontiguousarray(np.stack(plan['fc0']) if isinstance(plan['fc0'], list) else plan['fc0'])).to(device)
        out['fc1'] = torch.from_numpy(np.ascontiguousarray(np.stack(plan['fc1']) if isinstance(plan['fc1'], list) else plan['fc1'])).to(device)
    else:
        out['fc0'] = [torch.from_numpy(a).to(device) for a in plan['fc0']]
        out['fc1'] = [torch.from_numpy(a).to(device) for a in plan['fc1']]
    out['src'] = torch.from_numpy(plan['src']).to(device)
    out['tgt'] = torch.from_numpy(plan['tgt']).to(device)
    return out


def gather_rows(src, idx):
    """src [n, c] fp32 CUDA, idx int64 [...] CUDA -> src[idx] of shape [..., c] (eyoc_gather_rows: one coalesced launch)."""
    src = _C.f32c(src)
    idx = idx.contiguous()
    out = torch.empty(tuple(idx.shape) + (src.shape[1],), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _C.check(_C.lib().eyoc_gather_rows(_C.ptr(src), _C.ptr(idx), _C.c_int64(idx.numel()), _C.c_int(src.shape[1]), _C.ptr(out),
                                           _C.stream()))
    return out


class RegistrationPipeline:
    """features -> matching -> SC2-PCR for blocks of pairs on one GPU."""

    def __init__(self, model, matcher, subsample_size=5000, num_sample=5000, run_find_corr=True):
        self.model, self.matcher = model, matcher
        self.subsample_size, self.num_sample, self.run_find_corr = subsample_size, num_sample, run_find_corr

    def plan(self, sizes, fast=True):
        """Host side of one block: RNG draws + index composition.  sizes = [(n0, n1), ...] -> dict of int64 arrays
        holding GLOBAL row indices into the concatenated [cloud0_0, cloud1_0, cloud0_1, cloud1_1, ...] features.
        The draws come off the global numpy RandomState in the reference's order; ``fast`` runs them through
        eyoc_plan_draws (the same MT19937 stream and numpy's own algorithms in C, csrc/host_plan.cu) when the block
        has the fixed KITTI shape (every cloud larger than the find_corr subsample), else through numpy."""
        P = len(sizes)
        nn_ = self.matcher.num_node
        offs = np.zeros(2 * P + 1, np.int64)
        offs[1:] = np.cumsum([n for pair in sizes for n in pair])
        sub = self.subsample_size
        if fast and P and nn_ != 'all' and sub > 0 and all(n0 > sub and n1 >= sub for n0, n1 in sizes):
            return self._plan_fast(sizes, offs)
        fc0, fc1, src, tgt = [], [], [], []
        for p, (n0, n1) in enumerate(sizes):
            d = draw_indices(n0, n1, self.subsample_size, self.num_sample, nn_)
            fc0.append(d['fc0'] + offs[2 * p])
            fc1.append(d['fc1'] + offs[2 * p + 1])
            src.append(d['rs0'][d['mp0']] + offs[2 * p])
            tgt.append(d['rs1'][d['mp1']] + offs[2 * p + 1])
        same = len({len(a) for a in fc0}) == 1 and len({len(a) for a in fc1}) == 1
        return dict(offsets=offs, fc0=fc0, fc1=fc1, src=np.stack(src), tgt=np.stack(tgt), fc_uniform=same)

    def _plan_fast(self, sizes, offs):
        import ctypes
        P, sub, ns, nn_ = len(sizes), int(self.subsample_size), int(self.num_sample), int(self.matcher.num_node)
        name, key, pos, has_gauss, cached = np.random.get_state()
        if name != 'MT19937':
            raise RuntimeError('the global numpy RandomState is not MT19937')
        key = np.ascontiguousarray(key, dtype=np.uint32).copy()
        pos_c = ctypes.c_int32(int(pos))
        n0 = np.ascontiguousarray([a for a, _ in sizes], dtype=np.int64)
        n1 = np.ascontiguousarray([b for _, b in sizes], dtype=np.int64)
        fc0 = np.empty((P, sub), np.int64) if self.run_find_corr else None
        fc1 = np.empty((P, sub), np.int64) if self.run_find_corr else None
        src, tgt = np.empty((P, nn_), np.int64), np.empty((P, nn_), np.int64)

        def vp(a):
            return a.ctypes.data_as(ctypes.c_void_p) if a is not None else ctypes.c_void_p(0)

        if not self.run_find_corr:
            # find_corr is skipped by this pipeline, but its two draws still advance the reference's stream
            fc0, fc1 = np.empty((P, sub), np.int64), np.empty((P, sub), np.int64)
        _C.check(_C.lib().eyoc_plan_draws(vp(key), ctypes.byref(pos_c), ctypes.c_int(P), vp(n0), vp(n1), vp(offs),
                                          ctypes.c_int(sub), ctypes.c_int(ns), ctypes.c_int(nn_), vp(fc0), vp(fc1), vp(src), vp(tgt)))
        np.random.set_state((name, key, int(pos_c.value), has_gauss, cached))
        return dict(offsets=offs, fc0=fc0, fc1=fc1, src=src, tgt=tgt, fc_uniform=True)      # fc0 / fc1 [P, subsample] arrays

    def run(self, coords, xyz, sizes, plan=None, descriptors=None):
        """coords [sum N, 4] int32 (batch column = cloud id 0..2P-1), xyz [sum N, 3] fp32, both CUDA, clouds
        ordered [c0 of pair 0, c1 of pair 0, c0 of pair 1, ...].  Returns a dict of CUDA tensors.
        ``descriptors`` (optional [sum N, 32]) replaces the network output downstream of the forward pass (used
        by the benchmark to give the estimator a realistic inlier ratio with random-init weights); the forward
        pass is executed regardless."""
        plan = plan or self.plan(sizes)
        feats_in = ones_features(coords.shape[0], coords.device)   # lib/data_loaders.py:971-972: occupancy-only input
        F = self.model(SparseTensor(feats_in, coordinates=coords)).F
        return self.match(F, xyz, sizes, plan, descriptors)

    # ---- the same path in three stages (OverlappedRunner runs stage 1 of block k + 1 beside stage 3 of block k)
    def prepare(self, coords):
        """Stage 1: coordinate sets, first convolution (with its neighbour search), kernel maps and tile orders - the part of a
        block that is bound by dependent-access latency, not by bandwidth or math."""
        x = SparseTensor(ones_features(coords.shape[0], coords.device), coordinates=coords)
        return x, self.model.prepare(x)

    def forward(self, prepared):
        """Stage 2: the tensor-core convolutions.  -> (features [sum N, 32] fp32, pending range-flag read or None, x)."""
        x, y1 = prepared
        out, read = self.model.trunk(x, y1)
        return out.F, read, x

    def match(self, F, xyz, sizes, plan, descriptors=None):
        """Stage 3: find_corr NN, resampling, match_pair NN, SC2-PCR, inlier labels on the features of stage 2."""
        dev = F.device

        def dv(a):
            return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)

        Fm = F if descriptors is None else descriptors
        out = {'features': F}
        if self.run_find_corr:                                   # diagnostic NN of the reference (test_kitti.py:153-154)
            if plan['fc_uniform']:
                i0 = dv(plan['fc0'] if isinstance(plan['fc0'], (torch.Tensor, np.ndarray)) else np.stack(plan['fc0']))
                i1 = dv(plan['fc1'] if isinstance(plan['fc1'], (torch.Tensor, np.ndarray)) else np.stack(plan['fc1']))
                nn_idx = knn1(gather_rows(Fm, i0), gather_rows(Fm, i1), form=0)
                out['find_corr_src'] = i0
                out['find_corr_tgt'] = torch.gather(i1, 1, nn_idx)
            else:
                s_, t_ = [], []
                for a, b in zip(plan['fc0'], plan['fc1']):
                    a, b = dv(a), dv(b)
                    s_.append(a)
                    t_.append(b[knn1(Fm[a], Fm[b], form=0)])
                out['find_corr_src'], out['find_corr_tgt'] = s_, t_
        src_idx, tgt_idx = dv(plan['src']), dv(plan['tgt'])      # [P, num_node]
        nn_idx = knn1(gather_rows(Fm, src_idx), gather_rows(Fm, tgt_idx), form=1)          # SC2_PCR.py:296-298
        corr_tgt = torch.gather(tgt_idx, 1, nn_idx)
        src_corr, tgt_corr = gather_rows(xyz, src_idx), gather_rows(xyz, corr_tgt)
        trans, fitness, labels = self.matcher._run(src_corr, tgt_corr, want_labels=True)
        out.update(trans=trans, labels=labels, fitness=fitness, src_corr_idx=src_idx, tgt_corr_idx=corr_tgt,
                   src_corr=src_corr, tgt_corr=tgt_corr)
        return out

    @staticmethod
    def records(out, pair_ids):
        """Fixed-size per-pair result records [P, RECORD_FLOATS] fp32 (what the single NCCL all-gather carries)."""
        P = out['trans'].shape[0]
        rec = torch.zeros((P, RECORD_FLOATS), dtype=torch.float32, device=out['trans'].device)
        rec[:, :16] = out['trans'].reshape(P, 16)
        rec[:, 16] = out['labels'].sum(1)
        rec[:, 17] = out['fitness'].argmax(1).float()
        # (a pageable host -> device copy would synchronise the stream: pinned + non_blocking keeps the host running ahead)
        ids = pair_ids if isinstance(pair_ids, torch.Tensor) else torch.tensor(list(pair_ids), dtype=torch.float32).pin_memory()
        rec[:, 18] = ids.to(device=rec.device, dtype=torch.float32, non_blocking=True)
        rec[:, 19] = 1.0
        return rec


class OverlappedRunner:
    """Software pipeline over consecutive blocks on two CUDA streams.  Stage 3 of block k (nearest neighbours + SC2-PCR:
    compute-bound kernels) runs on ``match_stream`` while stage 1 of block k + 1 (coordinate sets, kernel maps, tile orders:
    small kernels bound by dependent-access latency) runs on the caller's stream; the two fill each other's idle issue slots.
    Stage 2 (the tensor-core convolutions, bandwidth-bound) always has the GPU to itself: it is queued behind an event that
    stage 3 of the previous block records.  ``submit`` returns the results of the PREVIOUS block; ``flush`` those of the last.

        runner = OverlappedRunner(pipe)
        for blk in blocks:
            out = runner.submit(**blk)          # None for the first block
        out = runner.flush()

    ``before_match`` / ``after_match`` (optional callables) run inside the match stream's context right before / after
    stage 3 is queued - the place to make that stream wait for uploads, to build the result records, and to release input
    buffers.  ``after_match(out)`` may return a replacement for ``out``."""

    def __init__(self, pipe, device=None):
        self.pipe = pipe
        # stages 1 + 2 on a HIGH-priority stream: while a stage-3 kernel with tens of thousands of CTAs is running, the CTAs of
        # the small stage-1 kernels are dispatched first whenever an SM has room (same-priority kernels would queue behind it)
        self.main_stream = torch.cuda.Stream(device=device, priority=-1)
        self.match_stream = torch.cuda.Stream(device=device)
        self._pending = None
        self._match_done = None
        self.trace = None            # set to a list: per block a dict of timing events (stage boundaries on both streams)

    def _ev(self, stream):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    def _match_pending(self):
        p, self._pending = self._pending, None
        if p is None:
            return None
        with torch.cuda.stream(self.match_stream):
            self.match_stream.wait_event(p['inputs_ready'])
            self.match_stream.wait_event(p['conv_done'])
            if p['before_match'] is not None:
                p['before_match']()
            if p['trace'] is not None:
                p['trace']['match_start'] = self._ev(self.match_stream)
            out = self.pipe.match(p['F'], p['xyz'], p['sizes'], p['plan'], p['descriptors'])
            if p['after_match'] is not None:
                r = p['after_match'](out)
                out = out if r is None else r
            self._match_done = torch.cuda.Event()
            self._match_done.record(self.match_stream)
            if p['trace'] is not None:
                p['trace']['match_end'] = self._ev(self.match_stream)
        p['F'].record_stream(self.match_stream)                  # allocated on the main stream, read on the match stream
        return out, p

    def _resolve_range(self, out, p):
        """The fp16 range flag of the block whose matching was just queued: its convolutions finished long ago, so the read
        does not wait.  A flagged block is redone synchronously through the fp32-activation path (rare, not overlapped)."""
        if p['read'] is not None and int(p['read'].get()[0]) & 1:
            torch.cuda.current_stream().wait_stream(self.match_stream)
            F = self.pipe.model.forward_fp32_activations(p['x']).F
            redo = self.pipe.match(F, p['xyz'], p['sizes'], p['plan'], p['descriptors'])
            if p['after_match'] is not None:                     # the hooks see the recomputed block as well
                r = p['after_match'](redo)
                redo = redo if r is None else r
            return redo
        return out

    def submit(self, coords, xyz, sizes, plan=None, descriptors=None, before_match=None, after_match=None):
        """Queue stage 3 of the previous block and stages 1 + 2 of this one.  Returns the previous block's result dict (None for
        the first block); its tensors are produced on ``match_stream`` - consume them in ``after_match``, or after
        ``match_stream.synchronize()`` / ``flush()``."""
        ready = torch.cuda.Event()
        ready.record()                                           # the inputs are complete on the caller's stream
        prev = self._match_pending()                             # stage 3 of the previous block -> match stream
        plan = plan or self.pipe.plan(sizes)
        out = None
        with torch.cuda.stream(self.main_stream):
            self.main_stream.wait_event(ready)
            tr = None
            if self.trace is not None:
                tr = dict(prep_start=self._ev(self.main_stream))
                self.trace.append(tr)
            prepared = self.pipe.prepare(coords)                 # stage 1 of this block, beside it
            if tr is not None:
                tr['prep_end'] = self._ev(self.main_stream)
            if prev is not None:
                out = self._resolve_range(*prev)
            if self._match_done is not None:
                self.main_stream.wait_event(self._match_done)    # stage 2 runs alone
            if tr is not None:
                tr['conv_start'] = self._ev(self.main_stream)
            F, read, x = self.pipe.forward(prepared)
            done = torch.cuda.Event()
            done.record(self.main_stream)
            if tr is not None:
                tr['conv_end'] = self._ev(self.main_stream)
        for t in (coords, xyz, descriptors):
            if isinstance(t, torch.Tensor):
                t.record_stream(self.main_stream)
                t.record_stream(self.match_stream)
        self._pending = dict(F=F, read=read, x=x, xyz=xyz, sizes=sizes, plan=plan, descriptors=descriptors, conv_done=done,
                             inputs_ready=ready, before_match=before_match, after_match=after_match, trace=tr)
        return out

    def flush(self):
        """Queue stage 3 of the last block; the caller's stream then waits for both streams of the runner."""
        prev = self._match_pending()
        out = None
        if prev is not None:
            with torch.cuda.stream(self.main_stream):
                out = self._resolve_range(*prev)
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.match_stream)
        cur.wait_stream(self.main_stream)
        return out


class PlanPrefetcher:
    """Runs ``pipe.plan(sizes)`` (the host RNG draws + index composition of the next block) on a worker thread so it
    overlaps the GPU work of the current block.  Plans come out in the order they were drawn, so the global numpy
    RNG stream is consumed exactly as by a sequential loop."""

    def __init__(self, pipe, sizes_iter, depth=2):
        import queue
        import threading
        self.q = queue.Queue(maxsize=depth)
        self._stop = False

        def work():
            for sizes in sizes_iter:
                if self._stop:
                    break
                self.q.put(pipe.plan(sizes))
            self.q.put(None)
        self.t = threading.Thread(target=work, daemon=True)
        self.t.start()

    def get(self):
        return self.q.get()

    def close(self):
        self._stop = True
        try:
            while self.q.get_nowait() is not None:
                pass
        except Exception:
            pass


class BlockUploader:
    """Host -> device copies of a block's inputs (coordinates, points, index plan) on a dedicated copy stream, one block
    ahead of the compute stream: ``start(...)`` queues the copies of block i + 1 while block i computes, ``ticket.wait()``
    makes the compute stream wait for them (an event, no host synchronisation), ``ticket.release()`` - called once the
    block's kernels are queued - lets the uploader reuse the buffers two blocks later.  Pinned staging buffers and device
    buffers are owned by the uploader (two sets, used alternately), so the steady state allocates nothing."""

    class _Ticket:
        def __init__(self, owner, slot, tensors, event):
            self.owner, self.slot, self.tensors, self.event = owner, slot, tensors, event

        def wait(self):
            torch.cuda.current_stream().wait_event(self.event)
            return self.tensors

        def release(self):
            ev = torch.cuda.Event()
            ev.record()                                     # on the compute stream: everything that reads the buffers is queued
            self.owner._free[self.slot] = ev

    def __init__(self, device, slots=2):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.slots = slots
        self._pinned = [{} for _ in range(slots)]
        self._dev = [{} for _ in range(slots)]
        self._free = [None] * slots       # compute-stream events: the device buffers of a set may be overwritten
        self._copied = [None] * slots     # copy-stream events: the pinned staging buffers of a set may be overwritten
        self._flip = 0
        self.bytes_last = 0

    @staticmethod
    def _buf(pool, key, like, make):
        buf = pool.get(key)
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = pool[key] = make()
        return buf

    def start(self, **host):
        """host: name -> numpy array / CPU tensor (or a list of equally shaped arrays, stacked).  Returns a ticket."""
        slot = self._flip
        self._flip = (self._flip + 1) % self.slots
        pinned, devp = self._pinned[slot], self._dev[slot]
        out, nbytes = {}, 0
        with torch.cuda.stream(self.stream):
            if self._free[slot] is not None:
                self.stream.wait_event(self._free[slot])    # the block that last used this set has been consumed
            for k, a in host.items():
                if a is None:
                    out[k] = None
                    continue
                if isinstance(a, (list, tuple)):
                    a = np.stack(a)
                t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
                if not t.is_pinned():
                    stage = self._buf(pinned, k, t, lambda: torch.empty(t.shape, dtype=t.dtype).pin_memory())
                    if self._copied[slot] is not None:
                        self._copied[slot].synchronize()    # this set's previous H2D (two blocks ago) has long completed
                    stage.copy_(t)
                    t = stage
                dst = self._buf(devp, k, t, lambda: torch.empty(t.shape, dtype=t.dtype, device=self.device))
                dst.copy_(t, non_blocking=True)
                out[k] = dst
                nbytes += t.numel() * t.element_size()
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._copied[slot] = ev
        self.bytes_last = nbytes
        return BlockUploader._Ticket(self, slot, out, ev)


class BlockFeeder:
    """Everything the host does for a block BEFORE its kernels are launched - the RNG index plan (``pipe.plan``), the
    pinned staging of the plan arrays and the host -> device copies of coordinates / points / plan on the copy stream - on a
    worker thread, ``depth`` blocks ahead of the compute loop.  ``get()`` returns (ticket, plan) in block order (the global
    numpy RNG stream is consumed exactly as by a sequential loop); the main thread only launches kernels.
    ``blocks`` yields dicts with ``coords`` / ``xyz`` (pinned CPU tensors or numpy arrays) and ``sizes``."""

    def __init__(self, pipe, blocks, device, depth=2):
        import queue
        import threading
        self.q = queue.Queue(maxsize=depth)
        self.uploader = BlockUploader(device, slots=depth + 2)      # a set is reused only after its block was consumed
        self._stop = False

        def work():
            torch.cuda.set_device(device)
            for blk in blocks:
                if self._stop:
                    break
                pl = pipe.plan(blk['sizes'])
                ticket = self.uploader.start(coords=blk['coords'], xyz=blk['xyz'], fc0=pl['fc0'], fc1=pl['fc1'], src=pl['src'],
                                             tgt=pl['tgt'])
                self.q.put((ticket, pl))
            self.q.put(None)
        self.t = threading.Thread(target=work, daemon=True)
        self.t.start()

    def get(self):
        return self.q.get()

    def close(self):
        self._stop = True
        try:
            while self.q.get_nowait() is not None:
                pass
        except Exception:      # noqa: BLE001
            pass


class AsyncRecords:
    """The per-block result records on their way to the host: the all-gather runs asynchronously (NCCL's own stream) and
    the device-to-host copy on a side stream, so the compute stream goes straight on to the next block; ``result()`` -
    normally called one block later - blocks the HOST only, on the copy's event."""

    def __init__(self, rec, num_pairs, group=None, host_out=None):
        import torch.distributed as dist
        self.num_pairs, self.group = num_pairs, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.work = None
        if self.world > 1:
            self.m = -(-num_pairs // self.world)
            pad = torch.zeros((self.m, rec.shape[1]), dtype=rec.dtype, device=rec.device)
            pad[:rec.shape[0]] = rec
            self.buf = torch.empty((self.world * self.m, rec.shape[1]), dtype=rec.dtype, device=rec.device)
            self.pad = pad
            self.work = dist.all_gather_into_tensor(self.buf, pad, group=group, async_op=True)
        else:
            self.buf = rec
        self.side = _side_stream(rec.device)
        self.host = host_out if host_out is not None else torch.empty(self.buf.shape, dtype=self.buf.dtype).pin_memory()
        ready = torch.cuda.Event()
        ready.record()                                        # rec (and, at world 1, buf) complete on the compute stream
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            if self.work is not None:
                self.work.wait()                              # makes the SIDE stream wait for the collective
            self.host[:self.buf.shape[0]].copy_(self.buf, non_blocking=True)
            self.buf.record_stream(self.side)
            self.done = torch.cuda.Event()
            self.done.record(self.side)

    def result(self):
        """[num_pairs, RECORD_FLOATS] pinned CPU tensor (padding of uneven shards stripped)."""
        self.done.synchronize()
        if self.world == 1:
            return self.host[:self.buf.shape[0]]
        keep = []
        for r in range(self.world):
            lo, hi = shard_range(self.num_pairs, self.world, r)
            keep.append(self.host[r * self.m: r * self.m + (hi - lo)])
        return torch.cat(keep, 0)


_SIDE = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def gather_records(rec, num_pairs, group=None, counts=None):
    """The path's only collective: ONE all-gather of the per-pair records.  Every rank pads its block to the largest
    block (block sizes follow from ``shard_range`` - or from ``counts``, the pairs each rank holds, when the caller shards
    differently - so no size exchange is needed); the padding is stripped after the gather.  NCCL on GPUs, gloo in the
    CPU tests."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rec
    world = dist.get_world_size(group)
    if counts is None:
        counts = [shard_range(num_pairs, world, r)[1] - shard_range(num_pairs, world, r)[0] for r in range(world)]
    m = max(max(counts), 1)
    pad = torch.zeros((m, rec.shape[1]), dtype=rec.dtype, device=rec.device)
    pad[:rec.shape[0]] = rec
    buf = torch.empty((world * m, rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    return torch.cat([buf[r * m: r * m + counts[r]] for r in range(world)], 0)
