"""BASELINE.json configs[4]: LoKITTI-shaped distant-pair sweep on synthetic pairs.

545 pairs (the shape of config/file_LoKITTI_50.npy), ids 1000..1544, ground-truth translation d ~ U[5, 50] m, evaluated through
the batched pipeline in blocks and reported in the reference's five distance buckets (scripts/test_kitti.sh:45) with the
reference's RTE / RRE / success formulas (scripts/test_kitti.py:188-210); a subsample is re-run through the CPU oracle for
parity (identical correspondence sets and inlier masks, pose within 1e-4 / 1e-3).

    python tools/lokitti_sweep.py [--pairs 545] [--block 64] [--oracle-pairs 6]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/lokitti_sweep.py --oracle-pairs 16

Under torchrun the 545 pairs are sharded over the ranks by contiguous blocks (pipeline.shard_range), every rank runs
its shard in blocks, the per-pair records travel through the path's one collective (pipeline.gather_records) and rank 0 prints the
bucket table - the pair-sharded 8 x B200 form of scripts/test_kitti.sh:45-75.  The oracle parity subsample runs on rank 0.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eyoc_b200 import synth  # noqa: E402
from eyoc_b200.pipeline import RegistrationPipeline  # noqa: E402
from eyoc_b200.scripts import test_kitti as tk  # noqa: E402

BUCKETS = [(5, 10), (10, 20), (20, 30), (30, 40), (40, 50)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=545)
    ap.add_argument('--block', type=int, default=64)
    ap.add_argument('--oracle-pairs', type=int, default=6)
    ap.add_argument('--out', default=None)
    ap.add_argument('--dump', default=None, help='write the gathered record table (.npy)')
    args = ap.parse_args()
    import torch.distributed as dist
    from eyoc_b200.pipeline import gather_records, shard_range
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if os.environ.get('EYOC_SWEEP_POWER_STEPWISE'):       # diagnostic: the per-iteration power-iteration kernel
        from eyoc_b200 import _C
        _C.lib().eyoc_debug_sc2_power_fused(0)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cfg = tk.make_config(tk.parse_args(['--use_RANSAC', 'false']))
    model = tk.build_model(cfg, dev)
    matcher = tk._matcher(cfg)
    pipe = RegistrationPipeline(model, matcher)
    all_ids = list(range(1000, 1000 + args.pairs))
    # shard by whole blocks, so that a pair sits in the same block - and sees the same RNG draws - at any world size:
    # the gathered table at world 8 is byte-identical to the world-1 table
    nblocks = -(-len(all_ids) // args.block)
    blocks_of = [shard_range(nblocks, world, r) for r in range(world)]
    counts = [min(b1 * args.block, len(all_ids)) - min(b0_ * args.block, len(all_ids)) for b0_, b1 in blocks_of]
    lo = blocks_of[rank][0] * args.block
    ids = all_ids[lo:lo + counts[rank]]
    t0 = time.time()
    rows = []
    seeds = {}
    gpu_s = 0.0
    recs = []
    for b0 in range(0, len(ids), args.block):
        blk = ids[b0:b0 + args.block]
        pairs = synth.make_pairs(blk)
        coords, xyz, desc, sizes = synth.collate_pairs(pairs)
        np.random.seed(lo + b0)                        # a function of the global position: the same draws at any world size
        seeds[b0] = blk
        torch.cuda.synchronize()
        t1 = time.time()
        out = pipe.run(torch.from_numpy(coords).to(dev), torch.from_numpy(xyz).to(dev), sizes, descriptors=torch.from_numpy(desc).to(dev))
        rec = pipe.records(out, blk)
        rec[:, 20:23] = torch.from_numpy(np.stack([p['T_gt'][:3, 3] for p in pairs])).to(dev)      # spare columns: GT translation
        rec[:, 23] = torch.tensor([tk.rte_rre(T.cpu(), torch.from_numpy(p['T_gt']))[1] for T, p in zip(out['trans'], pairs)], device=dev)
        recs.append(rec)
        torch.cuda.synchronize()
        gpu_s += time.time() - t1
    rec = torch.cat(recs, 0) if recs else torch.zeros((0, 24), device=dev)
    table_all = gather_records(rec, len(all_ids), counts=counts).cpu().numpy()   # the path's ONE collective
    t = torch.tensor([gpu_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gpu_s = float(t.item())
    if rank != 0:
        dist.destroy_process_group()
        return
    if args.dump:
        np.save(args.dump, table_all)
    for r in table_all:
        T = r[:16].reshape(4, 4)
        rte = float(np.linalg.norm(T[:3, 3] - r[20:23]))
        rre = float(r[23])
        rows.append(dict(id=int(r[18]), dist=float(np.linalg.norm(r[20:23])), rte=rte, rre_deg=float(np.degrees(rre)),
                         ok=tk.is_success(rte, rre), inliers=int(r[16])))
    assert [r['id'] for r in rows] == all_ids, 'gathered table is not in pair order'
    print(f'{len(rows)} pairs on {world} GPU(s) in {time.time() - t0:.1f} s wall (slowest rank {gpu_s:.2f} s in the pipeline incl. H2D: '
          f'{len(rows) / gpu_s:.0f} pairs/s)')
    print('bucket [m]   pairs   RR %    RTE cm (succ. mean)   RRE deg (succ. mean)   median inliers / 8000')
    table = []
    for lo, hi in BUCKETS + [(5, 50)]:
        sel = [r for r in rows if lo <= r['dist'] < hi or (hi == 50 and r['dist'] == 50)]
        if not sel:
            continue
        ok = [r for r in sel if r['ok']]
        rr = 100.0 * len(ok) / len(sel)
        rte = 100 * np.mean([r['rte'] for r in ok]) if ok else float('nan')
        rre = np.mean([r['rre_deg'] for r in ok]) if ok else float('nan')
        med = int(np.median([r['inliers'] for r in sel]))
        print(f'[{lo:2d}, {hi:2d}]    {len(sel):5d}   {rr:5.1f}   {rte:10.2f}            {rre:10.3f}             {med:6d}')
        table.append(dict(bucket=[lo, hi], pairs=len(sel), rr=rr, rte_cm=rte, rre_deg=rre, median_inliers=med))
    # ---- oracle parity on a subsample (first pairs of the first block: same numpy stream as the block run)
    parity = []
    if args.oracle_pairs > 0:
        from oracle import matching_oracle as MO, sc2pcr_oracle as O
        k = min(args.oracle_pairs, args.block, len(ids))
        blk = ids[:min(args.block, len(ids))]
        pairs = synth.make_pairs(blk)
        coords, xyz, desc, sizes = synth.collate_pairs(pairs)
        np.random.seed(0)
        out = pipe.run(torch.from_numpy(coords).to(dev), torch.from_numpy(xyz).to(dev), sizes, descriptors=torch.from_numpy(desc).to(dev))
        np.random.seed(0)
        ocfg = O.SC2Config(**{kk: tk.CONFIG_KITTI[kk] for kk in ('inlier_threshold', 'num_node', 'd_thre', 'num_iterations', 'ratio',
                                                                 'nms_radius', 'max_points', 'k1', 'k2')}, stable_ties=True)
        torch.set_num_threads(os.cpu_count() or 1)
        for j in range(k):                                 # the oracle consumes the RNG stream pair by pair, like the pipeline's plan
            p = pairs[j]
            F0, F1 = torch.from_numpy(p['desc0']), torch.from_numpy(p['desc1'])
            MO.find_corr(p['xyz0'], p['xyz1'], F0, F1, subsample_size=5000)
            x0, f0 = MO.random_sample(p['xyz0'], F0, 5000)
            x1, f1 = MO.random_sample(p['xyz1'], F1, 5000)
            T_o, lab_o, sc_o, tc_o, _ = O.estimator(torch.from_numpy(x0)[None], torch.from_numpy(x1)[None], f0[None], f1[None], ocfg,
                                                    dense_weight=False)
            T = out['trans'][j].cpu()
            rec = dict(id=blk[j], corr_equal=bool(torch.equal(out['src_corr'][j].cpu(), sc_o[0]) and torch.equal(out['tgt_corr'][j].cpu(), tc_o[0])),
                       mask_hamming=int((out['labels'][j].cpu() != lab_o[0]).sum()),
                       dR=float(torch.linalg.norm(T[:3, :3] - T_o[0, :3, :3])), dt=float(torch.linalg.norm(T[:3, 3] - T_o[0, :3, 3])))
            parity.append(rec)
            print('oracle parity', rec)
        assert all(r['corr_equal'] and r['mask_hamming'] == 0 and r['dR'] < 1e-4 and r['dt'] < 1e-3 for r in parity), 'parity violated'
        print(f'oracle parity OK on {k} pairs: identical correspondence sets and inlier masks, pose within 1e-4 / 1e-3')
    if args.out:
        import hashlib
        json.dump(dict(table=table, parity=parity, pairs=len(rows), pipeline_seconds=gpu_s, world=world,
                       records_sha256=hashlib.sha256(np.ascontiguousarray(table_all[:, :20]).tobytes()).hexdigest()), open(args.out, 'w'), indent=1)
        print('records sha256 (poses, inlier counts, best seeds, ids: world-size invariant):',
              hashlib.sha256(np.ascontiguousarray(table_all[:, :20]).tobytes()).hexdigest())
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
