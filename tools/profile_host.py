"""cProfile of the launching thread: where the host time of one block (pipe.run + records) goes.

    python tools/profile_host.py [--pairs 64]
"""
import argparse
import cProfile
import io
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from eyoc_b200 import synth  # noqa: E402
from eyoc_b200.pipeline import RegistrationPipeline, plan_to_device  # noqa: E402
from eyoc_b200.scripts.SC2_PCR.SC2_PCR import Matcher  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=64)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    pairs = synth.make_pairs(list(range(args.pairs)))
    coords_np, xyz_np, desc_np, sizes = synth.collate_pairs(pairs)
    coords_d, xyz_d, desc_d = (torch.from_numpy(a).to(dev) for a in (coords_np, xyz_np, desc_np))
    pipe = RegistrationPipeline(bench.build_model(dev), Matcher(**bench.KITTI_CFG))
    np.random.seed(0)
    plan_d = plan_to_device(pipe.plan(sizes), dev)
    ids = torch.arange(args.pairs, dtype=torch.float32, device=dev)
    for _ in range(3):
        pipe.records(pipe.run(coords_d, xyz_d, sizes, plan=plan_d, descriptors=desc_d), ids)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        pipe.records(pipe.run(coords_d, xyz_d, sizes, plan=plan_d, descriptors=desc_d), ids)
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28)
    print(s.getvalue())
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(30)
    print(s.getvalue())


if __name__ == '__main__':
    main()
