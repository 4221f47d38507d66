"""Timeline of pipeline.OverlappedRunner on one benchmark block repeated: per block, when (ms from the first event) stage 1
(prepare), stage 2 (convolutions) and stage 3 (matching) start and end on their streams."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from eyoc_b200 import synth  # noqa: E402
from eyoc_b200.pipeline import OverlappedRunner, RegistrationPipeline, plan_to_device  # noqa: E402
from eyoc_b200.scripts.SC2_PCR.SC2_PCR import Matcher  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    pairs = synth.make_pairs(list(range(P)))
    coords, xyz, desc, sizes = synth.collate_pairs(pairs)
    model = bench.build_model(dev)
    pipe = RegistrationPipeline(model, Matcher(**bench.KITTI_CFG))
    c, x, d = torch.from_numpy(coords).to(dev), torch.from_numpy(xyz).to(dev), torch.from_numpy(desc).to(dev)
    np.random.seed(0)
    plan = plan_to_device(pipe.plan(sizes), dev)
    runner = OverlappedRunner(pipe, dev)
    for _ in range(3):
        runner.submit(c, x, sizes, plan=plan, descriptors=d)
    runner.flush()
    torch.cuda.synchronize()
    runner.trace = []
    for _ in range(5):
        runner.submit(c, x, sizes, plan=plan, descriptors=d)
    runner.flush()
    torch.cuda.synchronize()
    t0 = runner.trace[0]['prep_start']
    for i, tr in enumerate(runner.trace):
        print(f'block {i}: ' + '  '.join(f'{k} {t0.elapsed_time(tr[k]):7.2f}' for k in ('prep_start', 'prep_end', 'conv_start', 'conv_end', 'match_start', 'match_end') if k in tr))


if __name__ == '__main__':
    main()
