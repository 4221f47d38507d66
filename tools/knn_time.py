"""Time the 1-NN kernels at the benchmark shapes (64 pairs: 5000 x 5000 form 0, 8000 x 8000 form 1), both data paths."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eyoc_b200.lib import eval as ev  # noqa: E402

g = torch.Generator().manual_seed(0)
for form, n in ((0, 5000), (1, 8000)):
    q = torch.randn(64, n, 32, generator=g); q = (q / q.norm(dim=-1, keepdim=True)).cuda()
    r = torch.randn(64, n, 32, generator=g); r = (r / r.norm(dim=-1, keepdim=True)).cuda()
    for mode in ('fp32', 'tc'):
        ev.KNN_MODE = mode
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); idx = ev.knn1(q, r, form=form); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f'form {form} 64 x {n} x {n} mode {mode:5s}: {min(ts[1:]):8.3f} ms', flush=True)
