"""Repeat the tensor-core pre-filtered 1-NN against the fp32-FMA kernel on the same data (single pair and batched): every run
must return identical indices.  python tools/knn_stress.py"""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eyoc_b200.lib import eval as ev
g = torch.Generator().manual_seed(5)
def _norm(x): return x / x.norm(dim=-1, keepdim=True)
for n, form in ((5000, 0), (8000, 1)):
    F0, F1 = _norm(torch.randn(n, 32, generator=g)).cuda(), _norm(torch.randn(n, 32, generator=g)).cuda()
    ev.KNN_MODE = 'fp32'; ref = ev.knn1(F0, F1, form=form)
    ev.KNN_MODE = 'tc'
    bad = 0
    for it in range(60):
        got = ev.knn1(F0, F1, form=form)
        d = int((got != ref).sum())
        if d: bad += 1; print('form', form, 'iter', it, 'mismatches', d, (got != ref).nonzero().flatten()[:8].tolist())
    print('form', form, 'n', n, 'bad runs', bad, 'of 60')
# batched too
F0, F1 = _norm(torch.randn(16, 8000, 32, generator=g)).cuda(), _norm(torch.randn(16, 8000, 32, generator=g)).cuda()
ev.KNN_MODE = 'fp32'; ref = ev.knn1(F0, F1, form=1)
ev.KNN_MODE = 'tc'
bad = 0
for it in range(20):
    got = ev.knn1(F0, F1, form=1)
    d = int((got != ref).sum())
    if d: bad += 1; print('batched iter', it, 'mismatches', d)
print('batched bad runs', bad, 'of 20')
