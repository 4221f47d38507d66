"""Repeat the tensor-core pre-filtered 1-NN on the same inputs and compare with the fp32-FMA kernel every time."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eyoc_b200 import synth  # noqa: E402
from eyoc_b200.lib import eval as ev  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    pairs = synth.make_pairs([1000, 1001, 1002, 1003])
    rng = np.random.default_rng(0)
    q, r = [], []
    for p in pairs:
        q.append(p['desc0'][rng.integers(0, len(p['desc0']), 8000)])
        r.append(p['desc1'][rng.integers(0, len(p['desc1']), 8000)])
    q, r = torch.from_numpy(np.stack(q)).to(dev), torch.from_numpy(np.stack(r)).to(dev)
    for B in (2, 4):
        for form in (1, 0):
            ev.KNN_MODE = 'fp32'
            want = ev.knn1(q[:B], r[:B], form=form)
            ev.KNN_MODE = 'tc'
            bad = 0
            for it in range(40):
                got = ev.knn1(q[:B], r[:B], form=form)
                n = int((got != want).sum())
                if n:
                    bad += 1
                    rows = (got != want).nonzero()[:3].tolist()
                    print(f'B {B} form {form} iteration {it}: {n} indices differ, e.g. {rows}')
            print(f'B {B} form {form}: {bad} of 40 runs differ from the fp32-FMA kernel')


if __name__ == '__main__':
    main()
