"""Feature error of the fused forward vs the oracle for both conv data paths on a full-size synthetic cloud pair."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eyoc_b200 import nn as enn, synth
from eyoc_b200.model import load_model
from eyoc_b200.sparse import SparseTensor
from oracle import resunet_oracle as RO
torch.set_num_threads(os.cpu_count())
pair = synth.make_pair(0)
coords = synth.collate([pair['coords0'], pair['coords1']])
feats = torch.ones(len(coords), 1)
for seed in (0, 1):
    sd = RO.make_state_dict(1, 32, 5, seed=seed)
    model = load_model('ResUNetBN2C')(1, 32, bn_momentum=0.05, normalize_feature=True, conv1_kernel_size=5, D=3)
    model.load_state_dict(sd); model = model.cuda().eval()
    want = RO.resunet_forward(coords, feats, sd, True, 5)
    for mode in ('fp32', 'tf32x3'):
        enn.CONV_MODE = mode
        got = model(SparseTensor(feats.cuda(), coordinates=torch.from_numpy(coords).cuda())).F.cpu()
        err = (got - want).abs()
        print(f'seed {seed} mode {mode}: N={len(coords)} max|dF|={float(err.max()):.3e} mean|dF|={float(err.mean()):.3e} min cos={float((got*want).sum(1).min()):.9f}')
