"""Hot-path helpers of the reference's inference driver scripts/test_kitti.py, same names and semantics:
``find_corr`` (:28-42), ``apply_transform`` (:44-47), ``evaluate_nn_dist`` (:49-52), ``random_sample`` (:54-73)
and the RTE / RRE / success formulas of ``main`` (:188-210).  The draw order on the global numpy RNG is the
reference's.

``main(config, test_loader)`` is the reference's per-pair evaluation loop (:75-233) on the drop-in symbols, fed by any
iterable that yields the reference's ``collate_pair_fn`` dictionaries (lib/data_loaders.py:31-85: pcd0/pcd1 lists,
sinput{0,1}_C [N,4] int32, sinput{0,1}_F [N,1], T_gt) - the reference's own ``make_data_loader`` plugs in unchanged;
``SyntheticPairLoader`` yields the same dictionaries from the SURVEY 8d generator and ``RawPairLoader`` from raw sweeps
voxelised on the device (eyoc_voxelize).  ``main_blocks`` runs the same evaluation through the batched pipeline
(blocks of pairs per launch sequence).  The command line keeps the reference's flags (:240-294):

    python -m eyoc_b200.scripts.test_kitti --synthetic 64 [--block 32] [--save_dir DIR] [--use_RANSAC false]

Open3D RANSAC (``--use_RANSAC true``, the reference's default) is outside the hot path and is refused.
"""
import argparse
import json
import logging
import os

import numpy as np
import torch

from ..lib.eval import find_nn_gpu
from ..lib.timer import AverageMeter, Timer

# scripts/SC2_PCR/config_json/config_KITTI.json (merged into the config when SC2-PCR is selected, :279-284)
CONFIG_KITTI = dict(inlier_threshold=0.6, num_node=8000, use_mutual=False, d_thre=0.1, num_iterations=20, ratio=0.2,
                    nms_radius=0.6, max_points=8000, k1=30, k2=20, downsample=0.3, re_thre=5, te_thre=60)
# the model / loader keys main() reads from <save_dir>/config.json, with the values scripts/train_kitti_EYOC.sh fixes
CONFIG_DEFAULTS = dict(model='ResUNetBN2C', model_n_out=32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True,
                       voxel_size=0.3, rte_thresh=2.0, rre_thresh=5.0, use_RANSAC=False, test_phase='test')


def find_corr(xyz0, xyz1, F0, F1, subsample_size=-1):
    """scripts/test_kitti.py:28-42."""
    subsample = len(F0) > subsample_size
    if subsample_size > 0 and subsample:
        N0 = min(len(F0), subsample_size)
        N1 = min(len(F1), subsample_size)
        inds0 = np.random.choice(len(F0), N0, replace=False)
        inds1 = np.random.choice(len(F1), N1, replace=False)
        F0, F1 = F0[torch.from_numpy(inds0).to(F0.device)], F1[torch.from_numpy(inds1).to(F1.device)]
    nn_inds = find_nn_gpu(F0, F1, nn_max_n=500)
    if subsample_size > 0 and subsample:
        return xyz0[inds0], xyz1[inds1[nn_inds]]
    return xyz0, xyz1[nn_inds]


def apply_transform(pts, trans):
    """scripts/test_kitti.py:44-47."""
    R = trans[:3, :3]
    T = trans[:3, 3]
    return pts @ R.t() + T


def evaluate_nn_dist(xyz0, xyz1, T_gth):
    """scripts/test_kitti.py:49-52."""
    xyz0 = apply_transform(xyz0, T_gth)
    dist = np.sqrt(((xyz0 - xyz1) ** 2).sum(1) + 1e-6)
    return dist.tolist()


def random_sample(pcd, feats, N):
    """scripts/test_kitti.py:54-73: exactly N points (permutation if more, with replacement if fewer)."""
    n1 = pcd.size(0) if isinstance(pcd, torch.Tensor) else pcd.shape[0]
    if n1 == N:
        return pcd, feats
    choice = np.random.permutation(n1)[:N] if n1 > N else np.random.choice(n1, N)
    fidx = torch.from_numpy(choice).to(feats.device) if isinstance(feats, torch.Tensor) else choice
    return pcd[choice], feats[fidx]


def rte_rre(T_est, T_gth):
    """scripts/test_kitti.py:188-191 (CPU torch [4,4] in, python floats out), including the diag clamp."""
    rte = np.linalg.norm(T_est[:3, 3] - T_gth[:3, 3])
    trace_matrix = T_est[:3, :3].t() @ T_gth[:3, :3]
    trace_matrix[[0, 1, 2], [0, 1, 2]] = torch.min(torch.ones(3), trace_matrix[[0, 1, 2], [0, 1, 2]])
    rre = np.arccos((np.trace(trace_matrix) - 1) / 2)
    return float(rte), float(rre)


def is_success(rte, rre, rte_thresh=2.0, rre_thresh=5.0):
    """scripts/test_kitti.py:206 (defaults :253-255)."""
    return bool(rte < rte_thresh and not np.isnan(rre) and rre < np.pi / 180 * rre_thresh)


# ------------------------------------------------------------------------------------------------ data sources
class SyntheticPairLoader:
    """Yields the reference's collate_pair_fn dictionaries (batch size 1) from the synthetic KITTI-shaped generator;
    ``planted`` adds ``desc0`` / ``desc1`` (descriptors with a known inlier ratio: random-init weights carry no signal)."""

    def __init__(self, pair_ids, planted=True):
        from .. import synth
        self.pairs = synth.make_pairs(list(pair_ids))
        self.planted = planted

    def __len__(self):
        return len(self.pairs)

    def __iter__(self):
        from .. import synth
        for p in self.pairs:
            d = {'pcd0': [torch.from_numpy(p['xyz0'])], 'pcd1': [torch.from_numpy(p['xyz1'])],
                 'sinput0_C': torch.from_numpy(synth.collate([p['coords0']])), 'sinput1_C': torch.from_numpy(synth.collate([p['coords1']])),
                 'sinput0_F': torch.ones((len(p['coords0']), 1)), 'sinput1_F': torch.ones((len(p['coords1']), 1)),
                 'T_gt': [torch.from_numpy(p['T_gt'].astype(np.float32))], 'len_batch': [[len(p['coords0']), len(p['coords1'])]]}
            if self.planted and 'desc0' in p:
                d['desc0'], d['desc1'] = torch.from_numpy(p['desc0']), torch.from_numpy(p['desc1'])
            yield d


class RawPairLoader:
    """Raw sweeps -> the same dictionaries, voxelised on the device exactly like the reference's loader
    (lib/data_loaders.py:936-979: sel = sparse_quantize(xyz / voxel), coords = floor(xyz[sel] / voxel), feats = 1).
    ``pairs`` = iterable of (xyz0 [n0,3] fp32, xyz1 [n1,3] fp32, T_gt [4,4])."""

    def __init__(self, pairs, voxel_size=0.3, device='cuda'):
        self.pairs, self.voxel_size, self.device = list(pairs), voxel_size, torch.device(device)

    def __len__(self):
        return len(self.pairs)

    def __iter__(self):
        from ..sparse import ones_features, voxelize_gpu
        for xyz0, xyz1, T in self.pairs:
            out = {}
            for i, xyz in enumerate((xyz0, xyz1)):
                x = torch.as_tensor(xyz, dtype=torch.float32).to(self.device)
                coords, sel = voxelize_gpu(x, self.voxel_size)
                out[f'pcd{i}'] = [x[sel].cpu()]
                out[f'sinput{i}_C'] = coords
                out[f'sinput{i}_F'] = ones_features(coords.shape[0], self.device)
            out['T_gt'] = [torch.as_tensor(T, dtype=torch.float32)]
            yield out


class _Cfg(dict):
    """dict with attribute access (the reference uses easydict)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def build_model(config, device):
    """scripts/test_kitti.py:81-93.  With ``--save_dir`` the checkpoint must exist (the reference raises inside torch.load;
    a mistyped directory must not turn into an evaluation of random weights).  Without a save_dir - the synthetic mode,
    where no checkpoint is reachable - the weights are seeded random-init, and the log says so."""
    from ..model import load_model
    Model = load_model(config.model)
    model = Model(1, config.model_n_out, bn_momentum=config.bn_momentum, conv1_kernel_size=config.conv1_kernel_size,
                  normalize_feature=config.normalize_feature)
    ckpt = os.path.join(config.save_dir, 'best_val_checkpoint.pth') if config.get('save_dir') else None
    if ckpt is not None:
        if not os.path.exists(ckpt):
            raise FileNotFoundError(f'{ckpt}: no checkpoint under --save_dir (omit --save_dir to evaluate seeded random-init '
                                    'weights on synthetic pairs)')
        model.load_state_dict(torch.load(ckpt, map_location='cpu')['state_dict'])
    else:
        logging.warning('no --save_dir: seeded RANDOM-INIT weights (recall figures carry no meaning)')
        torch.manual_seed(0)
        model.apply(lambda m: m.reset_parameters() if hasattr(m, 'reset_parameters') and not isinstance(m, torch.nn.BatchNorm1d) else None)
    return model.to(device).eval()


def _matcher(config):
    from .SC2_PCR.SC2_PCR import Matcher
    return Matcher(inlier_threshold=config.inlier_threshold, num_node=config.num_node, use_mutual=config.use_mutual,
                   d_thre=config.d_thre, num_iterations=config.num_iterations, ratio=config.ratio, nms_radius=config.nms_radius,
                   max_points=config.max_points, k1=config.k1, k2=config.k2)


class _Stats:
    """The meters and the final report of the reference's main (:107-128, :188-233)."""

    def __init__(self, config):
        self.success, self.rte, self.rre = AverageMeter(), AverageMeter(), AverageMeter()
        self.rte_thresh, self.rre_thresh = config.rte_thresh, config.rre_thresh
        self.list_rte, self.list_rre, self.T_est, self.max_dist = [], [], [], 0.0

    def add(self, T_est, T_gth):
        rte, rre = rte_rre(T_est, T_gth)
        self.max_dist = max(self.max_dist, float(np.linalg.norm(T_gth[:3, 3])))
        if rte < self.rte_thresh:
            self.rte.update(rte)
        if not np.isnan(rre) and rre < np.pi / 180 * self.rre_thresh:
            self.rre.update(rre * 180 / np.pi)
        ok = is_success(rte, rre, self.rte_thresh, self.rre_thresh)
        self.success.update(1 if ok else 0)
        if not ok:
            logging.info(f"Failed with RTE: {rte}, RRE: {rre * 180 / np.pi}")
        self.list_rte.append(rte)
        self.list_rre.append(rre)
        self.T_est.append(T_est)

    def report(self):
        print(f"rre thresh: {self.rre_thresh}; rte_thresh: {self.rte_thresh}")
        print(f"maximum frame dist: {self.max_dist:.2f}")
        msg = (f"RTE: {self.rte.avg}, var: {getattr(self.rte, 'var', 0.0)}, RRE: {self.rre.avg}, var: {getattr(self.rre, 'var', 0.0)}, "
               f"Success: {self.success.sum} / {self.success.count} ({self.success.avg * 100} %)")
        logging.info(msg)
        print(msg)
        return dict(rte=self.rte.avg, rre=self.rre.avg, success=self.success.sum, count=self.success.count,
                    recall=self.success.avg, T_est=self.T_est, list_rte=self.list_rte, list_rre=self.list_rre)


def main(config, test_loader):
    """scripts/test_kitti.py:75-233, one pair per iteration, on the drop-in symbols."""
    from ..sparse import SparseTensor
    if config.get('use_RANSAC'):
        raise NotImplementedError('Open3D RANSAC (scripts/test_kitti.py:171-177) is outside the hot path: pass --use_RANSAC false')
    device = torch.device('cuda')
    model = build_model(config, device)
    matcher = _matcher(config)
    stats = _Stats(config)
    data_timer, feat_timer, reg_timer = Timer(), Timer(sync=True), Timer(sync=True)
    print(f"rre thresh: {config.rre_thresh}; rte_thresh: {config.rte_thresh}")
    N = len(test_loader)
    it = iter(test_loader)
    dists_nn = []
    for i in range(N):
        data_timer.tic()
        data_dict = next(it)
        data_timer.toc()
        xyz0, xyz1 = data_dict['pcd0'][0], data_dict['pcd1'][0]
        T_gth = data_dict['T_gt'][0]
        xyz0np, xyz1np = xyz0.numpy(), xyz1.numpy()
        with torch.no_grad():
            feat_timer.tic()
            enc0 = model(SparseTensor(data_dict['sinput0_F'].to(device), coordinates=data_dict['sinput0_C'].to(device)))
            F0 = enc0.F.detach()
            enc1 = model(SparseTensor(data_dict['sinput1_F'].to(device), coordinates=data_dict['sinput1_C'].to(device)))
            F1 = enc1.F.detach()
            feat_timer.toc()
        if 'desc0' in data_dict:                   # synthetic data: descriptors with signal (the forward pass above still ran)
            F0, F1 = data_dict['desc0'].to(device), data_dict['desc1'].to(device)
        xyz0_corr, xyz1_corr = find_corr(xyz0, xyz1, F0, F1, subsample_size=5000)
        dists_nn.append(evaluate_nn_dist(xyz0_corr, xyz1_corr, T_gth))
        xyz0np, F0 = random_sample(xyz0np, F0, 5000)
        xyz1np, F1 = random_sample(xyz1np, F1, 5000)
        reg_timer.tic()
        x0, x1 = torch.from_numpy(xyz0np).to(device), torch.from_numpy(xyz1np).to(device)
        T_ransac, _, _, _, _ = matcher.estimator(x0[None, :], x1[None, :], F0[None, :], F1[None, :])
        T_ransac = T_ransac[0].to('cpu')
        reg_timer.toc()
        stats.add(T_ransac, T_gth)
        if i % 10 == 0:
            logging.info(f"{i} / {N}: Data time: {data_timer.avg}, Feat time: {feat_timer.avg}, Reg time: {reg_timer.avg}, "
                         f"RTE: {stats.rte.avg}, RRE: {stats.rre.avg}, Success: {stats.success.sum} / {stats.success.count} "
                         f"({stats.success.avg * 100} %)")
    out = stats.report()
    out['dists_nn'] = dists_nn
    return out


def main_blocks(config, test_loader, block=32):
    """The same evaluation through RegistrationPipeline: ``block`` pairs per launch sequence.  With the global numpy RNG
    seeded identically it produces the per-pair loop's poses bit for bit (tests/test_pipeline_gpu.py)."""
    from ..pipeline import RegistrationPipeline
    if config.get('use_RANSAC'):
        raise NotImplementedError('Open3D RANSAC is outside the hot path: pass --use_RANSAC false')
    device = torch.device('cuda')
    model = build_model(config, device)
    pipe = RegistrationPipeline(model, _matcher(config))
    stats = _Stats(config)
    print(f"rre thresh: {config.rre_thresh}; rte_thresh: {config.rte_thresh}")
    timer = Timer(sync=True)
    batch = []

    def flush():
        if not batch:
            return
        coords, xyz, descs, sizes = [], [], [], []
        for j, d in enumerate(batch):
            for side in (0, 1):
                c = d[f'sinput{side}_C'].to(device).clone()
                c[:, 0] = 2 * j + side                       # batch column = cloud id inside the block
                coords.append(c)
                xyz.append(d[f'pcd{side}'][0].to(device))
                if 'desc0' in d:
                    descs.append(d[f'desc{side}'].to(device))
            sizes.append((d['sinput0_C'].shape[0], d['sinput1_C'].shape[0]))
        timer.tic()
        out = pipe.run(torch.cat(coords).to(torch.int32), torch.cat(xyz).float(), sizes,
                       descriptors=torch.cat(descs) if descs else None)
        Ts = out['trans'].cpu()
        timer.toc()
        for j, d in enumerate(batch):
            stats.add(Ts[j], d['T_gt'][0])
        logging.info(f"block of {len(batch)} pairs: {timer.diff:.3f} s ({len(batch) / max(timer.diff, 1e-9):.1f} pairs/s), "
                     f"Success: {stats.success.sum} / {stats.success.count}")
        batch.clear()

    for d in test_loader:
        batch.append(d)
        if len(batch) == block:
            flush()
    flush()
    return stats.report()


def str2bool(v):
    return v.lower() in ('true', '1')


def parse_args(argv=None):
    """The reference's flags (scripts/test_kitti.py:240-255) plus the data-source / batching switches of this driver."""
    parser = argparse.ArgumentParser()
    parser.add_argument('--save_dir', default=None, type=str)
    parser.add_argument('--test_phase', default='test', type=str)
    parser.add_argument('--dataset', default=None, type=str)
    parser.add_argument('--LoKITTI', default=False, type=str2bool)
    parser.add_argument('--LoNUSCENES', default=False, type=str2bool)
    parser.add_argument('--LoWAYMO', default=False, type=str2bool)
    parser.add_argument('--test_num_thread', default=5, type=int)
    parser.add_argument('--pair_min_dist', default=None, type=int)
    parser.add_argument('--pair_max_dist', default=None, type=int)
    parser.add_argument('--downsample_single', default=1.0, type=float)
    parser.add_argument('--kitti_root', type=str, default="/data/kitti/")
    parser.add_argument('--use_RANSAC', type=str2bool, default=True)
    parser.add_argument('--rre_thresh', default=5.0, type=float)
    parser.add_argument('--rte_thresh', default=2.0, type=float)
    parser.add_argument('--synthetic', default=0, type=int, help='evaluate N synthetic KITTI-shaped pairs instead of a dataset')
    parser.add_argument('--block', default=0, type=int, help='pairs per batched launch sequence (0 = the reference per-pair loop)')
    parser.add_argument('--seed', default=0, type=int, help='np.random seed (the reference leaves the global RNG unseeded)')
    return parser.parse_args(argv)


def make_config(args):
    """scripts/test_kitti.py:257-292: <save_dir>/config.json, the CLI overrides, and config_KITTI.json when SC2-PCR is on."""
    config = _Cfg(CONFIG_DEFAULTS)
    if args.save_dir and os.path.exists(os.path.join(args.save_dir, 'config.json')):
        config.update(json.load(open(os.path.join(args.save_dir, 'config.json'))))
    config.save_dir = args.save_dir
    config.test_phase = args.test_phase
    config.kitti_root = args.kitti_root
    config.kitti_odometry_root = args.kitti_root + '/dataset'
    config.test_num_thread = args.test_num_thread
    config.LoKITTI, config.LoNUSCENES, config.LoWAYMO = args.LoKITTI, args.LoNUSCENES, args.LoWAYMO
    config.phase = 'test'
    config.use_RANSAC = args.use_RANSAC
    config.dataset = args.dataset
    config.supervised = False
    if not config.use_RANSAC:
        config.update(CONFIG_KITTI)
    if args.pair_min_dist is not None and args.pair_max_dist is not None:
        config.pair_min_dist, config.pair_max_dist = args.pair_min_dist, args.pair_max_dist
    config.downsample_single = args.downsample_single
    config.rte_thresh, config.rre_thresh = args.rte_thresh, args.rre_thresh
    return config


if __name__ == '__main__':
    logging.basicConfig(level=logging.INFO, format='%(asctime)s %(message)s')
    _args = parse_args()
    _config = make_config(_args)
    if _args.synthetic <= 0:
        raise SystemExit('no dataset loader is bundled (datasets are outside the hot path): pass --synthetic N, or call '
                         'main(config, make_data_loader(config, ...)) with the reference loader')
    np.random.seed(_args.seed)
    _loader = SyntheticPairLoader(range(_args.synthetic))
    (main_blocks(_config, _loader, _args.block) if _args.block > 0 else main(_config, _loader))
