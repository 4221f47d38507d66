"""CPU oracle: SC2-PCR estimator, weighted Kabsch, SE(3) helpers.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Dense, batch-size-1 restatement in
torch CPU fp32 of the reference estimator, written so that every tensor op is the
same torch op the reference issues -> on CPU the two are bit-identical, which is
what ``oracle/pin_against_reference.py`` asserts.  Citations are relative to
/root/reference.

The functions return *all* intermediate discrete quantities (seed list, top-k
index sets, iteration counts, inlier masks) so the CUDA path can be checked stage
by stage, each stage being fed the oracle's upstream tensors.
"""
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class SC2Config:
    """scripts/SC2_PCR/config_json/config_KITTI.json:2-14 (KITTI defaults)."""
    inlier_threshold: float = 0.6
    num_node: object = 8000
    d_thre: float = 0.1
    num_iterations: int = 20
    ratio: float = 0.2
    nms_radius: float = 0.6
    max_points: int = 8000
    k1: int = 30
    k2: int = 20
    # Tie rule of the three argsort(descending=True) calls (SC2_PCR.py:53,84,105).  False = exactly what
    # the reference issues (stable unspecified): torch-CPU then orders ties by its AVX-512 quicksort, which
    # is deterministic but not index-ordered (measured: 0.03 % agreement with a stable sort on 8000 integer
    # scores) - this is the mode pinned bit-for-bit against the reference.  True = 'descending value, lowest
    # index first', the rule the CUDA path implements (and what torch-CUDA's radix sort yields).
    stable_ties: bool = False


# --------------------------------------------------------------------------- SE3
def se3_transform(pts, trans):
    """scripts/SC2_PCR/utils/SE3.py:43-57 (batched branch): R @ p + t."""
    out = trans[:, :3, :3] @ pts.permute(0, 2, 1) + trans[:, :3, 3:4]
    return out.permute(0, 2, 1)


def se3_integrate(R, t):
    """scripts/SC2_PCR/utils/SE3.py:73-96 (batched torch branch)."""
    T = torch.eye(4)[None].repeat(R.shape[0], 1, 1).to(R.device)
    T[:, :3, :3] = R
    T[:, :3, 3:4] = t.view([-1, 3, 1])
    return T


# ------------------------------------------------------------------------ Kabsch
def kabsch_weighted(A, B, weights=None, weight_threshold=0):
    """scripts/SC2_PCR/common.py:7-45.  Mutates ``weights`` in place like :20."""
    bs = A.shape[0]
    if weights is None:
        weights = torch.ones_like(A[:, :, 0])
    weights[weights < weight_threshold] = 0
    wsum = torch.sum(weights, dim=1, keepdim=True)[:, :, None] + 1e-6
    cA = torch.sum(A * weights[:, :, None], dim=1, keepdim=True) / wsum
    cB = torch.sum(B * weights[:, :, None], dim=1, keepdim=True) / wsum
    Am = A - cA
    Bm = B - cB
    H = Am.permute(0, 2, 1) @ torch.diag_embed(weights) @ Bm        # :33-34
    U, S, V = torch.svd(H.cpu())                                     # :36 (LAPACK on the CPU whatever the device)
    U, S, V = U.to(weights.device), S.to(weights.device), V.to(weights.device)   # :37
    delta = torch.det(V @ U.permute(0, 2, 1))
    eye = torch.eye(3)[None].repeat(bs, 1, 1).to(A.device)
    eye[:, -1, -1] = delta
    R = V @ eye @ U.permute(0, 2, 1)
    t = cB.permute(0, 2, 1) - R @ cA.permute(0, 2, 1)
    return se3_integrate(R, t)


def kabsch_weighted_nodiag(A, B, weights):
    """Same maths as kabsch_weighted without the dense diag_embed (common.py:33 builds
    a [bs,n,n] matrix -> 256 MB at n=8000).  Used only to keep oracle memory sane in
    the refinement loop; (Am^T * w) @ Bm sums the same products in a different order,
    so it is tolerance-level (not bit-level) equivalent.  NOT used for pinning."""
    wsum = torch.sum(weights, dim=1, keepdim=True)[:, :, None] + 1e-6
    cA = torch.sum(A * weights[:, :, None], dim=1, keepdim=True) / wsum
    cB = torch.sum(B * weights[:, :, None], dim=1, keepdim=True) / wsum
    H = (A - cA).permute(0, 2, 1) * weights[:, None, :] @ (B - cB)
    U, S, V = torch.svd(H.cpu())
    U, V = U.to(A.device), V.to(A.device)
    eye = torch.eye(3)[None].repeat(A.shape[0], 1, 1).to(A.device)
    eye[:, -1, -1] = torch.det(V @ U.permute(0, 2, 1))
    R = V @ eye @ U.permute(0, 2, 1)
    return se3_integrate(R, cB.permute(0, 2, 1) - R @ cA.permute(0, 2, 1))


# ------------------------------------------------------------- power iteration
def power_iteration(M, num_iterations):
    """scripts/SC2_PCR/SC2_PCR.py:179-190.  Returns (vector [bs,n], iterations run)."""
    v = torch.ones_like(M[:, :, 0:1])
    v_last = v
    it = 0
    for it in range(1, num_iterations + 1):
        v = torch.bmm(M, v)
        v = v / (torch.norm(v, dim=1, keepdim=True) + 1e-6)
        if torch.allclose(v, v_last):
            break
        v_last = v
    return v.squeeze(-1), it


# ------------------------------------------------------------------ first order
def pairwise_dist(p):
    """SC2_PCR.py:333-334: torch.norm over the broadcast difference."""
    return torch.norm(p[:, :, None, :] - p[:, None, :, :], dim=-1)


def first_order(src, tgt, cfg):
    """SC2_PCR.py:333-342,357-358 -> src_dist, cross_dist, SC, hard, tight."""
    src_dist = pairwise_dist(src)
    cross = torch.abs(src_dist - pairwise_dist(tgt))
    SC = torch.clamp(1.0 - cross ** 2 / cfg.d_thre ** 2, min=0)
    hard = (cross < cfg.d_thre).float()
    tight = (cross < cfg.d_thre / 2).float()
    return src_dist, cross, SC, hard, tight


def pick_seeds(dists, scores, R, max_num, stable=False):
    """SC2_PCR.py:33-59 (bs = 1)."""
    assert scores.shape[0] == 1
    rel = scores.T >= scores
    rel = rel.bool() | (dists[0] >= R).bool()
    is_local_max = rel.min(-1)[0].float()
    order = torch.argsort(scores * is_local_max, dim=1, descending=True, stable=stable)
    return order[:, 0:max_num].detach()


# ------------------------------------------------------------------ seed stage
def seed_stage(seeds, SC2, src, tgt, cfg, detail=None):
    """SC2_PCR.py:61-168.  Returns (final_trans [bs,4,4], seedwise_fitness [bs,S])."""
    bs, _, num_channels = SC2.shape
    k1, k2 = cfg.k1, cfg.k2
    if k1 > num_channels:                                             # :76-78
        k1 = k2 = 4
    knn_idx = torch.argsort(SC2, dim=2, descending=True, stable=cfg.stable_ties)[:, :, 0:k1]  # :84-85
    flat = knn_idx.contiguous().view([bs, -1])[:, :, None].expand(-1, -1, 3)
    src_knn = src.gather(dim=1, index=flat).view([bs, -1, k1, 3])
    tgt_knn = tgt.gather(dim=1, index=flat).view([bs, -1, k1, 3])
    sd = ((src_knn[:, :, :, None, :] - src_knn[:, :, None, :, :]) ** 2).sum(-1) ** 0.5
    td = ((tgt_knn[:, :, :, None, :] - tgt_knn[:, :, None, :, :]) ** 2).sum(-1) ** 0.5
    local_hard = (torch.abs(sd - td) < cfg.d_thre).float()            # :99
    local_sc2 = torch.matmul(local_hard[:, :, :1, :], local_hard)     # :100
    fine = torch.argsort(local_sc2, dim=3, descending=True, stable=cfg.stable_ties)[:, :, :, 0:k2]   # :105-106
    num = fine.shape[1]
    fine_e = fine.contiguous().view([bs, num, -1])[:, :, :, None].expand(-1, -1, -1, 3)
    src_f = src_knn.gather(dim=2, index=fine_e).view([bs, -1, k2, 3])
    tgt_f = tgt_knn.gather(dim=2, index=fine_e).view([bs, -1, k2, 3])
    sd = ((src_f[:, :, :, None, :] - src_f[:, :, None, :, :]) ** 2).sum(-1) ** 0.5
    td = ((tgt_f[:, :, :, None, :] - tgt_f[:, :, None, :, :]) ** 2).sum(-1) ** 0.5
    cross = torch.abs(sd - td)
    local = torch.clamp(1 - cross ** 2 / cfg.d_thre ** 2, min=0).view([-1, k2, k2])   # :122-125
    ar = torch.arange(k2)
    local[:, ar, ar] = 0                                              # :131
    w, iters = power_iteration(local, cfg.num_iterations)             # :132 (global allclose)
    w = w.view([bs, -1, k2])
    w = w / (torch.sum(w, dim=-1, keepdim=True) + 1e-6)               # :134
    w = w.view([-1, k2])
    seed_T = kabsch_weighted(src_f.view([-1, k2, 3]), tgt_f.view([-1, k2, 3]), w)
    seed_T = seed_T.view([bs, -1, 4, 4])
    pred = torch.einsum('bsnm,bmk->bsnk', seed_T[:, :, :3, :3], src.permute(0, 2, 1)) \
        + seed_T[:, :, :3, 3:4]                                       # :153-155
    pred = pred.permute(0, 1, 3, 2)
    L2 = torch.norm(pred - tgt[:, None, :, :], dim=-1)
    fitness = torch.sum((L2 < cfg.inlier_threshold).float(), dim=-1)  # :161
    best = fitness.argmax(dim=1)
    final = seed_T.gather(dim=1, index=best[:, None, None, None].expand(-1, -1, 4, 4)).squeeze(1)
    if detail is not None:
        fine_global = knn_idx.gather(2, fine.view(bs, num, k2))       # indices into the N corr.
        detail.update(topk1=knn_idx, fine_local=fine.view(bs, num, k2), topk2=fine_global,
                      local_iters=iters, seed_weights=w.view(bs, -1, k2), seed_trans=seed_T,
                      best_seed=best)
    return final, fitness


# ------------------------------------------------------------------ refinement
def post_refinement(T, src, tgt, it_num, cfg, detail=None, dense_weight=True):
    """SC2_PCR.py:238-278 (bs = 1)."""
    assert T.shape[0] == 1
    thr_list = [0.10] * it_num if cfg.inlier_threshold == 0.10 else [1.2] * it_num   # :254-257
    prev = 0
    counts = []
    for thr in thr_list:
        warped = se3_transform(src, T)
        L2 = torch.norm(warped - tgt, dim=-1)
        inl = (L2 < thr)[0]
        n = torch.sum(inl)
        if abs(int(n - prev)) < 1:
            break
        prev = n
        counts.append(int(n))
        w = 1 / (1 + (L2 / thr) ** 2)[:, inl]
        fn = kabsch_weighted if dense_weight else kabsch_weighted_nodiag
        T = fn(src[:, inl, :], tgt[:, inl, :], w)
    if detail is not None:
        detail.update(refine_counts=counts)
    return T


# -------------------------------------------------------------------- SC2_PCR
def sc2_pcr(src, tgt, cfg, detail=None, dense_weight=True):
    """SC2_PCR.py:307-384 -> (final_trans [1,4,4], seedwise_fitness [1,S])."""
    num_corr = tgt.shape[1]
    if num_corr > cfg.max_points:                                     # :324-327
        src, tgt, num_corr = src[:, :cfg.max_points], tgt[:, :cfg.max_points], cfg.max_points
    src_dist, cross, SC, hard, tight = first_order(src, tgt, cfg)
    conf, it0 = power_iteration(SC, cfg.num_iterations)               # :349
    seeds = pick_seeds(src_dist, conf, cfg.nms_radius, int(num_corr * cfg.ratio), cfg.stable_ties)   # :350
    s_hard = hard.gather(1, seeds[:, :, None].expand(-1, -1, num_corr))
    s_tight = tight.gather(1, seeds[:, :, None].expand(-1, -1, num_corr))
    SC2 = torch.matmul(s_tight, tight) * s_hard                       # :363
    if detail is not None:
        detail.update(confidence=conf, global_iters=it0, seeds=seeds, SC2=SC2,
                      hard=hard, tight=tight)
    T, fitness = seed_stage(seeds, SC2, src, tgt, cfg, detail)
    if detail is not None:
        detail.update(initial_trans=T.clone())
    T = post_refinement(T, src, tgt, 20, cfg, detail, dense_weight)   # :374
    return T, fitness


def match_pair(src_kp, tgt_kp, src_f, tgt_f, cfg, detail=None):
    """SC2_PCR.py:280-305 with the hard-coded ``.cuda()`` (:299) dropped.  Draws from the
    global numpy RNG in the reference's order: choice(N_src, num_node) then choice(N_tgt, ...)."""
    n_src, n_tgt = src_f.shape[1], tgt_f.shape[1]
    if cfg.num_node == 'all':
        si, ti = np.arange(n_src), np.arange(n_tgt)
    else:
        si = np.random.choice(n_src, cfg.num_node)
        ti = np.random.choice(n_tgt, cfg.num_node)
    sd, td = src_f[:, si, :], tgt_f[:, ti, :]
    skp, tkp = src_kp[:, si, :], tgt_kp[:, ti, :]
    distance = torch.sqrt(2 - 2 * (sd[0] @ td[0].T) + 1e-6)
    idx = torch.argmin(distance, dim=1)
    if detail is not None:
        detail.update(src_sel=si, tgt_sel=ti, nn_idx=idx)
    return skp, tkp[:, idx]


def estimator(src_kp, tgt_kp, src_f, tgt_f, cfg, detail=None, dense_weight=True):
    """SC2_PCR.py:386-413 -> (trans, labels, src_corr, tgt_corr, seedwise_fitness)."""
    sc, tc = match_pair(src_kp, tgt_kp, src_f, tgt_f, cfg, detail)
    T, fitness = sc2_pcr(sc, tc, cfg, detail, dense_weight)
    warp = se3_transform(sc, T)
    dist = torch.sum((warp - tc) ** 2, dim=-1) ** 0.5
    labels = (dist < cfg.inlier_threshold).float()
    return T, labels, sc, tc, fitness


# ------------------------------------------------------- metrics (test_kitti.py)
def rte_rre(T_est, T_gt):
    """scripts/test_kitti.py:188-191 (torch CPU tensors in, python floats out)."""
    rte = np.linalg.norm(T_est[:3, 3] - T_gt[:3, 3])
    tm = T_est[:3, :3].t() @ T_gt[:3, :3]
    tm[[0, 1, 2], [0, 1, 2]] = torch.min(torch.ones(3), tm[[0, 1, 2], [0, 1, 2]])
    rre = np.arccos((np.trace(tm) - 1) / 2)
    return float(rte), float(rre)
