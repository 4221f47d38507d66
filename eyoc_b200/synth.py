"""Synthetic KITTI-shaped inputs (numpy only; no datasets or checkpoints are reachable here).

Follows the generator specified in SURVEY.md §8d: a 64-beam spinning LiDAR ray-cast against a
ground plane plus axis-aligned boxes, voxelised exactly like the reference's loader
(``floor(xyz / voxel)``, first-occurrence unique; /root/reference lib/data_loaders.py:936-972),
and collated with a leading batch column (lib/data_loaders.py:31-85).
"""
import numpy as np

VOXEL_SIZE = 0.3          # scripts/train_kitti_EYOC.sh:20
SENSOR_HEIGHT = 1.73
MAX_RANGE = 80.0


def _yaw(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def make_world(rng, n_box_mean=40):
    n = max(4, rng.poisson(n_box_mean))
    centre = rng.uniform(-MAX_RANGE, MAX_RANGE, (n, 2))
    half = rng.uniform(0.5, 7.5, (n, 2))
    height = rng.uniform(1.5, 8.0, n)
    lo = np.concatenate([centre - half, np.zeros((n, 1))], 1)
    hi = np.concatenate([centre + half, height[:, None]], 1)
    keep = np.linalg.norm(centre, axis=1) > 12.0          # keep the road around the sensor free
    return lo[keep], hi[keep]


def _ray_dirs(az_step_deg=0.17):
    # HDL-64E layout: upper block 32 beams at 1/3 deg, lower block 32 beams at 1/2 deg
    elev = np.deg2rad(np.concatenate([np.linspace(2.0, -8.33, 32), np.linspace(-8.83, -24.33, 32)]))
    az = np.deg2rad(np.arange(0.0, 360.0, az_step_deg))
    e, a = np.meshgrid(elev, az, indexing='ij')
    return np.stack([np.cos(e) * np.cos(a), np.cos(e) * np.sin(a), np.sin(e)], -1).reshape(-1, 3)


def scan(world, position, yaw, rng, az_step_deg=0.17, noise=0.02, ground_roughness=0.07):
    """Ray-cast one sweep; returns points in the SENSOR frame, float32 [M,3].  Ground returns get
    an extra vertical roughness term (grass / road texture) so that voxel counts land at ~30 k."""
    lo, hi = world
    d_local = _ray_dirs(az_step_deg=az_step_deg)
    d = d_local @ _yaw(yaw).T
    o = np.asarray(position, np.float64)
    t_hit = np.full(len(d), np.inf)
    down = d[:, 2] < -1e-9
    t_hit[down] = -o[2] / d[down, 2]
    with np.errstate(divide='ignore', invalid='ignore'):
        inv = 1.0 / d
        for b in range(len(lo)):
            t0 = (lo[b] - o) * inv
            t1 = (hi[b] - o) * inv
            tn = np.minimum(t0, t1).max(1)
            tf = np.maximum(t0, t1).min(1)
            ok = (tn <= tf) & (tn > 0.5)
            t_hit = np.where(ok & (tn < t_hit), tn, t_hit)
    r = t_hit + rng.normal(0.0, noise, len(d))
    ok = np.isfinite(t_hit) & (r < MAX_RANGE) & (r > 1.0)
    pts = d_local[ok] * r[ok, None]
    ground = pts[:, 2] < -(o[2] - 0.13)
    pts[ground, 2] += rng.normal(0.0, ground_roughness, int(ground.sum()))
    return pts.astype(np.float32)


def voxelize(xyz, voxel_size=VOXEL_SIZE):
    """lib/data_loaders.py:940-972: sparse_quantize(xyz/voxel) first-occurrence selection, then
    coords = floor(xyz[sel]/voxel).int().  Returns (xyz[sel] float32 [N,3], coords int32 [N,3])."""
    q = np.floor(xyz / np.float32(voxel_size)).astype(np.int64)
    key = ((q[:, 0] + 32768) << 32) | ((q[:, 1] + 32768) << 16) | (q[:, 2] + 32768)
    _, first = np.unique(key, return_index=True)
    sel = np.sort(first)
    return xyz[sel], q[sel].astype(np.int32)


def make_pair(pair_id, distance=None, az_step_deg=0.17):
    """One synthetic pair.  Returns dict(xyz0, xyz1 [N,3] f32 metres, coords0, coords1 [N,3] i32,
    T_gt [4,4] f32 mapping cloud-0 sensor coordinates to cloud-1 sensor coordinates)."""
    rng = np.random.default_rng(1234 + pair_id)
    world = make_world(rng)
    yaw0 = rng.uniform(-np.pi, np.pi)
    dyaw = np.deg2rad(rng.uniform(-10.0, 10.0))
    dist = rng.uniform(5.0, 50.0) if distance is None else distance
    c0 = np.array([0.0, 0.0, SENSOR_HEIGHT])
    c1 = c0 + dist * np.array([np.cos(yaw0), np.sin(yaw0), 0.0])
    p0 = scan(world, c0, yaw0, rng, az_step_deg)
    p1 = scan(world, c1, yaw0 + dyaw, rng, az_step_deg)
    R0, R1 = _yaw(yaw0), _yaw(yaw0 + dyaw)
    T = np.eye(4)
    T[:3, :3] = R1.T @ R0
    T[:3, 3] = R1.T @ (c0 - c1)
    xyz0, coords0 = voxelize(p0)
    xyz1, coords1 = voxelize(p1)
    return dict(xyz0=xyz0, xyz1=xyz1, coords0=coords0, coords1=coords1, T_gt=T.astype(np.float32))


def collate(coords_list):
    """ME.utils.sparse_collate for coordinates: prepend the batch index -> int32 [sum N, 4]."""
    out = [np.concatenate([np.full((len(c), 1), b, np.int32), c.astype(np.int32)], 1)
           for b, c in enumerate(coords_list)]
    return np.ascontiguousarray(np.concatenate(out, 0))


def planted_descriptors(xyz0, xyz1, T_gt, rng, dim=32, sigma=0.12, radius=VOXEL_SIZE):
    """SURVEY.md §8d 'planted descriptors': voxels of cloud 0 whose GT-transformed position has a
    cloud-1 voxel within ``radius`` share a random unit vector (+ N(0, sigma^2) noise, renormalised)
    with that voxel; everything else gets an independent random unit vector."""
    from scipy.spatial import cKDTree
    f0 = rng.normal(size=(len(xyz0), dim))
    f1 = rng.normal(size=(len(xyz1), dim))
    moved = xyz0.astype(np.float64) @ T_gt[:3, :3].T.astype(np.float64) + T_gt[:3, 3]
    dist, nn = cKDTree(xyz1).query(moved, k=1)
    hit = dist < radius
    f0 /= np.linalg.norm(f0, axis=1, keepdims=True)
    f1[nn[hit]] = f0[hit] + rng.normal(0.0, sigma, (int(hit.sum()), dim))
    f0[hit] += rng.normal(0.0, sigma, (int(hit.sum()), dim))
    f0 /= np.linalg.norm(f0, axis=1, keepdims=True)
    f1 /= np.linalg.norm(f1, axis=1, keepdims=True)
    return f0.astype(np.float32), f1.astype(np.float32), hit


def make_correspondences(n, inlier_ratio, seed, noise=0.03, extent=60.0, dup_ratio=0.2):
    """Synthetic putative correspondences with a planted pose, including the duplicate
    correspondences that with-replacement sampling creates.
    Returns src, tgt float32 [n,3], T [4,4] float32, inlier mask bool [n]."""
    rng = np.random.default_rng(seed)
    n_unique = max(4, int(n * (1.0 - dup_ratio)))
    src = rng.uniform(-extent, extent, (n_unique, 3)) * np.array([1.0, 1.0, 0.08])
    ang = rng.uniform(-0.3, 0.3)
    R = _yaw(ang) @ np.array([[1, 0, 0], [0, np.cos(0.02), -np.sin(0.02)], [0, np.sin(0.02), np.cos(0.02)]])
    t = rng.uniform(-20.0, 20.0, 3) * np.array([1.0, 1.0, 0.05])
    inl = rng.random(n_unique) < inlier_ratio
    tgt = src @ R.T + t + rng.normal(0.0, noise, src.shape)
    tgt[~inl] = rng.uniform(-extent, extent, (int((~inl).sum()), 3)) * np.array([1.0, 1.0, 0.08])
    pick = np.concatenate([np.arange(n_unique), rng.integers(0, n_unique, n - n_unique)])
    rng.shuffle(pick)
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    return (src[pick].astype(np.float32), tgt[pick].astype(np.float32), T.astype(np.float32), inl[pick])


# ---------------------------------------------------------------------------------------------- bulk generation
CACHE_DIR = '/tmp/eyoc_b200_synth'


def _make_one(args):
    pair_id, sigma = args
    import os
    path = os.path.join(CACHE_DIR, f'pair_{pair_id}_{sigma:.3f}.npz')
    if os.path.exists(path):
        try:
            with np.load(path) as z:
                return {k: z[k] for k in z.files}
        except Exception:
            pass
    p = make_pair(pair_id)
    f0, f1, hit = planted_descriptors(p['xyz0'], p['xyz1'], p['T_gt'], np.random.default_rng(99991 + pair_id), sigma=sigma)
    p.update(desc0=f0, desc1=f1, overlap=np.float32(hit.mean()))
    os.makedirs(CACHE_DIR, exist_ok=True)
    tmp = path + f'.{os.getpid()}.tmp.npz'
    np.savez(tmp, **p)
    os.replace(tmp, path)
    return p


def make_pairs(pair_ids, sigma=0.12, workers=None):
    """Generate (or load from the /tmp cache) synthetic pairs with planted descriptors, in parallel."""
    import multiprocessing as mp
    import os
    workers = workers or min(len(pair_ids), os.cpu_count() or 1)
    args = [(int(i), float(sigma)) for i in pair_ids]
    if workers <= 1 or len(args) <= 1:
        return [_make_one(a) for a in args]
    with mp.get_context('fork').Pool(workers) as pool:
        return pool.map(_make_one, args)


def collate_pairs(pairs):
    """[c0 of pair 0, c1 of pair 0, c0 of pair 1, ...] -> coords int32 [sum N, 4], xyz f32 [sum N, 3],
    descriptors f32 [sum N, 32] (or None), sizes [(n0, n1), ...]."""
    clouds = [c for p in pairs for c in (p['coords0'], p['coords1'])]
    coords = collate(clouds)
    xyz = np.ascontiguousarray(np.concatenate([x for p in pairs for x in (p['xyz0'], p['xyz1'])], 0), np.float32)
    desc = None
    if 'desc0' in pairs[0]:
        desc = np.ascontiguousarray(np.concatenate([x for p in pairs for x in (p['desc0'], p['desc1'])], 0), np.float32)
    sizes = [(len(p['coords0']), len(p['coords1'])) for p in pairs]
    return coords, xyz, desc, sizes
