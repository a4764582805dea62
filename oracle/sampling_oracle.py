"""TEST INFRASTRUCTURE — CPU restatement of the particle samplers in front of the dynamics path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

* `fps_rad_idx` follows the reference's src/dynamics/utils.py:10-24 line by line (the random first pick is passed in).
  PINNED: tests/golden/fps_rad_idx.npz holds outputs of the reference function itself (tests/golden/make_golden_fps.py).
* `farthest_point_sampler` restates dgl.geometry.farthest_point_sampler, the third-party sampler the reference calls at
  src/dynamics/dataset/graph.py:11-12 and src/planning/perception.py:271.  DGL is NOT in /root/reference and is not pinned by
  it (README.md:47 installs an unversioned `dgl`); the restatement follows DGL's published CPU operator
  (src/geometry/cpu/geometry_op_impl.cc, FarthestPointSampler, DGL 1.x/2.x): per cloud, running minimum of the SQUARED
  distance to the picks accumulated over x, y, z in the array's dtype, next pick = first index of the maximum.
  PARITY UNPINNED for this function: no DGL build is available offline to generate fixtures.
"""
import numpy as np


def farthest_point_sampler(pos, npoints, start_idx):
    """pos (B,N,3) float32, start_idx (B,) -> (B,npoints) int64."""
    pos = np.asarray(pos, dtype=np.float32)
    B, N, _ = pos.shape
    out = np.zeros((B, npoints), dtype=np.int64)
    for b in range(B):
        dist = np.full(N, np.inf, dtype=np.float32)
        cur = int(start_idx[b])
        out[b, 0] = cur
        for i in range(1, npoints):
            d = pos[b] - pos[b, cur]
            one = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]    # fp32, x then y then z
            dist = np.minimum(dist, one)
            cur = int(np.argmax(dist))                                             # first maximum (strict > scan)
            out[b, i] = cur
    return out


def fps_rad_idx(pcd, radius, rand_idx):
    """utils.py:10-24 with the random first pick made explicit."""
    pcd_fps_lst = [pcd[rand_idx]]
    idx_lst = [rand_idx]
    dist = np.linalg.norm(pcd - pcd_fps_lst[0], axis=1)
    while dist.max() > radius:
        pcd_fps_lst.append(pcd[dist.argmax()])
        idx_lst.append(dist.argmax())
        dist = np.minimum(dist, np.linalg.norm(pcd - pcd_fps_lst[-1], axis=1))
    return np.stack(pcd_fps_lst, axis=0), np.stack(idx_lst, axis=0)


def fps(obj_kp_start, max_nobj, fps_radius, start_1, start_2):
    """graph.py:8-36 with the random draws made explicit (float radius branch)."""
    particle = np.asarray(obj_kp_start, dtype=np.float32)
    idx1 = farthest_point_sampler(particle[None], min(max_nobj, particle.shape[0]), [start_1])[0].astype(np.int32)
    _, idx2 = fps_rad_idx(particle[idx1], fps_radius, start_2)
    return idx1[idx2.astype(np.int32)]
