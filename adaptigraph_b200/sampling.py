"""Particle sampling in front of the dynamics path, with the reference's signatures.

    fps(obj_kp_start, max_nobj, fps_radius_range, verbose=False)     src/dynamics/dataset/graph.py:8-36
    fps_rad_idx(pcd, radius)                                         src/dynamics/utils.py:10-24
    farthest_point_sampler(pos, npoints, start_idx)                  dgl.geometry (graph.py:11, perception.py:271)

The selections run on the GPU (`agx_fps`, csrc/sampling.cu); the wrappers draw the SAME numpy random numbers in the same
order as the reference (start index, radius, second start index), so a seeded reference run and a seeded run through these
functions pick the same particles.  There is no CPU path: numpy inputs are copied to the current CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _device(device=None) -> torch.device:
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def farthest_point_sampler(pos, npoints: int, start_idx=-1) -> torch.Tensor:
    """dgl.geometry.farthest_point_sampler: pos (B,N,3) -> (B,npoints) int64 indices; start_idx -1 draws one at random
    (numpy), an int is used for every cloud, a (B,) tensor per cloud."""
    pos = torch.as_tensor(pos)
    dev = pos.device if pos.is_cuda else _device()
    pos_d = pos.to(dev, torch.float32)
    B, N, _ = pos_d.shape
    if torch.is_tensor(start_idx):
        start = start_idx.to(dev, torch.int32).reshape(B)
    else:
        s = int(np.random.randint(0, N)) if start_idx == -1 else int(start_idx)
        if not 0 <= s < N:
            raise ValueError(f"start_idx {s} out of range for {N} points")
        start = torch.full((B,), s, dtype=torch.int32, device=dev)
    n_pts = torch.full((B,), N, dtype=torch.int32, device=dev)
    idx, _ = ops.fps(pos_d, n_pts, start, int(npoints), -1.0)
    return idx.to(torch.int64).to(pos.device if pos.is_cuda else "cpu")


def fps_rad_idx(pcd, radius):
    """utils.py:10-24: (n,3) numpy -> (kept points (m,3), kept indices (m,)); first pick = np.random.randint(n)."""
    pcd = np.asarray(pcd)
    n = pcd.shape[0]
    rand_idx = np.random.randint(n)
    dev = _device()
    pos_d = torch.from_numpy(np.ascontiguousarray(pcd, dtype=np.float32)).to(dev)[None]
    idx, cnt = ops.fps(pos_d, torch.tensor([n], dtype=torch.int32, device=dev),
                       torch.tensor([rand_idx], dtype=torch.int32, device=dev), n, float(radius))
    idx_lst = idx[0, :int(cnt.item())].cpu().numpy().astype(np.int64)
    return pcd[idx_lst], idx_lst


def fps(obj_kp_start, max_nobj, fps_radius_range, verbose=False):
    """graph.py:8-36: farthest-point sample to at most max_nobj particles, then thin to a uniform radius; returns the
    indices into obj_kp_start (int32 numpy)."""
    particle = torch.from_numpy(np.asarray(obj_kp_start)).float().unsqueeze(0)   # [1, N, 3]
    fps_idx_1 = farthest_point_sampler(particle, min(max_nobj, particle.shape[1]),
                                       start_idx=np.random.randint(0, particle.shape[1]))[0].numpy().astype(np.int32)
    downsample_particle = particle[0, fps_idx_1].numpy()
    if type(fps_radius_range) == float:
        fps_radius = fps_radius_range
    elif len(fps_radius_range) == 2:
        fps_radius = np.random.uniform(fps_radius_range[0], fps_radius_range[1])
    else:
        raise ValueError(f"Invalid fps_radius_range: {fps_radius_range}.")
    _, fps_idx_2 = fps_rad_idx(downsample_particle, fps_radius)
    fps_idx = fps_idx_1[fps_idx_2.astype(np.int32)]
    if verbose:
        print(f"FPS num particles: {len(fps_idx)} with index list \n {fps_idx}. \n")
    return np.array(fps_idx)


def fps_batch(pos: torch.Tensor, n_points: torch.Tensor, max_nobj: int, radius, start_idx: torch.Tensor,
              start_idx_2: torch.Tensor):
    """Device-resident batched form of `fps` for B clouds at once (no host round trip): pos (B,N,3) CUDA, n_points (B),
    start_idx (B) first pick among the cloud's points, start_idx_2 (B) first pick among the max_nobj survivors; `radius` is a
    float or a (B) tensor (one thinning radius per cloud, as a training batch draws them: graph.py:17-20).
    Returns idx (B, max_nobj) int32 indices into the clouds and counts (B)."""
    B, N, _ = pos.shape
    k = min(int(max_nobj), N)
    idx1, cnt1 = ops.fps(pos, n_points, start_idx, k, -1.0)
    cnt1 = torch.minimum(cnt1, n_points.to(torch.int32))          # a cloud with fewer than k points keeps them all once
    sub = torch.gather(pos, 1, idx1.to(torch.int64).unsqueeze(-1).expand(B, k, 3))
    if torch.is_tensor(radius):
        idx2, cnt2 = ops.fps_radii(sub, cnt1, start_idx_2, k, radius.to(pos.device))
    else:
        idx2, cnt2 = ops.fps(sub, cnt1, start_idx_2, k, float(radius))
    return torch.gather(idx1, 1, idx2.to(torch.int64)), cnt2
