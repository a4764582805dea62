"""Generate tests/golden/fps_cases.npz from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_fps.py

* `fps_rad_idx` cases are outputs of the reference's own function (src/dynamics/utils.py:10-24) under a seeded numpy RNG.
* `fps` cases run the reference's own wrapper (src/dynamics/dataset/graph.py:8-36) with the one dependency that is absent
  from the image, dgl.geometry.farthest_point_sampler, replaced by oracle/sampling_oracle.farthest_point_sampler (the
  restatement of DGL's published algorithm): they pin the wrapper's composition and its order of random draws, not DGL.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"


def main():
    from oracle import sampling_oracle as so
    sys.path.insert(0, REF)
    dgl = types.ModuleType("dgl")
    geo = types.ModuleType("dgl.geometry")

    def sampler(pos, npoints, start_idx=-1):
        B = pos.shape[0]
        return torch.from_numpy(so.farthest_point_sampler(pos.numpy(), npoints, [start_idx] * B))
    geo.farthest_point_sampler = sampler
    dgl.geometry = geo
    mp = types.ModuleType("moviepy")
    mpe = types.ModuleType("moviepy.editor")
    mp.editor = mpe
    sys.modules.update({"dgl": dgl, "dgl.geometry": geo, "moviepy": mp, "moviepy.editor": mpe})
    from dynamics.utils import fps_rad_idx
    from dynamics.dataset.graph import fps

    rng = np.random.default_rng(7)
    clouds = {
        "blob300": rng.normal(0, 1.0, (300, 3)).astype(np.float32),
        "cloth45": np.stack(np.meshgrid(np.arange(45) * 0.25, [0.0], np.arange(45) * 0.25, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
                   + rng.normal(0, 0.01, (2025, 3)).astype(np.float32),
        "rope500": np.stack([0.02 * np.arange(500), np.zeros(500), 0.3 * np.sin(0.02 * np.arange(500))], -1).astype(np.float32),
        "dup64": np.repeat(rng.normal(0, 1, (16, 3)).astype(np.float32), 4, axis=0),      # exact ties
        "single": np.zeros((1, 3), np.float32),
    }
    out = {}
    for name, pcd in clouds.items():
        out[f"{name}/pcd"] = pcd
        for k, radius in enumerate([0.05, 0.3, 0.8]):
            seed = 100 + k
            np.random.seed(seed)
            _, idx = fps_rad_idx(pcd, radius)
            out[f"{name}/rad{k}/radius"] = np.float64(radius)
            out[f"{name}/rad{k}/seed"] = np.int64(seed)
            out[f"{name}/rad{k}/idx"] = np.asarray(idx, np.int64).reshape(-1)
        for k, (max_nobj, rr) in enumerate([(100, 0.3), (40, [0.2, 0.5]), (5000, 0.1)]):
            seed = 200 + k
            np.random.seed(seed)
            idx = fps(pcd, max_nobj, rr)
            out[f"{name}/fps{k}/max_nobj"] = np.int64(max_nobj)
            out[f"{name}/fps{k}/range"] = np.asarray(rr, np.float64).reshape(-1)
            out[f"{name}/fps{k}/seed"] = np.int64(seed)
            out[f"{name}/fps{k}/idx"] = np.asarray(idx, np.int64).reshape(-1)
    np.savez_compressed(os.path.join(HERE, "fps_cases.npz"), **out)
    print("wrote fps_cases.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
