"""Generate tests/golden/dataset_rope.npz from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_dataset.py

Writes the synthetic on-disk data set of tests/agx_helpers.write_synthetic_dataset into a temporary directory, runs the
reference's own loaders and `DynDataset.__getitem__` (src/dynamics/dataset/load.py, dataset.py) on it under seeded numpy /
torch RNGs, and stores what they return: the frame-pair table, the physics parameters, every tensor of a few samples (the dense
`Rr` / `Rs` as per-row receiver / sender ids), the kept particle indices, and the index order a seeded shuffled
`DataLoader(num_workers=0)` visits.  Absent from the image and stubbed: `dgl.geometry.farthest_point_sampler` (replaced by
oracle/sampling_oracle's restatement, as in make_golden_fps.py), `moviepy`, `cv2`, `sim.utils.load_yaml` (unused by the calls made).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/src"


def stub_modules():
    from oracle import sampling_oracle as so
    dgl, geo = types.ModuleType("dgl"), types.ModuleType("dgl.geometry")

    def sampler(pos, npoints, start_idx=-1):
        return torch.from_numpy(so.farthest_point_sampler(pos.numpy(), npoints, [start_idx] * pos.shape[0]))
    geo.farthest_point_sampler = sampler
    dgl.geometry = geo
    mp, mpe = types.ModuleType("moviepy"), types.ModuleType("moviepy.editor")
    mp.editor = mpe
    sim, simu = types.ModuleType("sim"), types.ModuleType("sim.utils")
    simu.load_yaml = lambda path: None
    sim.utils = simu
    mods = {"dgl": dgl, "dgl.geometry": geo, "moviepy": mp, "moviepy.editor": mpe, "sim": sim, "sim.utils": simu}
    try:
        import cv2  # noqa: F401
    except Exception:
        mods["cv2"] = types.ModuleType("cv2")
    sys.modules.update(mods)


def ids(R):
    idx = R.argmax(-1).to(torch.int32)
    idx[R.sum(-1) == 0] = -1
    return idx.numpy()


def main():
    from agx_helpers import dataset_configs, write_synthetic_dataset
    stub_modules()
    sys.path.insert(0, REF)
    import dynamics.dataset.dataset as ref_ds
    from dynamics.dataset.load import load_dataset, load_positions
    out = {}
    with tempfile.TemporaryDirectory() as root:
        write_synthetic_dataset(root)
        kept = []
        real_fps = ref_ds.fps

        def spy(*a, **k):                      # records the kept particle indices of every sample
            r = real_fps(*a, **k)
            kept.append(np.asarray(r))
            return r
        ref_ds.fps = spy
        for tag, kw in [("noise", {}), ("plain", {"state_noise": 0.0, "fps_radius_range": 0.2}), ("phys", {"phys_noise": 0.05})]:
            dc, mc = dataset_configs(root, **kw)
            if tag == "plain":
                dc["randomness"]["use"] = False
            for phase in ["train", "valid"]:
                pairs, phys = load_dataset(dc, mc, phase)
                out[f"{tag}/{phase}/pairs"] = np.asarray(pairs)
                out[f"{tag}/{phase}/physics"] = np.stack([p["rope"] for p in phys])
            eef, obj = load_positions(dc)
            out[f"{tag}/n_obj"] = np.asarray([o.shape[1] for o in obj])
            ds = ref_ds.DynDataset(dc, mc, "train")
            order = [3, 0, 17, 5, 5, len(ds) - 1]
            out[f"{tag}/order"] = np.asarray(order)
            np.random.seed(1234)
            kept.clear()
            for j, i in enumerate(order):
                g = ds[i]
                for k, v in g.items():
                    if k in ("Rr", "Rs"):
                        out[f"{tag}/s{j}/{k}_ids"] = ids(v)
                    else:
                        out[f"{tag}/s{j}/{k}"] = v.numpy().copy()       # copy: the physics tensor aliases the data set's own array
                out[f"{tag}/s{j}/kept"] = kept[j].astype(np.int32)
        # the order a seeded shuffled DataLoader visits the samples in (train.py:44-49 with num_workers = 0)
        dc, mc = dataset_configs(root)
        ds = ref_ds.DynDataset(dc, mc, "train")

        class IndexOnly(torch.utils.data.Dataset):
            def __len__(self):
                return len(ds)

            def __getitem__(self, i):
                return i
        torch.manual_seed(42)
        loader = torch.utils.data.DataLoader(IndexOnly(), batch_size=16, shuffle=True, num_workers=0)
        out["loader/n"] = np.int64(len(ds))
        out["loader/epoch0"] = np.concatenate([b.numpy() for b in loader])
        out["loader/epoch1"] = np.concatenate([b.numpy() for b in loader])
    np.savez_compressed(os.path.join(HERE, "dataset_rope.npz"), **out)
    print("wrote dataset_rope.npz with", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "dataset_rope.npz")), "bytes")


if __name__ == "__main__":
    main()
