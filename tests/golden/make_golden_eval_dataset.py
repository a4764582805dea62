"""Generate tests/golden/eval_dataset.npz by running the UNMODIFIED reference `rollout_dataset`
(src/dynamics/rollout/rollout.py:205-268, with rollout_episode_pushes / construct_graph / rollout_from_start_graph under it)
on the synthetic on-disk data set of tests/agx_helpers.write_synthetic_dataset (build container only):

    python tests/golden/make_golden_eval_dataset.py

The reference's own DynamicsPredictor runs on the CPU with the weights of tests/golden/weights_seed0.npz.  Stubbed because absent
from the image and irrelevant to the numbers: dgl's sampler (oracle/sampling_oracle's restatement), moviepy, matplotlib (every
pyplot call is a no-op), sim.data_gen.data.  Stored: `error_short.txt` as the reference wrote it and every per-push error file.
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


class Anything:
    def __getattr__(self, name):
        return Anything()

    def __call__(self, *a, **k):
        return Anything()


def import_reference():
    from oracle import sampling_oracle as so
    sys.path.insert(0, REF)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def sampler(pos, npoints, start_idx=-1):
        return torch.from_numpy(so.farthest_point_sampler(pos.numpy(), npoints, [start_idx] * pos.shape[0]))
    geo = mod("dgl.geometry", farthest_point_sampler=sampler)
    mod("dgl", geometry=geo)
    mod("moviepy", editor=mod("moviepy.editor"))
    mod("matplotlib", use=lambda *a, **k: None, pyplot=mod("matplotlib.pyplot", __getattr__=lambda name: Anything()))
    mod("sim.data_gen.data", load_data=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub")))
    from dynamics.gnn.model import DynamicsPredictor
    from dynamics.rollout.rollout import rollout_dataset
    return DynamicsPredictor, rollout_dataset


def main():
    from adaptigraph_b200 import synthetic as syn
    from agx_helpers import dataset_configs, golden_weights, write_synthetic_dataset
    DP, rollout_dataset = import_reference()
    torch.set_num_threads(4)
    model_config = syn.configs("rope")[0]
    out = {}
    with tempfile.TemporaryDirectory() as root:
        write_synthetic_dataset(root)
        dc, mc = dataset_configs(root)
        model = DP(model_config, mc, dc, "cpu").eval()
        model.load_state_dict(golden_weights())
        save_dir = os.path.join(root, "rollout_out")
        os.makedirs(save_dir)
        np.random.seed(5)
        rollout_dataset(model, "cpu", {"dataset_config": dc, "material_config": mc}, save_dir, False)
        out["error_short"] = np.loadtxt(os.path.join(save_dir, "error_short.txt"))
        for ep in sorted(os.listdir(save_dir)):
            d = os.path.join(save_dir, ep, "short")
            if not os.path.isdir(d):
                continue
            for f in sorted(os.listdir(d)):
                if f.endswith(".txt"):
                    out[f"{ep}/{f[:-4]}"] = np.atleast_1d(np.loadtxt(os.path.join(d, f)))
    for k, v in out.items():
        print(k, v.shape, np.round(v.reshape(-1)[:3], 5))
    np.savez_compressed(os.path.join(HERE, "eval_dataset.npz"), **out)


if __name__ == "__main__":
    main()
