"""Shared test helpers: golden loading and edge-format conversion (CPU only)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def golden_weights():
    return {k: torch.from_numpy(v) for k, v in load_npz("weights_seed0.npz").items()}


def onehots_from_lists(recv, send, N):
    """int (B, n_rel) receiver/sender ids with -1 padding -> dense f32 (B, n_rel, N) Rr, Rs."""
    recv, send = torch.as_tensor(recv).long(), torch.as_tensor(send).long()
    B, R = recv.shape
    Rr, Rs = torch.zeros(B, R, N), torch.zeros(B, R, N)
    b, e = torch.nonzero(recv >= 0, as_tuple=True)
    Rr[b, e, recv[b, e]] = 1
    Rs[b, e, send[b, e]] = 1
    return Rr, Rs


def lists_from_onehots(Rr, Rs):
    def one(R):
        idx = R.argmax(-1).to(torch.int32)
        idx[R.sum(-1) == 0] = -1
        return idx
    return one(Rr), one(Rs)


def csr_from_lists(recv, send, N):
    """-1 padded (B, n_rel) lists (rows sorted by receiver, as the reference emits) -> CSR over B*N rows."""
    recv, send = torch.as_tensor(recv).long(), torch.as_tensor(send).long()
    B = recv.shape[0]
    deg = torch.zeros(B * N, dtype=torch.long)
    sends = []
    for b in range(B):
        ok = recv[b] >= 0
        r, s = recv[b][ok], send[b][ok]
        assert bool((r[1:] >= r[:-1]).all()), "relation rows must be receiver-sorted"
        deg.index_add_(0, r + b * N, torch.ones_like(r))
        sends.append(s)
    row_ptr = torch.zeros(B * N + 1, dtype=torch.int32)
    row_ptr[1:] = torch.cumsum(deg, 0).to(torch.int32)
    return row_ptr, torch.cat(sends).to(torch.int32)


def graph_case_names(cases):
    return sorted({k.split("/")[0] for k in cases})


# ---------------------------------------------------------------------------------------------- synthetic on-disk dataset
def dataset_configs(root, state_noise=0.05, phys_noise=0.0, fps_radius_range=(0.18, 0.22), max_nobj=60, max_nR=600):
    """dataset_config / material_config in the shape of src/config/dynamics/rope.yaml, pointing at `root`."""
    dataset_config = {
        "data_name": "rope", "materials": ["rope"], "data_dir": os.path.join(root, "sim_data"),
        "prep_data_dir": os.path.join(root, "preprocess"), "verbose": False, "n_his": 4, "n_future": 3,
        "ratio": {"train": [0, 0.75], "valid": [0.75, 1]},
        "datasets": [{"name": "rope", "max_nobj": max_nobj, "max_nR": max_nR,
                      "fps_radius_range": list(fps_radius_range) if not isinstance(fps_radius_range, float) else fps_radius_range,
                      "adj_radius_range": [0.48, 0.52], "topk": 10, "connect_tool_all": False}],
        "randomness": {"use": True, "state_noise": {"train": state_noise, "valid": 0.0}, "phys_noise": {"train": phys_noise, "valid": 0.0}},
    }
    material_config = {"material_index": {"rope": 0}, "rope": {"physics_params": [
        {"name": "particle_radius", "use": False, "min": 0.0, "max": 1.0},
        {"name": "stiffness", "use": True, "min": 0.0, "max": 1.0}]}}
    return dataset_config, material_config


def write_synthetic_dataset(root, n_episodes=8, seed=11):
    """A small rope data set in the reference's on-disk layout (load.py): sim_data/rope/<episode:06>/property_params.pkl,
    preprocess/rope/positions.pkl ({'eef_pos': [...], 'obj_pos': [...]}) and preprocess/rope/frame_pairs/<episode:06>_<push:02>.txt
    (rows of n_his + n_future frame indices; one file per episode has a single row, which the loader must skip)."""
    import pickle
    rng = np.random.default_rng(seed)
    data_dir, prep_dir = os.path.join(root, "sim_data", "rope"), os.path.join(root, "preprocess", "rope")
    os.makedirs(os.path.join(prep_dir, "frame_pairs"), exist_ok=True)
    eef_pos, obj_pos = [], []
    for ep in range(n_episodes):
        os.makedirs(os.path.join(data_dir, f"{ep:06}"), exist_ok=True)
        with open(os.path.join(data_dir, f"{ep:06}", "property_params.pkl"), "wb") as f:
            pickle.dump({"particle_radius": 0.03, "stiffness": float(rng.uniform(0.1, 0.9)), "length": 3.0}, f)
        T, n_obj = 30, int(rng.integers(150, 260))
        s = np.linspace(0, 6.0, n_obj)
        base = np.stack([s, np.zeros(n_obj), 0.4 * np.sin(s + rng.uniform(0, 3))], -1)
        drift = np.cumsum(rng.normal(0, 0.01, (T, 1, 3)), 0) * np.linspace(0.2, 1.0, n_obj)[None, :, None]
        obj = (base[None] + drift + rng.normal(0, 0.004, (T, n_obj, 3))).astype(np.float32 if ep % 2 else np.float64)
        eef = (np.array([[3.0, 0.3, 0.5]]) + np.cumsum(rng.normal(0, 0.02, (T, 1, 3)), 0)).astype(np.float64)
        obj_pos.append(obj)
        eef_pos.append(eef)
        for push in range(1, 3):      # rows [f - 3s .. f + 3s] at frame stride s = push: the pair after (f, f + s) exists, so rollouts chain
            rows = np.stack([f + push * np.arange(-3, 4) for f in range(3 * push, T - 3 * push)]).astype(float)
            np.savetxt(os.path.join(prep_dir, "frame_pairs", f"{ep:06}_{push:02}.txt"), rows)
        if ep < 6:    # one row: load_pairs skips the file (the evaluation driver of the reference cannot read one: validation episodes have none)
            np.savetxt(os.path.join(prep_dir, "frame_pairs", f"{ep:06}_03.txt"), (np.arange(7) + 2).astype(float)[None])
    with open(os.path.join(prep_dir, "positions.pkl"), "wb") as f:
        pickle.dump({"eef_pos": eef_pos, "obj_pos": obj_pos}, f)
