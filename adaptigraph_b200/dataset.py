"""Training data in front of the dynamics path: the reference's on-disk formats and its sample
recipe, emitting SPARSE relations built on the device instead of dense padded one-hots built on the host.

    load_pairs / load_dataset / load_positions      src/dynamics/dataset/load.py:6-83  (frame_pairs/*.txt, property_params.pkl,
                                                    positions.pkl — read exactly as the reference reads them)
    DynDataset(dataset_config, material_config,     src/dynamics/dataset/dataset.py:10-252: same constructor, same keys, same
               phase)                               numpy random draws in the same order per sample
    make_loader(dataset, batch_size, shuffle)       the DataLoader of train.py:44-49 (same sampler classes, so a seeded run
                                                    visits the samples in the reference's order)

What differs from the reference, by design:
* a batch is assembled at once (`DynDataset.get_batch`, or `dataset[list_of_indices]` through `make_loader`): the farthest-point
  thinning of all its samples is ONE pair of `agx_fps` launches (per-sample start index and radius), and the relations of all its
  samples ONE `agx_graph_build` call with the single-graph semantics of `construct_edges_from_states` (graph.py:38-89) on the
  uploaded `state[:, -1]` and the per-sample `adj_thresh` draw;
* the batch carries an `EdgeList` (`batch["edges"]`, CSR by receiver, capacity B * max_nR — exceeding max_nR raises like
  `pad_torch`, utils.py:37-46) and NO `Rr` / `Rs`: 2 * max_nR * N floats per sample (0.8 MB for the rope config) never exist on
  the host, the wire or the device.  `dense=True` adds the reference's padded `Rr` / `Rs` (from the same EdgeList) for callers
  that have not switched.
Everything else in `__getitem__` (gathers, zero padding, masks, noise, the xy rotation) is host-side data preparation on a few
kilobytes per sample, done with the same numpy operations as the reference so that a seeded run reproduces its tensors bit
for bit (the rotation is a float32 numpy matmul; re-implementing it elsewhere could change the rounding of its 3-term sums).
There is no CPU path for the sampling or the relations: the dataset needs a CUDA device.
"""
from __future__ import annotations

import glob
import os
import pickle
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib as L
from . import ops
from .graph import EdgeList
from .sampling import fps_batch


# ------------------------------------------------------------------------------------------ on-disk formats (load.py)
def load_pairs(pairs_path: str, episode_range) -> np.ndarray:
    """load.py:6-17: every `<episode:06>_<push:02>.txt` of the episodes in range -> (n, 1 + n_his + n_future) int array
    [episode, frame indices...]; files holding a single row (1-D after loadtxt) are skipped, as the reference skips them."""
    rows: List[np.ndarray] = []
    for ep in episode_range:
        n_pushes = len(glob.glob(os.path.join(pairs_path, f"{ep:06}_*.txt")))
        for push in range(1, n_pushes + 1):
            frame_pairs = np.loadtxt(os.path.join(pairs_path, f"{ep:06}_{push:02}.txt"))
            if frame_pairs.ndim == 1:
                continue
            rows.extend(np.concatenate([np.full((frame_pairs.shape[0], 1), float(ep)), frame_pairs], axis=1))
    return np.array(rows).astype(int)


def load_dataset(dataset_config: dict, material_config: dict, phase: str = "train"):
    """load.py:19-66: (frame pairs of the phase's episode slice, per-episode dict material -> normalised physics parameters)."""
    data_name = dataset_config["data_name"]
    data_dir = os.path.join(dataset_config["data_dir"], data_name)
    prep_dir = os.path.join(dataset_config["prep_data_dir"], data_name)
    lo, hi = dataset_config["ratio"][phase]
    num_epis = len([f for f in os.listdir(data_dir) if os.path.isdir(os.path.join(data_dir, f)) and f.isdigit()])
    pair_lists = load_pairs(os.path.join(prep_dir, "frame_pairs"), range(int(num_epis * lo), int(num_epis * hi)))
    physics_params = []
    for ep in range(num_epis):
        with open(os.path.join(data_dir, f"{ep:06}", "property_params.pkl"), "rb") as f:
            properties = pickle.load(f)
        per_material = {}
        for name in dataset_config["materials"]:
            used = [(properties[item["name"]] - item["min"]) / (item["max"] - item["min"] + 1e-6)
                    for item in material_config[name]["physics_params"] if item["name"] in properties.keys() and item["use"]]
            per_material[name] = np.array(used).astype(np.float32) * (1.0 - 0.0) + 0.0      # phys_norm_max / _min of load.py:49-50
        physics_params.append(per_material)
    return pair_lists, physics_params


def load_positions(dataset_config: dict):
    """load.py:68-83: positions.pkl -> (eef_pos, obj_pos), per episode (T, N_eef, 3) and (T, N_obj, 3)."""
    prep_dir = os.path.join(dataset_config["prep_data_dir"], dataset_config["data_name"])
    with open(os.path.join(prep_dir, "positions.pkl"), "rb") as f:
        positions = pickle.load(f)
    return positions["eef_pos"], positions["obj_pos"]


# ------------------------------------------------------------------------------------------ the dataset (dataset.py)
class DynDataset(torch.utils.data.Dataset):
    """dataset.py:10-252 with sparse, device-built relations.  `dataset[i]` is one sample as a batch of one with the batch axis
    removed; `dataset[[i, j, ...]]` / `get_batch` a whole batch (what `make_loader` requests)."""

    def __init__(self, dataset_config: dict, material_config: dict, phase: str = "train", device=None, dense: bool = False):
        assert phase in ["train", "valid"]
        self.phase, self.dataset_config, self.material_config = phase, dataset_config, material_config
        self.verbose = dataset_config["verbose"]
        self.n_his, self.n_future = dataset_config["n_his"], dataset_config["n_future"]
        self.add_randomness = dataset_config["randomness"]["use"]
        self.state_noise = dataset_config["randomness"]["state_noise"][phase]
        self.phys_noise = dataset_config["randomness"]["phys_noise"][phase]
        assert len(dataset_config["datasets"]) == 1, "Only one object type is supported."
        d = self.dataset = dataset_config["datasets"][0]
        self.max_nobj, self.fps_radius_range = d["max_nobj"], d["fps_radius_range"]
        self.max_nR, self.adj_radius_range = d["max_nR"], d["adj_radius_range"]
        self.topk, self.connect_tool_all = d["topk"], d["connect_tool_all"]
        self.pair_lists, self.physics_params = load_dataset(dataset_config, material_config, phase)
        self.pair_lists = np.array(self.pair_lists)
        self.materials = {k: v.shape[0] for k, v in self.physics_params[0].items()}
        self.eef_pos, self.obj_pos = load_positions(dataset_config)
        self.pos_dim = self.obj_pos[0].shape[-1]
        self.obj_dim = self.max_nobj
        self.eef_dim = self.eef_pos[0].shape[1]
        self.state_dim = self.obj_dim + self.eef_dim
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dense = dense

    def __len__(self) -> int:
        return len(self.pair_lists)

    # ---- the two device stages; tests substitute `_thin` to check the host recipe without a GPU
    def _thin(self, clouds: List[np.ndarray], start: np.ndarray, radius: np.ndarray, start2: np.ndarray) -> List[np.ndarray]:
        """graph.py:8-36 for every cloud of the batch: farthest-point sample to max_nobj particles from `start`, then thin to
        `radius` from survivor `start2`.  Returns the kept indices into each cloud, in selection order."""
        B, n_max = len(clouds), max(c.shape[0] for c in clouds)
        pos = np.zeros((B, n_max, 3), np.float32)
        for b, c in enumerate(clouds):
            pos[b, : c.shape[0]] = c
        dev = self.device
        n_pts = torch.tensor([c.shape[0] for c in clouds], dtype=torch.int32, device=dev)
        idx, cnt = fps_batch(torch.from_numpy(pos).to(dev), n_pts, self.max_nobj,
                             torch.from_numpy(np.asarray(radius, np.float64)).to(dev),
                             torch.from_numpy(np.asarray(start, np.int32)).to(dev), torch.from_numpy(np.asarray(start2, np.int32)).to(dev))
        idx, cnt = idx.cpu().numpy(), cnt.cpu().numpy()
        return [idx[b, : cnt[b]].astype(np.int32) for b in range(B)]

    def _relations(self, state_last: torch.Tensor, state_mask: torch.Tensor, eef_mask: torch.Tensor, adj_thresh: np.ndarray) -> EdgeList:
        """dataset.py:214-216 for the whole batch: construct_edges_from_states on every sample's last history frame."""
        B, N, _ = state_last.shape
        thr2 = torch.from_numpy((adj_thresh * adj_thresh).astype(np.float32)).to(state_last.device)    # graph.py:53: squared as a Python float
        row_ptr, send, recv, n_edges, status = ops.graph_build(state_last, state_mask, eef_mask, thr2, self.topk, self.connect_tool_all,
                                                               L.AGX_SEM_SINGLE, B * self.max_nR)
        worst = int(n_edges.max().item())
        if worst > self.max_nR:                                                                        # pad_torch (utils.py:37-46) fails the same way
            raise RuntimeError(f"a sample has {worst} relations, more than max_nR = {self.max_nR}")
        return EdgeList(row_ptr, send, recv, n_edges, status, B, N)

    def get_batch(self, indices: Sequence[int]) -> Dict[str, Union[torch.Tensor, EdgeList]]:
        n_his, n_future, max_nobj = self.n_his, self.n_future, self.max_nobj
        B = len(indices)
        episodes, obj_all, eef_all = [], [], []
        # ---- the reference's random numbers, in its order: per sample [FPS start, FPS radius, thinning start], physics noise,
        # [state noise, rotation], adjacency radius (graph.py:11-23, utils.py:14, dataset.py:178-179, :188-190, :213)
        start, radius, start2, rots, adj = np.zeros(B, np.int64), np.zeros(B, np.float64), np.zeros(B, np.int64), np.zeros(B), np.zeros(B)
        noises: List[Optional[np.ndarray]] = []
        phys: List[Dict[str, np.ndarray]] = []
        for b, idx in enumerate(indices):
            ep = self.pair_lists[idx][0].astype(int)
            pair = self.pair_lists[idx][1:].astype(int)
            assert len(pair) == n_his + n_future
            episodes.append(ep)
            obj_all.append(np.array([self.obj_pos[ep][f] for f in pair]))      # (T, N_obj_all, 3)
            eef_all.append(np.array([self.eef_pos[ep][f] for f in pair]))      # (T, N_eef, 3)
            n_all = obj_all[b].shape[1]
            start[b] = np.random.randint(0, n_all)
            if type(self.fps_radius_range) == float:
                radius[b] = self.fps_radius_range
            elif len(self.fps_radius_range) == 2:
                radius[b] = np.random.uniform(self.fps_radius_range[0], self.fps_radius_range[1])
            else:
                raise ValueError(f"Invalid fps_radius_range: {self.fps_radius_range}.")
            start2[b] = np.random.randint(min(max_nobj, n_all))
            physics_param = self.physics_params[ep]
            for name in self.dataset_config["materials"]:
                if name not in physics_param.keys():
                    raise ValueError(f'Physics parameter {name} not found in {self.dataset_config["data_dir"]}')
                # in place, as dataset.py:178: the episode's stored parameters accumulate the noise of every visit
                physics_param[name] += np.random.uniform(-self.phys_noise, self.phys_noise, size=physics_param[name].shape)
            phys.append({k: v.copy() for k, v in physics_param.items()})
            if self.add_randomness:
                noises.append(np.random.uniform(-self.state_noise, self.state_noise, size=(n_his, self.state_dim, self.pos_dim)))
                rots[b] = np.random.uniform(-np.pi, np.pi)
            else:
                noises.append(None)
            adj[b] = np.random.uniform(*self.adj_radius_range)
        kept = self._thin([o[n_his - 1] for o in obj_all], start, radius, start2)
        # ---- host assembly (dataset.py:96-206)
        N = self.state_dim
        out = {
            "state": np.zeros((B, n_his, N, self.pos_dim), np.float32), "action": np.zeros((B, N, self.pos_dim), np.float32),
            "eef_future": np.zeros((B, n_future - 1, N, self.pos_dim), np.float32),
            "action_future": np.zeros((B, n_future - 1, N, self.pos_dim), np.float32),
            "state_future": np.zeros((B, n_future, max_nobj, self.pos_dim), np.float32),
            "attrs": np.zeros((B, N, 2), np.float32), "p_rigid": np.zeros((B, 1), np.float32),
            "p_instance": np.zeros((B, max_nobj, 1), np.float32), "obj_mask": np.zeros((B, max_nobj), bool),
            "material_index": np.zeros((B, max_nobj, len(self.material_config["material_index"])), np.int64),
        }
        state_mask, eef_mask = np.zeros((B, N), bool), np.zeros((B, N), bool)
        assert len(self.dataset_config["materials"]) == 1, "only support single material"
        mat_col = self.material_config["material_index"][self.dataset_config["materials"][0]]
        for b in range(B):
            k, eef = len(kept[b]), eef_all[b]
            n_eef = eef.shape[1]
            obj = np.zeros((n_his + n_future, max_nobj, self.pos_dim), np.float32)         # pad(obj_kps[:, fps_idx], max_nobj, dim=1)
            obj[:, :k] = obj_all[b][:, kept[b]]
            state = out["state"][b]
            state[:, :max_nobj] = obj[:n_his]
            state[:, max_nobj:max_nobj + n_eef] = eef[:n_his]
            out["action"][b, max_nobj:max_nobj + n_eef] = eef[n_his] - eef[n_his - 1]
            out["state_future"][b] = obj[n_his:]
            for fi in range(n_future - 1):
                out["eef_future"][b, fi, max_nobj:max_nobj + n_eef] = eef[n_his + fi]
                out["action_future"][b, fi, max_nobj:max_nobj + n_eef] = eef[n_his + fi + 1] - eef[n_his + fi]
            state_mask[b, :k] = True
            state_mask[b, max_nobj:max_nobj + n_eef] = True
            eef_mask[b, max_nobj:max_nobj + n_eef] = True
            out["obj_mask"][b, :k] = True
            out["attrs"][b, :k, 0] = 1.0
            out["attrs"][b, max_nobj:max_nobj + n_eef, 1] = 1.0
            out["p_instance"][b, :k, 0] = 1
            out["material_index"][b, :k, mat_col] = 1
            if self.add_randomness:                                                           # dataset.py:187-199
                state += noises[b]
                rot_mat = np.array([[np.cos(rots[b]), -np.sin(rots[b]), 0], [np.sin(rots[b]), np.cos(rots[b]), 0], [0, 0, 1]], dtype=state.dtype)
                out["state"][b] = state @ rot_mat[None]
                out["action"][b] = out["action"][b] @ rot_mat
                out["eef_future"][b] = out["eef_future"][b] @ rot_mat[None]
                out["action_future"][b] = out["action_future"][b] @ rot_mat[None]
                out["state_future"][b] = out["state_future"][b] @ rot_mat[None]
        dev = self.device
        batch: Dict[str, Union[torch.Tensor, EdgeList]] = {k: torch.from_numpy(v).to(dev) for k, v in out.items()}
        batch["state_mask"], batch["eef_mask"] = torch.from_numpy(state_mask).to(dev), torch.from_numpy(eef_mask).to(dev)
        batch["adj_thresh"] = torch.from_numpy(adj).to(dev)
        for name, dim in self.materials.items():                                              # dataset.py:245-250
            vals = [torch.from_numpy(p[name]).float() if name in p else torch.zeros(dim) for p in phys]
            batch[name + "_physics_param"] = torch.stack(vals).to(dev)
        edges = self._relations(batch["state"][:, -1].contiguous(), batch["state_mask"], batch["eef_mask"], adj)
        batch["edges"] = edges
        if self.dense:
            batch["Rr"], batch["Rs"] = edges.to_dense(self.max_nR)
        return batch

    def __getitem__(self, idx):
        if isinstance(idx, (list, tuple, np.ndarray)) or (torch.is_tensor(idx) and idx.dim() > 0):
            return self.get_batch([int(i) for i in idx])
        batch = self.get_batch([int(idx)])
        return {k: (v if isinstance(v, EdgeList) else v[0]) for k, v in batch.items()}


def make_loader(dataset: DynDataset, batch_size: int, shuffle: bool) -> torch.utils.data.DataLoader:
    """The DataLoader of train.py:44-49 (`batch_size`, `shuffle = phase == 'train'`), asking the dataset for whole batches:
    the same RandomSampler / SequentialSampler and BatchSampler classes a `DataLoader(dataset, batch_size, shuffle)` builds
    internally, so the torch RNG is consumed identically and a seeded run draws the reference's batches.  No worker processes:
    the per-batch device work replaces what the reference spreads over `num_workers` CPU processes."""
    from torch.utils.data import BatchSampler, DataLoader, RandomSampler, SequentialSampler
    sampler = RandomSampler(dataset) if shuffle else SequentialSampler(dataset)
    return DataLoader(dataset, batch_size=None, sampler=BatchSampler(sampler, batch_size, drop_last=False), num_workers=0,
                      collate_fn=lambda batch: batch)
