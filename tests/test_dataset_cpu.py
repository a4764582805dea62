"""Host side of the training data path (adaptigraph_b200/dataset.py) against what the UNMODIFIED reference returns on the same
on-disk data set (tests/golden/dataset_rope.npz, written by tests/golden/make_golden_dataset.py): the file loaders, the sample
recipe with the reference's random draws, and the sample order of the loader.  The two device stages (particle thinning,
relations) are substituted here — tests/test_dataset_gpu.py runs them."""
import numpy as np
import pytest
import torch

from agx_helpers import dataset_configs, load_npz, write_synthetic_dataset

G = load_npz("dataset_rope.npz")
VARIANTS = {"noise": {}, "plain": {"state_noise": 0.0, "fps_radius_range": 0.2}, "phys": {"phys_noise": 0.05}}


def configs(root, tag):
    dc, mc = dataset_configs(root, **VARIANTS[tag])
    if tag == "plain":
        dc["randomness"]["use"] = False
    return dc, mc


@pytest.fixture(scope="module")
def root(tmp_path_factory):
    r = str(tmp_path_factory.mktemp("agx_dataset"))
    write_synthetic_dataset(r)
    return r


def host_dataset(root, tag):
    """DynDataset with the device stages replaced: thinning returns the reference's kept indices, relations are not built."""
    from adaptigraph_b200.dataset import DynDataset
    from adaptigraph_b200.graph import EdgeList

    class HostOnly(DynDataset):
        served = 0

        def _thin(self, clouds, start, radius, start2):
            out = []
            for c, s, s2 in zip(clouds, start, start2):
                kept = G[f"{tag}/s{self.served}/kept"]
                assert 0 <= s < c.shape[0] and 0 <= s2 < min(self.max_nobj, c.shape[0])
                assert kept.max() < c.shape[0]
                out.append(kept)
                self.served += 1
            return out

        def _relations(self, state_last, state_mask, eef_mask, adj_thresh):
            z = torch.zeros(1, dtype=torch.int32)
            return EdgeList(z, z, z, z, z, state_last.shape[0], state_last.shape[1])
    dc, mc = configs(root, tag)
    return HostOnly(dc, mc, "train", device="cpu")


@pytest.mark.parametrize("tag", sorted(VARIANTS))
def test_loaders_match_reference(root, tag):
    from adaptigraph_b200.dataset import load_dataset, load_positions
    dc, mc = configs(root, tag)
    for phase in ["train", "valid"]:
        pairs, phys = load_dataset(dc, mc, phase)
        assert pairs.dtype == G[f"{tag}/{phase}/pairs"].dtype and np.array_equal(pairs, G[f"{tag}/{phase}/pairs"])
        got = np.stack([p["rope"] for p in phys])
        assert got.dtype == np.float32 and np.array_equal(got, G[f"{tag}/{phase}/physics"])
    eef, obj = load_positions(dc)
    assert [o.shape[1] for o in obj] == list(G[f"{tag}/n_obj"]) and len(eef) == len(obj)


def check_sample(tag, j, sample, skip=("edges", "state_mask", "eef_mask", "adj_thresh")):
    want = {k.split("/")[-1]: v for k, v in G.items() if k.startswith(f"{tag}/s{j}/") and not k.endswith(("_ids", "/kept"))}
    assert set(want) == set(sample) - set(skip) - {"Rr", "Rs"}
    for k, v in want.items():
        got = sample[k].cpu().numpy()
        assert got.dtype == v.dtype and got.shape == v.shape, (k, got.dtype, v.dtype, got.shape, v.shape)
        assert np.array_equal(got, v), f"{tag} sample {j}: {k} differs by {np.abs(got.astype(np.float64) - v).max()}"


@pytest.mark.parametrize("tag", sorted(VARIANTS))
def test_samples_match_reference_one_by_one(root, tag):
    ds = host_dataset(root, tag)
    np.random.seed(1234)
    for j, i in enumerate(G[f"{tag}/order"]):
        check_sample(tag, j, ds[int(i)])


@pytest.mark.parametrize("tag", sorted(VARIANTS))
def test_samples_match_reference_as_one_batch(root, tag):
    """A batch draws its random numbers sample by sample in the reference's order, so it equals the stacked samples."""
    ds = host_dataset(root, tag)
    np.random.seed(1234)
    order = [int(i) for i in G[f"{tag}/order"]]
    batch = ds[order]
    assert batch["state"].shape[0] == len(order)
    for j in range(len(order)):
        check_sample(tag, j, {k: (v if k == "edges" else v[j]) for k, v in batch.items()})
    # masks: object rows kept by the thinning, then the tool rows (dataset.py:150-158)
    n_kept = [len(G[f"{tag}/s{j}/kept"]) for j in range(len(order))]
    assert batch["state_mask"].sum(1).tolist() == [k + 1 for k in n_kept] and batch["eef_mask"].sum(1).tolist() == [1] * len(order)
    lo, hi = ds.adj_radius_range
    assert bool(((batch["adj_thresh"] >= lo) & (batch["adj_thresh"] <= hi)).all())


def test_loader_visits_samples_in_the_reference_order():
    from adaptigraph_b200.dataset import make_loader

    class Indices(torch.utils.data.Dataset):
        def __len__(self):
            return int(G["loader/n"])

        def __getitem__(self, idx):
            return list(idx)
    torch.manual_seed(42)
    loader = make_loader(Indices(), 16, shuffle=True)
    for epoch in ["epoch0", "epoch1"]:
        batches = list(loader)
        assert [len(b) for b in batches] == [16] * (int(G["loader/n"]) // 16) + ([int(G["loader/n"]) % 16] if int(G["loader/n"]) % 16 else [])
        assert np.array_equal(np.concatenate(batches), G[f"loader/{epoch}"])
    assert [list(b) for b in make_loader(Indices(), 20, shuffle=False)][0] == list(range(20))


def test_single_row_pair_files_are_skipped(root):
    from adaptigraph_b200.dataset import load_pairs
    import os
    dc, _ = configs(root, "noise")
    pairs = load_pairs(os.path.join(dc["prep_data_dir"], "rope", "frame_pairs"), range(2))
    assert pairs.shape == (84, 8) and set(pairs[:, 0]) == {0, 1}       # 24 + 18 rows per episode; the one-row third file adds none
