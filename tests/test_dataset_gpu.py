"""The training data path on the GPU (adaptigraph_b200/dataset.py): particle thinning (agx_fps / agx_fps_radii) and relations
(agx_graph_build, single-graph semantics) of whole batches against what the UNMODIFIED reference's DynDataset returns for the
same samples under the same seeds (tests/golden/dataset_rope.npz), and the epoch loop of train.py on that data set."""
import os

import numpy as np
import pytest
import torch

from agx_helpers import dataset_configs, load_npz, write_synthetic_dataset
from test_dataset_cpu import VARIANTS, check_sample, configs

pytestmark = pytest.mark.gpu
G = load_npz("dataset_rope.npz")


@pytest.fixture(scope="module")
def root(tmp_path_factory):
    r = str(tmp_path_factory.mktemp("agx_dataset_gpu"))
    write_synthetic_dataset(r)
    return r


def check_relations(tag, j, recv_ids, send_ids):
    want_r, want_s = G[f"{tag}/s{j}/Rr_ids"], G[f"{tag}/s{j}/Rs_ids"]
    assert np.array_equal(recv_ids, want_r) and np.array_equal(send_ids, want_s), f"{tag} sample {j}: relation rows differ"


@pytest.mark.parametrize("tag", sorted(VARIANTS))
def test_batch_equals_reference_samples(root, tag):
    from adaptigraph_b200.dataset import DynDataset
    from adaptigraph_b200.graph import relation_lists
    dc, mc = configs(root, tag)
    ds = DynDataset(dc, mc, "train", dense=True)
    order = [int(i) for i in G[f"{tag}/order"]]
    np.random.seed(1234)
    batch = ds[order]
    assert batch["state"].is_cuda and batch["edges"].B == len(order)
    n_edges = batch["edges"].n_edges.cpu().numpy()
    for j in range(len(order)):
        sample = {k: (v if k == "edges" else v[j]) for k, v in batch.items()}
        check_sample(tag, j, sample)
        kept = G[f"{tag}/s{j}/kept"]
        assert int(sample["obj_mask"].sum()) == len(kept)
        r, s = relation_lists(sample["Rr"], sample["Rs"])           # dense rows in the reference's order, zero rows = padding
        check_relations(tag, j, r.cpu().numpy(), s.cpu().numpy())
        assert n_edges[j] == int((G[f"{tag}/s{j}/Rr_ids"] >= 0).sum())
    # one by one: same draws, same samples
    ds1 = DynDataset(dc, mc, "train", dense=True)
    np.random.seed(1234)
    for j, i in enumerate(order):
        sample = ds1[i]
        check_sample(tag, j, sample)
        r, s = relation_lists(sample["Rr"], sample["Rs"])
        check_relations(tag, j, r.cpu().numpy(), s.cpu().numpy())


def test_relation_capacity_overflow_raises(root):
    from adaptigraph_b200.dataset import DynDataset
    dc, mc = dataset_configs(root, max_nR=50)
    ds = DynDataset(dc, mc, "train")
    np.random.seed(0)
    with pytest.raises(RuntimeError, match="max_nR"):
        ds[[0, 1]]


def test_epoch_loop_trains_and_writes_the_reference_checkpoints(root, tmp_path):
    import adaptigraph_b200 as agx
    from adaptigraph_b200 import synthetic
    from adaptigraph_b200.train import train
    dc, mc = dataset_configs(root)
    config = {
        "dataset_config": dc, "material_config": mc,
        "model_config": synthetic.configs("rope")[0],
        "train_config": {"out_dir": str(tmp_path), "phases": ["train", "valid"], "num_workers": 0, "random_seed": 42, "verbose": False,
                         "batch_size": 8, "n_epochs": 3, "n_iters_per_epoch": {"train": 12, "valid": -1}, "log_interval": 1},
    }
    lines = []
    curves = train(config, log=lines.append)
    assert len(curves["train"]) == 3 and len(curves["valid"]) == 3 and all(np.isfinite(curves["train"] + curves["valid"]))
    assert curves["train"][-1] < curves["train"][0] and curves["valid"][-1] < curves["valid"][0]
    ck = os.path.join(str(tmp_path), "rope", "checkpoints")
    assert sorted(os.listdir(ck)) == ["latest.pth", "latest_optim.pth"]
    sd = torch.load(os.path.join(ck, "latest.pth"))
    model = agx.DynamicsPredictor(config["model_config"], mc, dc, torch.device("cuda"))
    model.load_state_dict(sd)
    opt = torch.optim.Adam(model.parameters(), lr=0.001)
    opt.load_state_dict(torch.load(os.path.join(ck, "latest_optim.pth")))      # the reference's optimiser reads the file
    assert int(opt.state_dict()["state"][0]["step"]) == 36


def test_rollout_dataset_matches_reference(root, tmp_path):
    """evaluation.rollout / rollout_dataset (all pushes of all validation episodes in one device batch) against the files the
    reference's rollout_dataset wrote for the same data set, weights and numpy seed (tests/golden/eval_dataset.npz)."""
    import adaptigraph_b200 as agx
    from adaptigraph_b200 import evaluation, synthetic
    from agx_helpers import golden_weights
    E = load_npz("eval_dataset.npz")
    dc, mc = dataset_configs(root)
    model_config = synthetic.configs("rope")[0]
    ck = os.path.join(str(tmp_path), "log", "rope", "checkpoints")
    os.makedirs(ck)
    torch.save(golden_weights(), os.path.join(ck, "latest.pth"))
    config = {"dataset_config": dc, "material_config": mc, "model_config": model_config,
              "train_config": {"out_dir": os.path.join(str(tmp_path), "log"), "random_seed": 5},
              "rollout_config": {"out_dir": os.path.join(str(tmp_path), "rollout")}}
    res = evaluation.rollout(config, "latest")
    save_dir = os.path.join(str(tmp_path), "rollout", "rollout-rope-model_latest")
    got = np.loadtxt(os.path.join(save_dir, "error_short.txt"))
    assert got.shape == E["error_short"].shape
    assert np.abs(got - E["error_short"]).max() <= 1e-4, np.abs(got - E["error_short"]).max()
    assert np.array_equal(got, res["step_error"]) or np.allclose(got, res["step_error"], rtol=0, atol=1e-12)
    files = [k for k in E if k != "error_short"]
    assert len(files) == len(res["errors"]) == 4
    for k in files:
        mine = np.atleast_1d(np.loadtxt(os.path.join(save_dir, k.split("/")[0], "short", k.split("/")[1] + ".txt")))
        assert mine.shape == E[k].shape, (k, mine.shape, E[k].shape)          # same number of steps: same frame schedule
        assert np.abs(mine - E[k]).max() <= 1e-4, (k, np.abs(mine - E[k]).max())
    assert np.allclose(res["median"], np.median(E["error_short"], axis=1), atol=1e-4)


def test_construct_graph_dense_matches_sparse_start(root):
    """construct_graph(dense=True) carries the reference's padded Rr / Rs; the sparse start (relations built inside
    rollout_episodes) gives the same first-step errors."""
    from adaptigraph_b200 import evaluation, synthetic
    from adaptigraph_b200.dataset import load_dataset, load_positions
    import adaptigraph_b200 as agx
    from agx_helpers import golden_weights
    dc, mc = dataset_configs(root)
    pairs, phys = load_dataset(dc, mc, "valid")
    eef, obj = load_positions(dc)
    ep = int(pairs[0, 0])
    pe = pairs[pairs[:, 0] == ep][:, 1:]
    model = agx.DynamicsPredictor(synthetic.configs("rope")[0], mc, dc, torch.device("cuda"))
    model.load_state_dict(golden_weights())
    model.to("cuda").eval()
    outs = []
    for dense in (False, True):
        np.random.seed(9)
        g, idx = evaluation.construct_graph(dc, mc, eef[ep], obj[ep], dc["n_his"], pe[0], phys[ep], dense=dense)
        assert ("Rr" in g) == dense
        if dense:
            assert g["Rr"].shape == (dc["datasets"][0]["max_nR"], g["state"].shape[1])
        outs.append(evaluation.rollout_from_start_graph(g, idx, dc, mc, model, "cuda", eef[ep], obj[ep], pe[0][3], pe[0][4],
                                                        evaluation.get_next_pair_or_break_episode_pushes, pe))
    assert len(outs[0]) == len(outs[1]) > 3 and np.allclose(outs[0], outs[1], rtol=0, atol=1e-6)
