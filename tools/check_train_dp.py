"""Data-parallel check of the Trainer (run under torchrun on >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_train_dp.py
Every rank trains on its contiguous shard of one batch (one NCCL all-reduce of the flat gradient bucket per step); rank 0 also
trains a second model on the whole batch alone.  The two must agree (MSE is a mean over equal shards)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import adaptigraph_b200 as agx  # noqa: E402
from adaptigraph_b200 import synthetic as syn  # noqa: E402
from adaptigraph_b200.shard import shard_slice  # noqa: E402
from adaptigraph_b200.train import Trainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
B = 4 * world
w = syn.make_workload("rope", 60, B, seed=11).to("cuda")
n_p = w.p_instance.shape[1]


def batch(ww):
    d = ww.graph_dict()
    cur = ww.state[:, -1]
    d["state_future"] = torch.stack([cur[:, :n_p] + 0.01 * (i + 1) for i in range(3)], 1)
    d["eef_future"] = torch.stack([cur, cur], 1)
    d["action_future"] = torch.stack([ww.action, ww.action], 1)
    el = agx.build_edges(cur, ww.adj_thresh, ww.state_mask, ww.eef_mask, ww.topk, ww.connect_tools_all).check()
    return d, el


def model():
    torch.manual_seed(0)
    return agx.DynamicsPredictor(*syn.configs("rope", 2), "cuda").cuda().train()


m = model()
tr = Trainer(m, n_future=3)
d, el = batch(w.take(shard_slice(B, world, rank)))
for _ in range(3):
    loss = tr.step(d, el)
torch.cuda.synchronize()
if rank == 0:
    dist_mod = tr.world
    m1 = model()
    tr1 = Trainer(m1, n_future=3)
    tr1.world = 1                      # the whole batch on one GPU, no collective
    d1, el1 = batch(w)
    for _ in range(3):
        tr1.step(d1, el1)
    err = max(float((a - b).abs().max()) for a, b in zip(m.parameters(), m1.parameters()))
    print(json.dumps({"world": world, "ranks_used": dist_mod, "max_param_diff_vs_single_gpu": err, "ok": err <= 2e-5}))
dist.barrier()
dist.destroy_process_group()
