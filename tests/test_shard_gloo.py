"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous batch shards, ordered gather,
max-over-ranks timing, and bench.py's reference arm under torchrun (rank 0 prints, the others exit 0)."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_slices_partition_the_batch():
    from adaptigraph_b200.shard import shard_slice
    for B in (1, 7, 128, 129):
        for world in (1, 2, 4, 8):
            sl = [shard_slice(B, world, r) for r in range(world)]
            assert sl[0].start == 0 and sl[-1].stop == B
            assert all(a.stop == b.start for a, b in zip(sl, sl[1:]))
            sizes = [s.stop - s.start for s in sl]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adaptigraph_b200 import synthetic as syn
    from adaptigraph_b200.shard import gather_batch, max_over_ranks, shard_graph_dict, shard_slice
    from oracle import dynamics_oracle as orc
    w = syn.make_workload("rope", 30, B, seed=3)
    p = orc.init_params(1)
    g = shard_graph_dict(w.graph_dict(), world, rank)
    sl = shard_slice(B, world, rank)
    assert g["state"].shape[0] == sl.stop - sl.start and torch.equal(g["state"], w.state[sl])
    # each rank runs its shard (here: the CPU oracle stands in for the engine), results gathered in batch order
    Rr, Rs = orc.edges_dense_batch(g["state"][:, -1], w.adj_thresh, g["state_mask"], g["eef_mask"], w.topk, False)
    pos, _ = orc.forward_dense(p, 2, g["state"], g["attrs"], Rr, Rs, g["p_instance"], g["action"], g["rope_physics_param"])
    full = gather_batch(pos, B)
    t = max_over_ranks(float(rank + 1), "cpu")
    if rank == 0:
        Rr, Rs = orc.edges_dense_batch(w.state[:, -1], w.adj_thresh, w.state_mask, w.eef_mask, w.topk, False)
        ref, _ = orc.forward_dense(p, 2, w.state, w.attrs, Rr, Rs, w.p_instance, w.action, w.physics_param)
        torch.save({"err": float((full - ref).abs().max()), "t": t}, out)
    dist.destroy_process_group()


def test_sharded_forward_equals_whole_batch_gloo(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, 29531, 5, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["err"] <= 1e-6      # shard == whole (uneven split 3 + 2)
    assert res["t"] == 2.0         # max over ranks


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adaptigraph_b200 import synthetic as syn
    from adaptigraph_b200.shard import allreduce_gradients, shard_slice
    from oracle import dynamics_oracle as orc
    B = 4
    w = syn.make_workload("rope", 20, B, seed=9)
    p = {k: v.clone().requires_grad_(True) for k, v in orc.init_params(2).items()}

    def loss_of(sl):
        ww = w.take(sl)
        Rr, Rs = orc.edges_dense_batch(ww.state[:, -1], w.adj_thresh, ww.state_mask, ww.eef_mask, w.topk, False)
        pos, _ = orc.forward_dense(p, 2, ww.state, ww.attrs, Rr, Rs, ww.p_instance, ww.action, ww.physics_param)
        return pos.square().mean()

    loss_of(shard_slice(B, world, rank)).backward()              # every rank: its shard (the oracle stands in for the engine)

    class _P:                                                     # minimal parameter-like view
        def __init__(self, t): self.grad = t.grad
    allreduce_gradients([_P(t) for t in p.values()])
    if rank == 0:
        q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
        ww = w
        Rr, Rs = orc.edges_dense_batch(ww.state[:, -1], w.adj_thresh, ww.state_mask, ww.eef_mask, w.topk, False)
        pos, _ = orc.forward_dense(q, 2, ww.state, ww.attrs, Rr, Rs, ww.p_instance, ww.action, ww.physics_param)
        pos.square().mean().backward()                            # whole batch on one rank
        err = max(float((p[k].grad - q[k].grad).abs().max()) for k in p)
        torch.save({"err": err}, out)
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_equals_whole_batch_gloo(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(2, 29533, out), nprocs=2, join=True)
    assert torch.load(out)["err"] <= 1e-6      # mean over equal shards == whole-batch mean


def test_bench_reference_arm_under_torchrun_prints_one_line():
    env = dict(os.environ, AGX_BENCH_CPU_B="1", OMP_NUM_THREADS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29532", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
           "--rollout-steps", "1"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["value"] > 0 and j["cpu_baseline"]["kind"] in ("reference", "port")
    assert j["cpu_processes"] == 1 and j["scales_with_gpus"] is False
    assert j["e2e"]["h2d_bytes_per_step"] == 0
