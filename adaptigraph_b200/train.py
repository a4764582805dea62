"""Training-loop glue for the dynamics model, mirroring src/dynamics/train/train.py:61-130.

    unroll_loss(model, data, n_future, edges=None)   the inner loop of train.py:88-112 (n_future forwards on fixed relations,
                                                     MSE each step, predicted particles written back into the history)
    Trainer(model, lr=1e-3, ...)                     optimizer.zero_grad / loss.backward / optimizer.step of train.py:84-115 with
                                                     * every parameter and gradient re-homed as a view of ONE flat fp32 bucket,
                                                     * Adam as one fused kernel over that bucket (`agx_adam_step`),
                                                     * data parallelism = one all-reduce of the flat gradient bucket (NCCL),
                                                     * optionally the whole step captured in a CUDA graph and replayed.

The optimizer state round-trips through torch.optim.Adam's state_dict format, so `latest_optim.pth` files written by the
reference (train.py:121) load here and vice versa.  Nothing here computes on the CPU.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import ops
from .graph import EdgeList, edges_from_onehots

_IGNORED_BY_FORWARD = ("state_future", "eef_future", "action_future")


def unroll_loss(model, data: Dict[str, torch.Tensor], n_future: int, edges: Optional[EdgeList] = None) -> torch.Tensor:
    """Sum over n_future autoregressive steps of MSE(pred_state, state_future[:, fi]) — train.py:88-112 verbatim in behaviour:
    relations stay fixed, the prediction replaces the object rows of the newest history frame built from eef_future, the
    action becomes action_future[:, fi].  `data` is not modified."""
    data = dict(data)
    own = data.pop("edges", None)                               # a batch of dataset.DynDataset carries its EdgeList
    future_state, future_eef, future_action = data["state_future"], data["eef_future"], data["action_future"]
    if edges is None:
        edges = own if own is not None else edges_from_onehots(data["Rr"], data["Rs"])     # once, not once per forward
    loss_sum = 0
    for fi in range(n_future):
        gt_state = future_state[:, fi]
        pred_state, _ = model(**data, edges=edges)
        pred_state_p = pred_state[:, :gt_state.shape[1], :3]
        loss_sum = loss_sum + torch.nn.functional.mse_loss(pred_state_p, gt_state)
        if fi < n_future - 1:
            next_state = future_eef[:, fi].clone().unsqueeze(1)                     # (B, 1, n_p+n_s, 3)
            next_state[:, -1, :pred_state_p.shape[1]] = pred_state_p
            data["state"] = torch.cat([data["state"][:, 1:], next_state], dim=1)    # (B, n_his, n_p+n_s, 3)
            data["action"] = future_action[:, fi]
    return loss_sum


class FlatParameters:
    """Re-homes the parameters of `model` (and their .grad) as views into two flat fp32 device buffers, in
    model.parameters() order — the order torch.optim.Adam indexes its state by."""

    def __init__(self, model: torch.nn.Module):
        self.params = [p for p in model.parameters()]
        if not self.params or not all(p.is_cuda and p.dtype == torch.float32 for p in self.params):
            raise RuntimeError("FlatParameters needs float32 CUDA parameters (no CPU path)")
        dev = self.params[0].device
        self.sizes = [p.numel() for p in self.params]
        n = sum(self.sizes)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p, k in zip(self.params, self.sizes):
                self.flat[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat[off:off + k].view(p.shape)
                off += k
        self.attach_grads()

    def attach_grads(self) -> None:
        """(Re-)points every parameter's .grad at its slice of the flat gradient bucket.  `zero_grad(set_to_none=True)` — the
        default of torch's optimizers and of nn.Module.zero_grad — detaches them; a backward would then accumulate into fresh
        tensors the fused Adam never sees.  Called before every step, so that usage is harmless."""
        off = 0
        for p, k in zip(self.params, self.sizes):
            g = p.grad
            if g is None or g.data_ptr() != self.grad.data_ptr() + 4 * off or g.shape != p.shape:
                p.grad = self.grad[off:off + k].view(p.shape)
            off += k

    def views(self, flat: torch.Tensor):
        out, off = [], 0
        for p, k in zip(self.params, self.sizes):
            out.append(flat[off:off + k].view(p.shape))
            off += k
        return out


class Trainer:
    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, n_future: int = 3,
                 group: Optional[dist.ProcessGroup] = None, cuda_graph: bool = False):
        self.model, self.n_future = model, n_future
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.bucket = FlatParameters(model)
        dev = self.bucket.flat.device
        self.exp_avg = torch.zeros_like(self.bucket.flat)
        self.exp_avg_sq = torch.zeros_like(self.bucket.flat)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.cuda_graph = cuda_graph
        self._graph = None
        self._static: Optional[Dict[str, torch.Tensor]] = None
        self._static_edges: Optional[EdgeList] = None
        self._static_loss: Optional[torch.Tensor] = None
        model._packed = None     # the parameters moved

    # ------------------------------------------------------------------ one optimisation step
    def _step_impl(self, data, edges) -> torch.Tensor:
        self.bucket.attach_grads()                                          # survives a caller's zero_grad(set_to_none=True)
        self.bucket.grad.zero_()                                            # optimizer.zero_grad()          train.py:85
        self.model._packed = None                                           # parameters change under the kernels' feet
        loss = unroll_loss(self.model, data, self.n_future, edges)          #                                train.py:88-112
        loss.backward()                                                     # accumulates into the flat bucket's views
        if self.world > 1:                                                  # data parallel: ONE collective per step
            dist.all_reduce(self.bucket.grad, op=dist.ReduceOp.SUM, group=self.group)
        ops.adam_step(self.bucket.flat, self.bucket.grad, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr,
                      self.betas[0], self.betas[1], self.eps, 1.0 / self.world)   # optimizer.step()          train.py:115
        # the fused Adam writes the flat bucket through raw pointers (no _version bump): the packed blob the kernels read is
        # stale from here on -- drop it so that the next forward / rollout / evaluation repacks the updated weights
        self.model._packed = None
        return loss.detach()

    def step(self, data: Dict[str, torch.Tensor], edges: Optional[EdgeList] = None) -> torch.Tensor:
        """One training iteration on a batch in the reference's dict format (dataset.py:212-252 keys).  Returns the summed
        loss as a device scalar (no host synchronisation)."""
        self.model.train()
        if edges is None:
            edges = data["edges"] if data.get("edges") is not None else edges_from_onehots(data["Rr"], data["Rs"])
        if not self.cuda_graph:
            return self._step_impl(data, edges)
        return self._step_graphed(data, edges)

    # ------------------------------------------------------------------ CUDA-graph replay
    def _step_graphed(self, data, edges) -> torch.Tensor:
        tens = {k: v for k, v in data.items() if torch.is_tensor(v) and k not in ("Rr", "Rs")}    # an "edges" entry is not a tensor
        if self._graph is None:
            self._static = {k: v.clone() for k, v in tens.items()}
            self._static_edges = EdgeList(edges.row_ptr.clone(), edges.send.clone(), edges.recv.clone(), edges.n_edges.clone(),
                                          edges.status.clone(), edges.B, edges.N)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                                   # warm-up outside capture (attribute setup, allocator)
                snap = [t.clone() for t in (self.bucket.flat, self.exp_avg, self.exp_avg_sq, self.step_count)]
                for _ in range(2):
                    self._step_impl(self._static, self._static_edges)
                for t, s in zip((self.bucket.flat, self.exp_avg, self.exp_avg_sq, self.step_count), snap):
                    t.copy_(s)                                              # the warm-up steps must not train
            torch.cuda.current_stream().wait_stream(side)
            # The warm-up filled the static EdgeList's sender-list cache (autograd.sender_lists).  Were it still valid inside the
            # capture, the sort / scan that builds send_ptr / send_perm would not be recorded and every replay would pair the NEW
            # relations with the FIRST batch's sender lists (silently wrong dQs, d_state and upstream gradients).  Dropping it
            # forces the rebuild to be part of the graph, writing the same static buffers on every replay.
            self._static_edges._sender_cache = None
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):                             # records, does not execute
                self._static_loss = self._step_impl(self._static, self._static_edges)
        for k, v in tens.items():
            if self._static[k].shape != v.shape:
                raise RuntimeError(f"cuda_graph=True needs constant shapes; {k} changed {tuple(self._static[k].shape)} -> {tuple(v.shape)}")
            self._static[k].copy_(v)
        se = self._static_edges
        if edges.send.numel() != se.send.numel() or edges.row_ptr.numel() != se.row_ptr.numel():
            raise RuntimeError("cuda_graph=True needs a constant relation capacity (pad to max_nR as the reference's DataLoader does)")
        se.row_ptr.copy_(edges.row_ptr); se.send.copy_(edges.send); se.recv.copy_(edges.recv)
        self._graph.replay()
        self.model._packed = None                                           # the replayed Adam changed the weights (see _step_impl)
        return self._static_loss.clone()

    # ------------------------------------------------------------------ torch.optim.Adam-compatible state (train.py:121)
    def optimizer_state_dict(self) -> dict:
        m, v = self.bucket.views(self.exp_avg), self.bucket.views(self.exp_avg_sq)
        step = self.step_count.to(torch.float32).reshape(()).cpu()
        state = {i: {"step": step.clone(), "exp_avg": m[i].clone(), "exp_avg_sq": v[i].clone()} for i in range(len(m))} \
            if int(step) > 0 else {}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "params": list(range(len(m)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd: dict) -> None:
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(g["lr"]), (float(g["betas"][0]), float(g["betas"][1])), float(g["eps"])
        m, v = self.bucket.views(self.exp_avg), self.bucket.views(self.exp_avg_sq)
        with torch.no_grad():
            steps = set()
            for i in range(len(m)):
                st = sd["state"].get(i)
                if st is None:
                    m[i].zero_(); v[i].zero_()
                    continue
                m[i].copy_(st["exp_avg"]); v[i].copy_(st["exp_avg_sq"])
                steps.add(int(st["step"]))
            if len(steps) > 1:
                raise ValueError("per-parameter step counts differ; the fused optimiser keeps one")
            self.step_count.fill_(steps.pop() if steps else 0)
        self._graph = None     # the hyper-parameters are baked into a captured step: capture again on the next call


# ---------------------------------------------------------------------- the epoch loop of train.py:19-147
def train(config: dict, cuda_graph: bool = False, log=print) -> Dict[str, list]:
    """`train(config)` of src/dynamics/train/train.py with the reference's YAML dict: the same phases, iteration counts, seeds,
    sample order, loss and checkpoint files (`checkpoints/model_<epoch>.pth`, `latest.pth`, `latest_optim.pth` in torch.optim.Adam's
    format), on `dataset.DynDataset` batches (sparse relations built on the device) and `Trainer.step`.  The loss-curve PNG of
    train.py:125-141 is not drawn; the two curves are returned instead ({"train": [...], "valid": [...]}, one mean per epoch)."""
    import os
    import random
    import time

    import numpy as np

    from .dataset import DynDataset, make_loader
    from .model import DynamicsPredictor
    dataset_config, train_config = config["dataset_config"], config["train_config"]
    model_config, material_config = config["model_config"], config["material_config"]
    out_dir = os.path.join(train_config["out_dir"], dataset_config["data_name"])
    os.makedirs(os.path.join(out_dir, "checkpoints"), exist_ok=True)
    seed = train_config["random_seed"]                                       # set_seed (utils.py:108-114)
    torch.manual_seed(seed); torch.cuda.manual_seed_all(seed); np.random.seed(seed); random.seed(seed)
    device = torch.device("cuda", torch.cuda.current_device())
    n_future, phases = dataset_config["n_future"], train_config["phases"]
    datasets = {phase: DynDataset(dataset_config, material_config, phase, device=device) for phase in phases}
    loaders = {phase: make_loader(datasets[phase], train_config["batch_size"], shuffle=(phase == "train")) for phase in phases}

    def forever(loader):                                                     # dataloader_wrapper (utils.py:116-122)
        while True:
            for data in loader:
                yield data
    streams = {phase: forever(loaders[phase]) for phase in phases}
    model = DynamicsPredictor(model_config, material_config, dataset_config, device)
    model.to(device)
    trainer = Trainer(model, lr=0.001, n_future=n_future, cuda_graph=cuda_graph)
    curves: Dict[str, list] = {"train": [], "valid": []}
    for epoch in range(train_config["n_epochs"]):
        t0 = time.time()
        for phase in phases:
            losses = []
            n_iters = train_config["n_iters_per_epoch"][phase] if train_config["n_iters_per_epoch"][phase] != -1 else len(datasets[phase])
            for i in range(n_iters):
                data = next(streams[phase])
                if phase == "train":
                    loss = trainer.step(data)
                    if i % train_config["log_interval"] == 0:
                        log(f"Epoch {epoch}, iter {i}, loss {loss.item()}")
                        losses.append(loss.item())
                else:
                    model.eval()
                    with torch.no_grad():
                        losses.append(unroll_loss(model, data, n_future).item())
            if phase == "valid":
                log(f"\nEpoch {epoch}, valid loss {np.mean(losses)}")
            curves[phase].append(float(np.mean(losses)) if losses else float("nan"))
        if ((epoch + 1) < 100 and (epoch + 1) % 10 == 0) or (epoch + 1) % 100 == 0:
            torch.save(model.state_dict(), os.path.join(out_dir, "checkpoints", f"model_{epoch + 1}.pth"))
        torch.save(model.state_dict(), os.path.join(out_dir, "checkpoints", "latest.pth"))
        torch.save(trainer.optimizer_state_dict(), os.path.join(out_dir, "checkpoints", "latest_optim.pth"))
        log(f"Epoch {epoch} time: {time.time() - t0}\n")
    return curves
