"""BASELINE configs[1]: rope ~300 particles, batch 32, 4 propagation steps, forward + backward on one B200,
next to the oracle's dense CPU restatement with torch autograd on the host cores (a reported baseline).
Prints one JSON line (particle-steps/s = B * n_p / time per fwd+bwd)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import adaptigraph_b200 as agx  # noqa: E402
from adaptigraph_b200 import synthetic as syn  # noqa: E402
from oracle import dynamics_oracle as orc  # noqa: E402

c = syn.BASELINE_CONFIGS[2]
w_host = syn.baseline_workload(2)
torch.manual_seed(0)
m = agx.DynamicsPredictor(*syn.configs(c["material"], c["pstep"]), "cuda").cuda().train()
w = w_host.to("cuda")
el = agx.build_edges(w.state[:, -1], w.adj_thresh, w.state_mask, w.eef_mask, w.topk, w.connect_tools_all).check()
E = int(el.row_ptr[-1])


def step():
    m.zero_grad(set_to_none=False)
    st = w.state.clone().requires_grad_(True)
    d = w.graph_dict()
    d["state"] = st
    pos, _ = m(**d, edges=el)
    loss = pos.square().mean()
    loss.backward()
    return loss, st.grad


for _ in range(3):
    step()
torch.cuda.synchronize()
K = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    loss, gs = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K

# CPU: dense oracle + autograd (bounded: 2 passes)
torch.set_num_threads(os.cpu_count())
p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
Rr, Rs = orc.edges_dense_batch(w_host.state[:, -1], w_host.adj_thresh, w_host.state_mask, w_host.eef_mask, w_host.topk, False)


def cpu_step():
    for v in p.values():
        v.grad = None
    st = w_host.state.clone().requires_grad_(True)
    pos, _ = orc.forward_dense(p, c["pstep"], st, w_host.attrs, Rr, Rs, w_host.p_instance, w_host.action, w_host.physics_param)
    l = pos.square().mean()
    l.backward()
    return l, st.grad


cpu_step()
t0 = time.perf_counter()
for _ in range(2):
    l_cpu, gs_cpu = cpu_step()
cpu_ms = (time.perf_counter() - t0) / 2 * 1e3
gerr = max(float((dict(m.named_parameters())[k].grad.cpu() - v.grad).abs().max() / max(1e-12, float(v.grad.abs().max()))) for k, v in p.items())
# the reference's whole training iteration (train.py:84-115: zero_grad, n_future = 3 unroll, backward, Adam) through the Trainer,
# eager and as one replayed CUDA graph
from adaptigraph_b200.train import Trainer  # noqa: E402


def trainer_ms(cuda_graph):
    torch.manual_seed(0)
    mm = agx.DynamicsPredictor(*syn.configs(c["material"], c["pstep"]), "cuda").cuda().train()
    tr = Trainer(mm, n_future=3, cuda_graph=cuda_graph)
    d = w.graph_dict()
    n_p = w.p_instance.shape[1]
    cur = w.state[:, -1]
    d["state_future"] = torch.stack([cur[:, :n_p] + 0.01 * (i + 1) for i in range(3)], 1)
    d["eef_future"] = torch.stack([cur, cur], 1)
    d["action_future"] = torch.stack([w.action, w.action], 1)
    for _ in range(4):
        tr.step(d, el)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        last = tr.step(d, el)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / K, float(last)


tr_eager_ms, tr_eager_loss = trainer_ms(False)
tr_graph_ms, tr_graph_loss = trainer_ms(True)

print(json.dumps({
    "trainer_n_future3_ms_per_iter_eager": tr_eager_ms, "trainer_n_future3_ms_per_iter_cuda_graph": tr_graph_ms,
    "trainer_loss_after_24_iters": [tr_eager_loss, tr_graph_loss],
    "workload": "rope 300 particles, batch 32, pstep 4, forward+backward (BASELINE configs[1])", "relations": E,
    "gpu_ms_per_step": ms, "gpu_particle_steps_per_s": c["B"] * c["n_p"] / (ms * 1e-3),
    "cpu_ms_per_step": cpu_ms, "cpu_particle_steps_per_s": c["B"] * c["n_p"] / (cpu_ms * 1e-3), "cpu_cores": os.cpu_count(),
    "loss_abs_diff": abs(float(loss) - float(l_cpu)), "state_grad_max_abs_diff": float((gs.cpu() - gs_cpu).abs().max()),
    "param_grad_max_rel_diff": gerr}))
