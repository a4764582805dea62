"""Per-parameter gradient error of one forward + backward against torch autograd over the oracle's dense restatement.
    python tests/bench/grad_check.py [material n_p B pstep]      (AGX_TRAIN_PRECISION=fp32 selects the FFMA training layers)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import adaptigraph_b200 as agx  # noqa: E402
from adaptigraph_b200 import synthetic as syn  # noqa: E402
from oracle import dynamics_oracle as orc  # noqa: E402

material, n_p, B, pstep = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else ("granular", 150, 3, 2)
w = syn.make_workload(material, n_p, B, seed=77, n_pad=6)
torch.manual_seed(3)
m = agx.DynamicsPredictor(*syn.configs(material, pstep), "cuda").cuda().train()
wd = w.to("cuda")
el = agx.build_edges(wd.state[:, -1], w.adj_thresh, wd.state_mask, wd.eef_mask, w.topk, w.connect_tools_all).check()
st = wd.state.clone().requires_grad_(True)
d = wd.graph_dict()
d["state"] = st
pos, motion = m(**d, edges=el)
tgt = torch.randn(pos.shape, generator=torch.Generator().manual_seed(5)).cuda() * 0.05
loss = torch.nn.functional.mse_loss(pos, wd.state[:, -1, :pos.shape[1]] + tgt) + 0.1 * motion.square().mean()
loss.backward()
p = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
Rr, Rs = orc.edges_dense_batch(w.state[:, -1], w.adj_thresh, w.state_mask, w.eef_mask, w.topk, w.connect_tools_all)
st_c = w.state.clone().requires_grad_(True)
pos_c, motion_c = orc.forward_dense(p, pstep, st_c, w.attrs, Rr, Rs, w.p_instance, w.action, w.physics_param)
loss_c = torch.nn.functional.mse_loss(pos_c, w.state[:, -1, :pos_c.shape[1]] + tgt.cpu()) + 0.1 * motion_c.square().mean()
loss_c.backward()
print("precision", os.environ.get("AGX_TRAIN_PRECISION", "tc"), "loss", loss.item(), loss_c.item(), "fwd max diff", (pos.detach().cpu() - pos_c.detach()).abs().max().item())
print(f"{'state':45s} rel err {((st.grad.cpu() - st_c.grad).abs().max() / st_c.grad.abs().max()).item():.3e}  max |g| {st_c.grad.abs().max().item():.3e}")
for k, v in m.named_parameters():
    ref = p[k].grad
    print(f"{k:45s} rel err {((v.grad.cpu() - ref).abs().max() / ref.abs().max()).item():.3e}  max |g| {ref.abs().max().item():.3e}")
