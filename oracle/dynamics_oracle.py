"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU (torch fp32) restatement of the AdaptiGraph particle-graph dynamics hot
path, written to be read next to the reference:

* dense relation builder, batched   src/dynamics/dataset/graph.py:91-156   -> `edges_dense_batch`
* dense relation builder, single    src/dynamics/dataset/graph.py:38-89    -> `edges_dense_single`
* pad / truncate of relation rows   src/dynamics/utils.py:37-46, 127-137   -> `pad_rows`, `truncate_rows`
* model forward                     src/dynamics/gnn/model.py:129-313      -> `forward_dense`
  (blocks: Encoder :4-21, Propagator :23-41, ParticlePredictor :43-60)
* rollout step                      src/planning/forward_dynamics.py:156-197 -> `rollout_dense`
* training unroll                   src/dynamics/train/train.py:90-108     -> `unroll_loss_dense`

plus a sparse (edge-list) variant of the forward with the algebraic hoist the
CUDA engine uses (`forward_sparse`), which tests prove equal to the dense one.

Parity pinning: the reference has no tests or golden vectors, so
this oracle is pinned against outputs of the reference itself, imported from
/root/reference in the build container by tests/golden/make_golden.py and
committed as tests/golden/*.npz; tests/test_oracle_golden.py checks every
function here against them (bit-exact edges, <=2e-6 max-abs floats).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

MOTION_CLAMP = 100.0  # model.py:85

PARAM_SHAPES = lambda d_in, d_rel, F: {  # noqa: E731  (names as in the reference state_dict)
    "particle_encoder.model.0": (F, d_in), "particle_encoder.model.2": (F, F), "particle_encoder.model.4": (F, F),
    "relation_encoder.model.0": (F, d_rel), "relation_encoder.model.2": (F, F), "relation_encoder.model.4": (F, F),
    "particle_propagator.linear": (F, 2 * F), "relation_propagator.linear": (F, 3 * F),
    "non_rigid_predictor.linear_0": (F, F), "non_rigid_predictor.linear_1": (F, F),
    "non_rigid_predictor.linear_2": (3, F),
}


def init_params(seed: int, F: int = 150, d_in: int = 6, d_rel: int = 17) -> Dict[str, torch.Tensor]:
    """nn.Linear-style U(-1/sqrt(fan_in), 1/sqrt(fan_in)) parameters keyed like the
    reference state_dict.  (Golden fixtures carry the reference's own init.)"""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, (o, i) in PARAM_SHAPES(d_in, d_rel, F).items():
        bound = 1.0 / (i ** 0.5)
        out[name + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        out[name + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound
    return out


# --------------------------------------------------------------------------- graph builders
def _pair_tables(pos, mask, tool_mask):
    """graph.py:100-121 (batched) / :47-66 (single) on a (B,N,3) input."""
    diff = pos[:, :, None, :] - pos[:, None, :, :]
    dis = torch.sum(diff ** 2, -1)                       # ((dx^2+dy^2)+dz^2), fp32
    valid_pair = mask[:, :, None] & mask[:, None, :]
    tool_pair = tool_mask[:, :, None] & tool_mask[:, None, :]
    dis = dis.masked_fill(~valid_pair, 1e10).masked_fill(tool_pair, 1e10)
    recv_tool_send_valid = tool_mask[:, :, None] & mask[:, None, :]   # obj_tool_mask_1
    send_tool_recv_valid = tool_mask[:, None, :] & mask[:, :, None]   # obj_tool_mask_2
    return dis, tool_pair, recv_tool_send_valid, send_tool_recv_valid


def _radius_topk(dis, thr2, topk):
    """graph.py:125-132: radius test AND membership in the row's top-k smallest."""
    adj = (dis - thr2) < 0
    k = min(dis.shape[-1], topk)
    idx = torch.topk(dis, k=k, dim=-1, largest=False)[1]
    keep = torch.zeros_like(adj)
    keep.scatter_(-1, idx, True)
    return adj & keep


def adjacency_batch(pos, adj_thresh, mask, tool_mask, topk=10, connect_tools_all=False):
    """Boolean (B,N,N) adjacency [b, receiver, sender] of graph.py:91-144."""
    B = pos.shape[0]
    if isinstance(adj_thresh, float):
        adj_thresh = torch.tensor(adj_thresh, dtype=pos.dtype).repeat(B)
    thr2 = (adj_thresh * adj_thresh)[:, None, None]
    dis, _, m1, m2 = _pair_tables(pos, mask, tool_mask)
    adj = _radius_topk(dis, thr2, topk)
    if connect_tools_all:
        probe = tool_mask[:, :, None] & ~tool_mask[:, None, :]            # graph.py:123
        on = (adj & probe).flatten(1).any(1)[:, None, None]               # graph.py:135
        adj = adj & ~m1           # :141 and :143 together clear every tool-receiver/valid-sender entry
        adj = adj & ~m2           # :144 (and the entries :142 is about to set)
        adj = adj | (m2 & on)     # :142
    return adj


def adjacency_single(pos, adj_thresh, mask, tool_mask, topk=10, connect_tools_all=False):
    """Boolean (N,N) adjacency of graph.py:38-80 (threshold squared in Python float)."""
    thr2 = adj_thresh * adj_thresh
    dis, tool_pair, m1, m2 = _pair_tables(pos[None], mask[None], tool_mask[None])
    adj = _radius_topk(dis, thr2, topk)
    if connect_tools_all:
        adj = ((adj & ~m1) | m2) & ~tool_pair                              # :78-80
    return adj[0]


def onehots_from_adjacency(adj, dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """graph.py:146-155: relation rows ordered by (b, receiver, sender), zero-padded to max_b E_b."""
    B, N, _ = adj.shape
    counts = adj.flatten(1).sum(1)
    n_rel = int(counts.max())
    Rr = torch.zeros(B, n_rel, N, dtype=dtype)
    Rs = torch.zeros(B, n_rel, N, dtype=dtype)
    for b in range(B):
        rs = adj[b].nonzero()
        e = torch.arange(rs.shape[0])
        Rr[b, e, rs[:, 0]] = 1
        Rs[b, e, rs[:, 1]] = 1
    return Rr, Rs


def edges_dense_batch(pos, adj_thresh, mask, tool_mask, topk=10, connect_tools_all=False):
    return onehots_from_adjacency(adjacency_batch(pos, adj_thresh, mask, tool_mask, topk, connect_tools_all), pos.dtype)


def edges_dense_single(pos, adj_thresh, mask, tool_mask, topk=10, connect_tools_all=False):
    Rr, Rs = onehots_from_adjacency(adjacency_single(pos, adj_thresh, mask, tool_mask, topk, connect_tools_all)[None], pos.dtype)
    return Rr[0], Rs[0]


def edge_lists_from_adjacency(adj):
    """CSR by receiver over the flattened (B*N) node index: row_ptr int32 (B*N+1), send int32 (E_tot)
    with batch-local sender ids, the layout the CUDA builder emits."""
    B, N, _ = adj.shape
    deg = adj.sum(-1).flatten()
    row_ptr = torch.zeros(B * N + 1, dtype=torch.int32)
    row_ptr[1:] = torch.cumsum(deg, 0).to(torch.int32)
    send = adj.nonzero()[:, 2].to(torch.int32)
    return row_ptr, send


def pad_rows(R, max_rows):
    """utils.py:37-46 (dim=1 case): raises when the graph has more rows than max_rows."""
    out = torch.zeros(R.shape[0], max_rows, R.shape[2], dtype=R.dtype)
    out[:, :R.shape[1]] = R
    return out


def truncate_rows(Rr, Rs):
    """utils.py:127-137."""
    n = max(int((Rr.sum(-1) > 0).sum(1).max()), int((Rs.sum(-1) > 0).sum(1).max()))
    return Rr[:, :n], Rs[:, :n]


# --------------------------------------------------------------------------- model forward
def _lin(p, name, x):
    return torch.nn.functional.linear(x, p[name + ".weight"], p[name + ".bias"])


def _encoder(p, prefix, x, relu=None, names=None):
    """model.py:4-21 — three Linear layers, ReLU after each (including the last).  `relu(name, x)`, when given, stands in for
    torch.relu at the site `names[i]` (gradient parity tests substitute the activity pattern of the engine under test)."""
    for n, i in enumerate((0, 2, 4)):
        x = _lin(p, f"{prefix}.model.{i}", x)
        x = torch.relu(x) if relu is None else relu(names[n], x)
    return x


def node_and_relation_inputs(state, attrs, p_instance, action, physics_param):
    """model.py:155-195 for the shipped configs (state_dim 0, action_dim 3, density 0)."""
    B, H, N, D = state.shape
    n_p = p_instance.shape[1]
    hist = torch.cat([state[:, 1:] - state[:, :-1], state[:, -1:]], 1)       # :155-164
    hist = hist.transpose(1, 2).reshape(B, N, H * D)                         # :165
    phys = torch.cat([physics_param[:, None, :].repeat(1, n_p, 1),
                      torch.zeros(B, N - n_p, physics_param.shape[1])], 1)   # :186-189
    p_in = torch.cat([attrs, phys, action], 2)                               # :190,195
    group = torch.cat([p_instance, torch.zeros(B, N - n_p, p_instance.shape[2])], 1)  # :235
    return hist, p_in, group


def forward_dense(p, pstep, state, attrs, Rr, Rs, p_instance, action, physics_param):
    """DynamicsPredictor.forward (model.py:129-313) with the reference's dense one-hot algebra."""
    n_p = p_instance.shape[1]
    hist, p_in, group = node_and_relation_inputs(state, attrs, p_instance, action, physics_param)
    Rr_t = Rr.transpose(1, 2).contiguous()                                   # :151
    rel_in = torch.cat([Rr.bmm(attrs), Rs.bmm(attrs),                        # :224-228
                        (Rr.bmm(group) - Rs.bmm(group)).abs().sum(2, keepdim=True),  # :236-238
                        Rr.bmm(hist) - Rs.bmm(hist)], 2)                     # :248-250
    penc = _encoder(p, "particle_encoder", p_in)                             # :268
    renc = _encoder(p, "relation_encoder", rel_in)                           # :274
    eff = penc
    for _ in range(pstep):                                                   # :278-301
        e_in = torch.cat([renc, Rr.bmm(eff), Rs.bmm(eff)], 2)
        e_out = torch.relu(_lin(p, "relation_propagator.linear", e_in))
        agg = Rr_t.bmm(e_out)
        eff = torch.relu(_lin(p, "particle_propagator.linear", torch.cat([penc, agg], 2)) + eff)
    h = torch.relu(_lin(p, "non_rigid_predictor.linear_0", eff[:, :n_p]))    # :306
    h = torch.relu(_lin(p, "non_rigid_predictor.linear_1", h))
    motion = _lin(p, "non_rigid_predictor.linear_2", h)
    pos = state[:, -1, :n_p] + motion.clamp(-MOTION_CLAMP, MOTION_CLAMP)     # :309
    return pos, motion


def forward_sparse(p, pstep, state, attrs, row_ptr, send, p_instance, action, physics_param, relu=None):
    """Same function on CSR edge lists with the propagator weights split per operand
    (W[:, :F] relation part, W[:, F:2F] receiver part, W[:, 2F:] sender part) so the
    per-edge 450->150 product becomes gathers of per-node products.  Equal to
    `forward_dense` up to fp32 summation order.

    relu: optional `relu(site, x)` replacing torch.relu at the sites h1, h2, penc (particle encoder), g1, g2, renc (relation
    encoder), e<k> (relation effects of propagation step k), P<k+1> (particle effects), u1, u2 (predictor)."""
    act = (lambda name, x: torch.relu(x)) if relu is None else relu
    B, H, N, _ = state.shape
    n_p = p_instance.shape[1]
    hist, p_in, group = node_and_relation_inputs(state, attrs, p_instance, action, physics_param)
    deg = (row_ptr[1:] - row_ptr[:-1]).long()
    recv = torch.repeat_interleave(torch.arange(B * N), deg)
    snd = send.long() + (recv // N) * N
    fl = lambda t: t.reshape(B * N, -1)  # noqa: E731
    a, g, hs = fl(attrs), fl(group), fl(hist)
    rel_in = torch.cat([a[recv], a[snd], (g[recv] - g[snd]).abs().sum(1, keepdim=True), hs[recv] - hs[snd]], 1)
    penc = _encoder(p, "particle_encoder", fl(p_in), relu, ("h1", "h2", "penc"))
    renc = _encoder(p, "relation_encoder", rel_in, relu, ("g1", "g2", "renc"))
    F = penc.shape[1]
    Wr, br = p["relation_propagator.linear.weight"], p["relation_propagator.linear.bias"]
    Wp, bp = p["particle_propagator.linear.weight"], p["particle_propagator.linear.bias"]
    c_edge = renc @ Wr[:, :F].T + br
    a_node = penc @ Wp[:, :F].T + bp
    eff = penc
    for k in range(pstep):
        q_r, q_s = eff @ Wr[:, F:2 * F].T, eff @ Wr[:, 2 * F:].T
        e_out = act(f"e{k}", c_edge + q_r[recv] + q_s[snd])
        agg = torch.zeros_like(eff).index_add_(0, recv, e_out)
        eff = act(f"P{k + 1}", a_node + agg @ Wp[:, F:].T + eff)
    eff = eff.reshape(B, N, F)[:, :n_p]
    h = act("u1", _lin(p, "non_rigid_predictor.linear_0", eff))
    h = act("u2", _lin(p, "non_rigid_predictor.linear_1", h))
    motion = _lin(p, "non_rigid_predictor.linear_2", h)
    return state[:, -1, :n_p] + motion.clamp(-MOTION_CLAMP, MOTION_CLAMP), motion


# --------------------------------------------------------------------------- rollout / unroll
def rollout_dense(p, pstep, state, attrs, p_instance, action, physics_param, state_mask, eef_mask,
                  adj_thresh, topk, connect_tools_all, n_steps, y_mode="min", gripper_raise=0.0,
                  frozen_edges=None):
    """forward_dynamics.py:156-197 (and :351-393 with y_mode='masked_mean').

    Per step: forward on the current graph; tool particles move by their action
    delta and take y = min over predicted object y (:163-168); relations are rebuilt
    on [pred ; tools] (:171); history shifts by one frame (:176).  Returns
    (B, n_steps, n_p, 3) predicted positions.  `frozen_edges` = list of (Rr, Rs) per
    step to replay a recorded edge sequence instead of rebuilding.
    """
    B, H, N, _ = state.shape
    n_p = p_instance.shape[1]
    build = lambda pos: edges_dense_batch(pos, adj_thresh, state_mask, eef_mask, topk, connect_tools_all)  # noqa: E731
    Rr, Rs = frozen_edges[0] if frozen_edges else build(state[:, -1])
    out, edges = [], []
    for t in range(n_steps):
        edges.append((Rr, Rs))
        pred, _ = forward_dense(p, pstep, state, attrs, Rr, Rs, p_instance, action, physics_param)
        out.append(pred)
        if y_mode == "min":
            y = pred[:, :, 1].min(1).values                                   # :163
        else:
            m = state_mask[:, :n_p].to(pred.dtype)
            y = (pred[:, :, 1] * m).sum(1) / m.sum(1)                         # :359
        tool = state[:, -1, n_p:] + action[:, n_p:]                           # :164
        tool[:, :, 1] = y[:, None] + gripper_raise                            # :166-168
        cur = torch.cat([pred, tool], 1)                                      # :170
        if t + 1 < n_steps:
            Rr, Rs = frozen_edges[t + 1] if frozen_edges else build(cur)      # :171
        state = torch.cat([state[:, 1:], cur[:, None]], 1)                    # :176
    return torch.stack(out, 1), edges


def unroll_loss_dense(p, pstep, state, attrs, Rr, Rs, p_instance, action, physics_param,
                      state_future, eef_future, action_future):
    """train.py:90-108: n_future forwards on fixed relations, MSE each, summed; the next
    state's last frame is eef_future with the predicted object rows written in."""
    n_future = state_future.shape[1]
    loss = 0.0
    for fi in range(n_future):
        pred, _ = forward_dense(p, pstep, state, attrs, Rr, Rs, p_instance, action, physics_param)
        loss = loss + torch.nn.functional.mse_loss(pred, state_future[:, fi])
        if fi < n_future - 1:
            nxt = eef_future[:, fi].clone()
            nxt = torch.cat([pred, nxt[:, pred.shape[1]:]], 1)
            state = torch.cat([state[:, 1:], nxt[:, None]], 1)
            action = action_future[:, fi]
    return loss
