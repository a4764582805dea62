"""Host logic of the training unroll without a GPU (train.unroll_loss = train.py:88-112): with a stand-in model, check the data flow
of the n_future steps — fixed relations, predicted particles written into the newest history frame built from eef_future, the action
replaced by action_future, MSE summed over the steps, the caller's dict left untouched."""
import torch

from adaptigraph_b200.train import unroll_loss


class _Recorder:
    def __init__(self, n_p):
        self.n_p, self.seen = n_p, []

    def __call__(self, state=None, action=None, edges=None, **kw):
        assert kw["rope_physics_param"].shape[1] == 1 and "state_future" in kw       # extra keys ride along and are ignored
        self.seen.append((state.clone(), action.clone(), edges))
        pred = state[:, -1, :self.n_p] + 0.5 * action[:, -1:, :]                      # depends on both inputs
        return pred, pred - state[:, -1, :self.n_p]


def test_unroll_follows_the_reference_data_flow():
    torch.manual_seed(0)
    B, H, n_p, n_s, F = 2, 4, 5, 2, 3
    N = n_p + n_s
    data = {"state": torch.randn(B, H, N, 3), "attrs": torch.zeros(B, N, 2), "action": torch.randn(B, N, 3), "p_instance": torch.ones(B, n_p, 1),
            "rope_physics_param": torch.zeros(B, 1), "state_future": torch.randn(B, F, n_p, 3), "eef_future": torch.randn(B, F - 1, N, 3),
            "action_future": torch.randn(B, F - 1, N, 3), "Rr": None, "Rs": None}
    before = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in data.items()}
    m, edges = _Recorder(n_p), object()
    loss = unroll_loss(m, data, F, edges=edges)
    assert len(m.seen) == F and all(e is edges for _, _, e in m.seen)                 # same relations every step
    want_loss, state, action = 0.0, before["state"], before["action"]
    for fi in range(F):
        s_seen, a_seen, _ = m.seen[fi]
        assert torch.equal(s_seen, state) and torch.equal(a_seen, action)
        pred = state[:, -1, :n_p] + 0.5 * action[:, -1:, :]
        want_loss = want_loss + torch.nn.functional.mse_loss(pred, before["state_future"][:, fi])
        if fi < F - 1:
            frame = before["eef_future"][:, fi].clone()
            frame[:, :n_p] = pred                                                     # train.py:104-105
            state = torch.cat([state[:, 1:], frame[:, None]], 1)                      # :106
            action = before["action_future"][:, fi]                                   # :108
    assert torch.allclose(loss, want_loss)
    for k, v in before.items():                                                       # the batch dict is not modified
        assert (data[k] is v) if not torch.is_tensor(v) else torch.equal(data[k], v)
