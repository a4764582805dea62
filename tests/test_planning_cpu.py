"""Host logic of the MPC drivers without a GPU: samples are sorted by their repeat count and leave the batch once captured
(planning._rollout_captured).  A stand-in model whose `rollout` is a closed form (every step adds the tool delta of the sample to its
particles) lets the bookkeeping — ordering, segment lengths, history hand-over, capture indices, zero-repeat samples — be checked
exactly on CPU tensors."""
import torch

from adaptigraph_b200 import planning


class _LinearModel:
    def __init__(self):
        self.calls = []

    def rollout(self, state, attrs, action, p_instance, physics_param, state_mask, eef_mask, adj_thresh, topk, connect_tools_all, n_steps,
                max_nR, y_mode="min", gripper_raise=0.0, check=True):
        B, H, N, _ = state.shape
        n_obj = p_instance.shape[1]
        self.calls.append((B, n_steps))
        step = action[:, n_obj:].mean(1, keepdim=True)                        # (B, 1, 3): the sample's tool delta moves every particle
        cur = state[:, -1].clone()
        seq = []
        hist = state.clone()
        for _ in range(n_steps):
            cur = cur + step
            hist = torch.cat([hist[:, 1:], cur[:, None]], 1)
            seq.append(cur[:, :n_obj])
        return {"state_seqs": torch.stack(seq, 1), "n_edges": torch.zeros(n_steps, B, dtype=torch.int32), "state": hist,
                "status": torch.zeros(1, dtype=torch.int32)}


def test_samples_leave_the_batch_at_their_own_repeat_count():
    torch.manual_seed(0)
    bsz, n_obj, n_eef, H = 9, 5, 2, 4
    N = n_obj + n_eef
    pos0 = torch.randn(bsz, N, 3)
    states = pos0[:, None].repeat(1, H, 1, 1)
    delta = torch.zeros(bsz, N, 3)
    delta[:, n_obj:] = torch.randn(bsz, 1, 3)
    repeat = torch.tensor([3, 0, 5, 1, 3, 5, 2, 0, 1], dtype=torch.int32)
    m = _LinearModel()
    pred = planning._rollout_captured(m, states, torch.zeros(bsz, N, 2), delta, torch.ones(bsz, n_obj, 1), torch.zeros(bsz, 1),
                                      torch.ones(bsz, N, dtype=torch.bool), torch.zeros(bsz, N, dtype=torch.bool), 0.5, 10, False, repeat,
                                      max_nR=100, y_mode="min", raise_=0.0)
    want = pos0[:, :n_obj] + repeat[:, None, None].float() * delta[:, n_obj:].mean(1, keepdim=True)
    want[repeat == 0] = 0.0                                                   # never captured: zeros (forward_dynamics.py:150)
    assert torch.allclose(pred, want, atol=1e-6)
    # segments: distinct counts 1, 2, 3, 5 -> batches of 7, 5, 4, 2 samples advanced by 1, 1, 1, 2 steps = sum(repeat) sample-steps
    assert m.calls == [(7, 1), (5, 1), (4, 1), (2, 2)]
    assert sum(b * t for b, t in m.calls) == int(repeat.sum())


def test_all_zero_repeats_do_no_work():
    m = _LinearModel()
    pred = planning._rollout_captured(m, torch.zeros(3, 4, 6, 3), torch.zeros(3, 6, 2), torch.zeros(3, 6, 3), torch.ones(3, 5, 1),
                                      torch.zeros(3, 1), torch.ones(3, 6, dtype=torch.bool), torch.zeros(3, 6, dtype=torch.bool), 0.5, 10,
                                      False, torch.zeros(3, dtype=torch.int32), max_nR=10, y_mode="min", raise_=0.0)
    assert m.calls == [] and pred.shape == (3, 5, 3) and not pred.any()


def test_decode_action_matches_plan_utils():
    a = torch.tensor([[[0.1, 0.2, 0.0, 3.7], [1.0, -1.0, 1.5707963, 2.0]]])
    dec, rep = planning.decode_action(a, push_length=0.1)
    assert rep.tolist() == [[3, 2]] and rep.dtype == torch.int32               # plan_utils.py:14: int() truncation of the length
    assert torch.allclose(dec[0, 0], torch.tensor([0.1, 0.2, 0.0, 0.2]), atol=1e-7)
    assert torch.allclose(dec[0, 1], torch.tensor([1.0, -1.0, 1.0, -1.1]), atol=1e-6)
