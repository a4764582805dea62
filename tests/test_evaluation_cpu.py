"""Host logic of the batched evaluation rollout (no GPU): the frame schedule walked up front must visit exactly the frames the
reference's loop visits (rollout.py:65-96) — pinned by the number of steps and the ground-truth frames of tests/golden/eval_rollout.npz."""
import numpy as np
import pytest

import agx_helpers as H

G = H.load_npz("eval_rollout.npz")


@pytest.mark.parametrize("name", ["stride2", "stride3_gap"])
def test_schedule_matches_the_reference_loop(name):
    from adaptigraph_b200 import evaluation as ev
    n_his = int(G["n_his"])
    pairs, (start, end) = G[f"{name}/pairs"], G[f"{name}/start_end"]
    n_frames = G[f"{name}/obj_pos"].shape[0]
    next_fn = ev.get_next_pair_or_break_episode_pushes if int(G[f"{name}/pushes"]) else ev.get_next_pair_or_break_episode
    sched = ev._schedule(pairs, n_his, n_frames, int(start), int(end), next_fn)
    assert len(sched) == len(G[f"{name}/errors"])                      # one forward per visited pair, like the reference
    assert sched[0] == [int(start), int(end)]
    ends = [e for _, e in sched]
    assert all(b > a for a, b in zip(ends, ends[1:]))                  # "avoid loop": strictly forward in time
    for s, e in sched[1:]:
        assert any((row[n_his - 1] == s and row[n_his] == e) for row in pairs)   # every step is one of the recorded pairs
    # the walk stops exactly where the reference's does: no further pair from the last end frame
    assert next_fn(pairs, n_his, n_frames, ends[-1]) is None or len(sched) == 100


def test_next_pair_functions_follow_the_reference():
    from adaptigraph_b200 import evaluation as ev
    pairs = np.array([[0, 1, 2, 3, 5], [0, 1, 2, 3, 6], [0, 1, 2, 3, 7], [4, 5, 6, 7, 9]])
    assert list(ev.get_next_pair_or_break_episode_pushes(pairs, 4, 12, 3)) == [0, 1, 2, 3, 6]     # the middle one of three
    assert ev.get_next_pair_or_break_episode_pushes(pairs, 4, 12, 5) is None
    assert list(ev.get_next_pair_or_break_episode(pairs, 4, 12, 5)) == [4, 5, 6, 7, 9]            # walks forward to frame 7
    assert ev.get_next_pair_or_break_episode(pairs, 4, 12, 9) is None
