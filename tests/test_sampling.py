"""Particle samplers: the oracle against vectors produced by the reference's own fps_rad_idx / fps
(tests/golden/make_golden_fps.py), and — on the GPU — agx_fps through the reference-signature wrappers against both."""
import numpy as np
import pytest
import torch

import agx_helpers as H
from oracle import sampling_oracle as so

FPS = H.load_npz("fps_cases.npz")
CLOUDS = sorted({k.split("/")[0] for k in FPS})


def _replay_fps_draws(pcd, max_nobj, rr, seed):
    """The reference's order of numpy draws inside fps() (graph.py:11-27, utils.py:14): start, [radius], second start."""
    np.random.seed(seed)
    start_1 = np.random.randint(0, pcd.shape[0])
    radius = float(rr[0]) if len(rr) == 1 else np.random.uniform(rr[0], rr[1])
    start_2 = np.random.randint(min(max_nobj, pcd.shape[0]))
    return start_1, radius, start_2


@pytest.mark.parametrize("name", CLOUDS)
def test_oracle_matches_reference_fps_rad_idx(name):
    pcd = FPS[f"{name}/pcd"]
    for k in range(3):
        np.random.seed(int(FPS[f"{name}/rad{k}/seed"]))
        rand_idx = np.random.randint(pcd.shape[0])
        pts, idx = so.fps_rad_idx(pcd, float(FPS[f"{name}/rad{k}/radius"]), rand_idx)
        np.testing.assert_array_equal(np.asarray(idx).reshape(-1), FPS[f"{name}/rad{k}/idx"])
        np.testing.assert_array_equal(pts, pcd[FPS[f"{name}/rad{k}/idx"]])


@pytest.mark.parametrize("name", CLOUDS)
def test_oracle_matches_reference_fps_wrapper(name):
    pcd = FPS[f"{name}/pcd"]
    for k in range(3):
        max_nobj, rr = int(FPS[f"{name}/fps{k}/max_nobj"]), FPS[f"{name}/fps{k}/range"]
        s1, radius, s2 = _replay_fps_draws(pcd, max_nobj, rr, int(FPS[f"{name}/fps{k}/seed"]))
        idx = so.fps(pcd, max_nobj, radius, s1, s2)
        np.testing.assert_array_equal(idx, FPS[f"{name}/fps{k}/idx"])


def test_oracle_sampler_properties():
    rng = np.random.default_rng(0)
    pos = rng.normal(size=(3, 200, 3)).astype(np.float32)
    idx = so.farthest_point_sampler(pos, 50, [0, 7, 199])
    assert idx.shape == (3, 50) and list(idx[:, 0]) == [0, 7, 199]
    for b in range(3):
        assert len(set(idx[b])) == 50                                   # distinct points: no re-picks before exhaustion
        d = np.linalg.norm(pos[b][:, None] - pos[b][idx[b, :2]][None], axis=-1)
        assert idx[b, 1] == np.argmax(d[:, 0])                          # second pick = farthest from the first


# ------------------------------------------------------------------------------------------- known answers for the DGL sampler
# dgl.geometry.farthest_point_sampler is a third-party operator absent from /root/reference and from this image, so no run of it
# can pin the restatement.  These vectors pin its PUBLISHED semantics instead (DGL src/geometry/cpu/geometry_op_impl.cc,
# FarthestPointSampler): the first pick is start_idx; every later pick scans the points in index order keeping, per point, the
# minimum SQUARED distance to the picks so far, and takes the point whose kept distance is largest under a strict `>` against a
# running maximum that starts at 0 with candidate index 0 -- i.e. the LOWEST index among ties, and index 0 once every kept
# distance is 0.  The expected index lists below were derived by hand from that description on integer lattices (all squared
# distances are exact small integers in fp32, so no rounding question arises); they were NOT produced by oracle/sampling_oracle.py.
#
#   line     x = 0 1 2 3 4 (y = z = 0), start 0, 5 picks
#            after {0}: d = 0 1 4 9 16 -> 4;  after {0,4}: d = 0 1 4 1 0 -> 2;  after {0,4,2}: d = 0 1 0 1 0 -> tie {1,3} -> 1;  then 3
#   square   corners (0,0) (2,0) (0,2) (2,2) and centre (1,1) in the xy plane, start 4 (the centre), 5 picks
#            after {4}: d = 2 2 2 2 0 -> four-way tie -> 0;  after {4,0}: d = 0 2 2 2 0 -> (d to 0 is 4, 4, 8; to the centre 2) tie -> 1;
#            after {4,0,1}: d = 0 0 2 2 0 -> tie {2,3} -> 2;  then 3
#   stack    the same point five times, start 3, 4 picks: every kept distance is 0, nothing beats the running maximum 0 -> 0, 0, 0
#   cube     the 8 corners of the unit cube scaled by 3 in index order zyx (index = 4z + 2y + x), start 5 = (3,0,3), 4 picks
#            after {5}: squared distances 18 9 27 18 9 0 18 9 -> 2 = (0,3,0);  after {5,2}: d = 9 9 0 9 9 0 9 9 -> tie -> 0;
#            after {5,2,0}: d = 0 9 0 9 9 0 9 9 -> tie {1,3,4,6,7} -> 1
#   batch    two clouds of unequal content in one call with per-cloud start indices: `line` with start 2 -> 2, then d = 4 1 0 1 4 ->
#            tie {0,4} -> 0, then d = 0 1 0 1 4 -> 4, then tie {1,3} -> 1;  and `line` reversed (x = 4 3 2 1 0) with start 0 -> 0 4 2 1
DGL_KAT = {
    "line": (np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0], [4, 0, 0]], np.float32), 0, [0, 4, 2, 1, 3]),
    "square": (np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0], [2, 2, 0], [1, 1, 0]], np.float32), 4, [4, 0, 1, 2, 3]),
    "stack": (np.tile(np.array([[0.5, -1.25, 2.0]], np.float32), (5, 1)), 3, [3, 0, 0, 0]),
    "cube": (np.array([[3 * (i & 1), 3 * ((i >> 1) & 1), 3 * (i >> 2)] for i in range(8)], np.float32), 5, [5, 2, 0, 1]),
}
DGL_KAT_BATCH = (np.stack([DGL_KAT["line"][0], DGL_KAT["line"][0][::-1].copy()]), [2, 0], [[2, 0, 4, 1], [0, 4, 2, 1]])


@pytest.mark.parametrize("name", sorted(DGL_KAT))
def test_oracle_sampler_matches_dgl_known_answers(name):
    pos, start, want = DGL_KAT[name]
    np.testing.assert_array_equal(so.farthest_point_sampler(pos[None], len(want), [start])[0], want)


def test_oracle_sampler_matches_dgl_known_answers_batched():
    pos, starts, want = DGL_KAT_BATCH
    np.testing.assert_array_equal(so.farthest_point_sampler(pos, 4, starts), want)


# ------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def sampling():
    import adaptigraph_b200.ops  # noqa: F401  (loads the .so; raises if missing)
    from adaptigraph_b200 import sampling as s
    assert torch.cuda.is_available()
    return s


@pytest.mark.gpu
@pytest.mark.parametrize("name", CLOUDS)
def test_gpu_fps_rad_idx_matches_reference(sampling, name):
    pcd = FPS[f"{name}/pcd"]
    for k in range(3):
        np.random.seed(int(FPS[f"{name}/rad{k}/seed"]))
        pts, idx = sampling.fps_rad_idx(pcd, float(FPS[f"{name}/rad{k}/radius"]))
        np.testing.assert_array_equal(idx, FPS[f"{name}/rad{k}/idx"])
        np.testing.assert_array_equal(pts, pcd[idx])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CLOUDS)
def test_gpu_fps_wrapper_matches_reference(sampling, name):
    pcd = FPS[f"{name}/pcd"]
    for k in range(3):
        rr = FPS[f"{name}/fps{k}/range"]
        np.random.seed(int(FPS[f"{name}/fps{k}/seed"]))
        idx = sampling.fps(pcd, int(FPS[f"{name}/fps{k}/max_nobj"]), float(rr[0]) if len(rr) == 1 else list(rr))
        np.testing.assert_array_equal(idx, FPS[f"{name}/fps{k}/idx"])


@pytest.mark.gpu
def test_gpu_sampler_matches_oracle_batched(sampling):
    rng = np.random.default_rng(3)
    for B, N, k in [(4, 257, 64), (2, 2025, 300), (3, 33, 33), (1, 12800, 40)]:   # 12800 = the largest single-CTA cloud
        pos = rng.normal(size=(B, N, 3)).astype(np.float32)
        start = rng.integers(0, N, B)
        ref = so.farthest_point_sampler(pos, k, start)
        got = sampling.farthest_point_sampler(torch.from_numpy(pos).cuda(), k, torch.from_numpy(start).cuda())
        assert got.dtype == torch.int64 and got.is_cuda
        np.testing.assert_array_equal(got.cpu().numpy(), ref)
    # scalar start index and CPU input (the reference's call form, graph.py:11-12): result comes back on the CPU
    pos = rng.normal(size=(1, 500, 3)).astype(np.float32)
    got = sampling.farthest_point_sampler(torch.from_numpy(pos), 100, start_idx=17)
    np.testing.assert_array_equal(got.numpy(), so.farthest_point_sampler(pos, 100, [17]))
    with pytest.raises(ValueError):
        sampling.farthest_point_sampler(torch.from_numpy(pos), 10, start_idx=500)
    with pytest.raises(ValueError):   # beyond the cluster staging limit: an error, not a fallback
        sampling.farthest_point_sampler(torch.zeros(1, 204801, 3), 4, start_idx=0)


@pytest.mark.gpu
@pytest.mark.parametrize("N,k", [(12801, 40), (20000, 64), (60000, 48), (150000, 32), (204800, 8)])
def test_gpu_cluster_sampler_matches_oracle(sampling, N, k):
    """Clouds beyond one CTA's shared memory run on a thread-block cluster (2, 2, 8, 16, 16 CTAs): same picks as the oracle,
    in both modes, for ragged lengths too."""
    rng = np.random.default_rng(N)
    B = 2
    pos = rng.normal(size=(B, N, 3)).astype(np.float32)
    start = rng.integers(0, N // 2, B)
    got = sampling.farthest_point_sampler(torch.from_numpy(pos).cuda(), k, torch.from_numpy(start).cuda())
    np.testing.assert_array_equal(got.cpu().numpy(), so.farthest_point_sampler(pos, k, start))
    # radius mode on a ragged prefix (second cloud uses 60 % of its points)
    from adaptigraph_b200 import ops
    n_pts = np.array([N, int(0.6 * N)])
    radius = 2.2
    idx, cnt = ops.fps(torch.from_numpy(pos).cuda(), torch.from_numpy(n_pts).int().cuda(), torch.from_numpy(start).int().cuda(), 4096, radius)
    for b in range(B):
        _, ref = so.fps_rad_idx(pos[b, :n_pts[b]], radius, int(start[b]))
        assert int(cnt[b]) == len(ref)
        np.testing.assert_array_equal(idx[b, :len(ref)].cpu().numpy(), ref)


@pytest.mark.gpu
def test_gpu_fps_batch_on_device(sampling):
    """Ragged clouds sampled without a host round trip: every cloud agrees with the oracle's fps() on its own points."""
    rng = np.random.default_rng(5)
    B, N, max_nobj, radius = 5, 600, 100, 0.35
    n_pts = np.array([600, 431, 100, 57, 1])
    pos = rng.normal(0, 0.8, size=(B, N, 3)).astype(np.float32)
    s1 = np.array([rng.integers(0, n) for n in n_pts])
    s2 = np.array([rng.integers(0, min(max_nobj, n)) for n in n_pts])
    idx, cnt = sampling.fps_batch(torch.from_numpy(pos).cuda(), torch.from_numpy(n_pts).int().cuda(), max_nobj, radius,
                                  torch.from_numpy(s1).int().cuda(), torch.from_numpy(s2).int().cuda())
    idx, cnt = idx.cpu().numpy(), cnt.cpu().numpy()
    for b in range(B):
        ref = so.fps(pos[b, :n_pts[b]], max_nobj, radius, int(s1[b]), int(s2[b]))
        assert cnt[b] == len(ref)
        np.testing.assert_array_equal(idx[b, :cnt[b]], ref)


def test_wrapper_host_logic_without_a_gpu(monkeypatch):
    """The reference-signature wrappers (random draws, index composition, numpy in / out) around a stand-in for the CUDA op: with the
    oracle behind `ops.fps`, seeded calls must reproduce the reference's own outputs — the draws happen in the reference's order."""
    from adaptigraph_b200 import sampling

    def fake_fps(pos, n_points, start_idx, max_samples, radius):
        pos, n_points, start_idx = pos.cpu().numpy(), n_points.cpu().numpy(), start_idx.cpu().numpy()
        B = pos.shape[0]
        idx = np.zeros((B, max_samples), np.int32)
        cnt = np.zeros(B, np.int32)
        for b in range(B):
            pts = pos[b, :n_points[b]]
            if radius < 0:
                sel = so.farthest_point_sampler(pts[None], max_samples, [start_idx[b]])[0]
            else:
                sel = so.fps_rad_idx(pts, radius, int(start_idx[b]))[1][:max_samples]
            idx[b, :len(sel)], cnt[b] = sel, len(sel)
        return torch.from_numpy(idx), torch.from_numpy(cnt)

    monkeypatch.setattr(sampling.ops, "fps", fake_fps)
    monkeypatch.setattr(sampling, "_device", lambda device=None: torch.device("cpu"))
    for name in ("blob300", "dup64", "single"):
        pcd = FPS[f"{name}/pcd"]
        for k in range(3):
            np.random.seed(int(FPS[f"{name}/rad{k}/seed"]))
            pts, idx = sampling.fps_rad_idx(pcd, float(FPS[f"{name}/rad{k}/radius"]))
            np.testing.assert_array_equal(idx, FPS[f"{name}/rad{k}/idx"])
            rr = FPS[f"{name}/fps{k}/range"]
            np.random.seed(int(FPS[f"{name}/fps{k}/seed"]))
            got = sampling.fps(pcd, int(FPS[f"{name}/fps{k}/max_nobj"]), float(rr[0]) if len(rr) == 1 else list(rr))
            np.testing.assert_array_equal(got, FPS[f"{name}/fps{k}/idx"])


@pytest.mark.gpu
def test_gpu_sampler_matches_dgl_known_answers(sampling):
    """The CUDA sampler against the hand-derived vectors of the published DGL operator (ties -> lowest index, all-zero -> 0)."""
    for name, (pos, start, want) in DGL_KAT.items():
        got = sampling.farthest_point_sampler(torch.from_numpy(pos[None]).cuda(), len(want), start_idx=start)
        np.testing.assert_array_equal(got.cpu().numpy()[0], want, err_msg=name)
    pos, starts, want = DGL_KAT_BATCH
    got = sampling.farthest_point_sampler(torch.from_numpy(pos).cuda(), 4, torch.tensor(starts).cuda())
    np.testing.assert_array_equal(got.cpu().numpy(), want)
