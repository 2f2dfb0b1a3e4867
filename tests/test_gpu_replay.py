"""GPU parity: the HBM arena (ring cursor, gathers, write-time hindsight flush, return recurrence, sample-time
relabelling) vs goldens produced by running the reference (tests/golden/{ring,nstep,her}.npz) and vs the CPU oracle.
Bit-exact for indices, goals, done masks, episode steps; <= 1e-5 relative (in practice ~1e-7) for rewards / returns."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import cpu_restatement as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R(fdql):
    from fastdeepqlearning_b200 import Replay
    return Replay


def npy(t):
    return t.detach().cpu().numpy()


def test_ring_cursor_and_gather_golden(R):
    g = load_golden("ring")
    ring = R.ReplayMemory(int(g["maxlen"]), int(g["B"]), int(g["T"]), stage_rows=37)
    for i in range(len(g["in_reward"])):
        ring.add({"obs_1d": g["in_obs"][i], "action": g["in_act"][i], "desired_goal": g["in_goal"][i],
                  "reward": float(g["in_reward"][i]), "task_done": bool(g["in_done"][i]), "episode_step": int(g["in_step"][i])})
        assert ring._top == g["tops"][i] and len(ring) == g["lens"][i]
        if i == 199:
            part = ring.temporal_sample(starts=g["partial_starts"])
            assert len(ring) == g["partial_len"]
            for k in ring.keys:
                np.testing.assert_array_equal(npy(part[k]), g[f"partial_{k}"].astype(np.float32), err_msg=k)
    assert len(ring) == int(g["maxlen"]) - 1  # quirk Q1
    win = ring.temporal_sample(starts=g["starts"])
    flat = ring.sample(idx=g["flat_idx"])
    mem = ring.memory
    for k in ring.keys:
        np.testing.assert_array_equal(npy(mem[k]), g[f"mem_{k}"].astype(np.float32), err_msg=k)
        np.testing.assert_array_equal(npy(win[k]), g[f"win_{k}"].astype(np.float32), err_msg=k)
        np.testing.assert_array_equal(npy(flat[k]), g[f"flat_{k}"].astype(np.float32), err_msg=k)
    # device-drawn streams honour the reference's ranges: starts in [0, len-T), sample() in [0, len)
    s, _, _ = ring.draw_streams(4096)
    assert int(s.min()) >= 0 and int(s.max()) < len(ring) - int(g["T"])
    assert int(s.max()) > (len(ring) - int(g["T"])) * 0.9
    out = ring.temporal_sample()
    assert tuple(out["obs_1d"].shape) == (int(g["T"]), int(g["B"]), 6)


def test_oversample_error(R, fdql):
    small = R.ReplayMemory(100, 8, 5)
    for i in range(7):
        small.add({"x": float(i)})
    with pytest.raises(fdql.OversampleError):
        small.temporal_sample()
    with pytest.raises(fdql.OversampleError):
        small.sample()


def test_temporal_consistency_like_reference_test(R):
    """tests/test_replays.py:60-84 of the reference: [T=10, B=256, obs=10] windows are time-contiguous."""
    ring = R.AsyncReplayMemory(1000, 256, 10)
    for i in range(600):
        ring.add({"obs": np.full(10, i, np.float32), "reward": 0.0})
    assert len(ring) == 600
    from fastdeepqlearning_b200.Replay.wrappers import TorchDataLoader
    loader = TorchDataLoader(ring, device="cuda:0")
    assert loader.ready()
    xp = loader.temporal_sample()
    obs = xp["obs"]
    assert tuple(obs.shape) == (10, 256, 10)
    assert bool((obs[1:] == obs[:-1] + 1).all())


def test_size_like_reference_test(R):
    """tests/test_replays.py:36-57: len tracks adds and saturates at maxlen for the Async front."""
    ring = R.AsyncReplayMemory(50, 4, 2)
    for i in range(120):
        ring.add({"x": float(i)})
        assert len(ring) == min(i + 1, 50)


def test_nstep_known_answer_commit_kernel(R):
    """tests/test_replays.py:16-33: reward 1 only at the last of 1000 steps -> mc_return == gamma^(n-1-step); the
    commit kernel runs the reference's sequential fp64-step/fp32-store recurrence, so the result is bit-exact."""
    g = load_golden("nstep")
    n = 1000
    ring = R.ReplayMemory(1001, 128, 1)
    r = np.zeros((n, 1), np.float32)
    r[-1] = 1
    ring.add_rows({"reward": r, "episode_step": np.arange(n, dtype=np.float32).reshape(-1, 1),
                   "episode_done": (np.arange(n) == n - 1).astype(np.float32).reshape(-1, 1),
                   "mc_return": np.zeros((n, 1), np.float32)}, episode_lengths=[n], with_returns=True, gamma=0.99)
    got = npy(ring.memory["mc_return"])[:n]
    np.testing.assert_array_equal(got, g["kat_mc_return"])
    assert np.allclose(got.reshape(-1), 0.99 ** (n - 1 - np.arange(n)))
    xp = ring.sample()
    assert np.allclose(npy(xp["mc_return"]), 0.99 ** (n - 1 - npy(xp["episode_step"])))
    # random episodes, bit-exact against the reference's stored returns
    ring2 = R.ReplayMemory(400, 8, 2)
    L = g["rand_lengths"]
    N = int(L.sum())
    ring2.add_rows({"reward": g["rand_reward"].reshape(-1, 1), "episode_done": g["rand_done"].astype(np.float32).reshape(-1, 1),
                    "mc_return": np.zeros((N, 1), np.float32)}, episode_lengths=L, with_returns=True, gamma=float(g["rand_gamma"]))
    np.testing.assert_array_equal(npy(ring2.memory["mc_return"])[:N], g["rand_mc_return"])


def _episode_cols(g, name):
    lengths = g[f"{name}_lengths"]
    N = int(lengths.sum())
    ends = np.cumsum(lengths) - 1
    starts_ep = ends - lengths + 1
    ep_of = np.repeat(np.arange(len(lengths)), lengths)
    cols = {"obs_1d": g[f"{name}_in_obs"], "action": g[f"{name}_in_action"],
            "achieved_goal": g[f"{name}_in_ag"].astype(np.float32), "desired_goal": g[f"{name}_in_dg"].astype(np.float32),
            "reward": g[f"{name}_in_reward"].astype(np.float32).reshape(-1, 1),
            "task_done": g[f"{name}_in_task_done"].astype(np.float32).reshape(-1, 1),
            "episode_done": np.isin(np.arange(N), ends).astype(np.float32).reshape(-1, 1),
            "episode_step": (np.arange(N) - starts_ep[ep_of]).astype(np.float32).reshape(-1, 1),
            "mc_return": np.zeros((N, 1), np.float32)}
    return cols, lengths, starts_ep, ends, ep_of


@pytest.mark.parametrize("name", ["bitflip", "all_geq", "first_geq"])
@pytest.mark.parametrize("mode", ["final", "random"])
def test_write_time_hindsight_flush_golden(R, fdql, name, mode):
    """her.py:24-95 + nstep_return.py: per episode the ring receives L real rows then L hindsight rows; the golden is
    the row stream the unmodified reference stored for the same episodes and the same injected goal picks."""
    g = load_golden("her")
    gamma = float(g["gamma"])
    cols, lengths, starts_ep, ends, _ = _episode_cols(g, name)
    picks = g[f"{name}_picks"]
    ring = R.ReplayMemory(4096, 8, 2)
    ring.set_reward_op(fdql.RewardOp.coerce(name), gamma)
    for e, L in enumerate(lengths):
        sl = slice(starts_ep[e], ends[e] + 1)
        begin = ring.add_rows({k: v[sl] for k, v in cols.items()}, episode_lengths=[L], with_returns=True)
        goal = begin + (L - 1 if mode == "final" else int(picks[e]))
        ring.add_hindsight_rows([begin], [L], [goal], with_returns=True)
    n = len(ring)
    assert n == 2 * lengths.sum()
    mem = {k: npy(v)[:n] for k, v in ring.memory.items()}
    for k in ("obs_1d", "action", "achieved_goal", "desired_goal", "task_done", "episode_done", "episode_step"):
        np.testing.assert_array_equal(mem[k], g[f"{name}_{mode}_{k}"].astype(np.float32), err_msg=k)
    np.testing.assert_array_equal(mem["reward"], g[f"{name}_{mode}_reward"])
    np.testing.assert_array_equal(mem["mc_return"], g[f"{name}_{mode}_mc_return"])  # exact sequential recurrence


@pytest.mark.parametrize("name", ["bitflip", "all_geq", "first_geq"])
def test_sample_time_relabel_equals_reference_rows(R, fdql, name):
    """Sample-time relabelling of every row (T=1, exact episode_step) reproduces the hindsight rows the reference
    stored at write time with the same goal pick (golden her.npz, mode=random)."""
    g = load_golden("her")
    gamma = float(g["gamma"])
    cols, lengths, starts_ep, ends, ep_of = _episode_cols(g, name)
    picks = g[f"{name}_picks"]
    N = int(lengths.sum())
    ring = R.ReplayMemory(N + 5, 8, 1)
    ring.set_reward_op(fdql.RewardOp.coerce(name), gamma)
    ring.add_rows(cols, episode_lengths=lengths, with_returns=True)
    starts = np.arange(N)
    out = ring.temporal_sample(starts=starts, flags=np.ones(N, np.uint8), goal_rows=(starts_ep + picks)[ep_of],
                               exact_episode_step=True, length=N)
    hs = np.concatenate([np.arange(2 * s + L, 2 * s + 2 * L) for s, L in zip(starts_ep, lengths)])
    for k in ("obs_1d", "action", "achieved_goal", "desired_goal", "task_done", "episode_done", "episode_step"):
        np.testing.assert_array_equal(npy(out[k])[0], g[f"{name}_random_{k}"][hs].astype(np.float32), err_msg=k)
    np.testing.assert_allclose(npy(out["reward"])[0], g[f"{name}_random_reward"][hs], rtol=1e-6)
    np.testing.assert_allclose(npy(out["mc_return"])[0], g[f"{name}_random_mc_return"][hs], rtol=1e-5, atol=1e-6)


def _synthetic(rng, n_eps, max_len, G, obs=5, act=2, p_hit=0.15, fixed_len=None):
    lengths = np.full(n_eps, fixed_len) if fixed_len else rng.integers(1, max_len + 1, n_eps)
    N = int(lengths.sum())
    ends = np.cumsum(lengths) - 1
    starts_ep = ends - lengths + 1
    ep_of = np.repeat(np.arange(n_eps), lengths)
    ag = rng.integers(0, 2, (N, G)).astype(np.float32)
    dg = rng.integers(0, 2, (n_eps, G)).astype(np.float32)[ep_of]
    hit = rng.random(N) < p_hit  # make some achieved goals equal the desired / a revisited goal
    ag[hit] = dg[hit]
    rr, dd = O.reward_bitflip(ag, dg)
    cols = {"obs_1d": rng.standard_normal((N, obs)).astype(np.float32), "action": rng.standard_normal((N, act)).astype(np.float32),
            "achieved_goal": ag, "desired_goal": dg, "reward": rr.astype(np.float32).reshape(-1, 1),
            "task_done": dd.astype(np.float32).reshape(-1, 1),
            "episode_done": np.isin(np.arange(N), ends).astype(np.float32).reshape(-1, 1),
            "episode_step": (np.arange(N) - starts_ep[ep_of]).astype(np.float32).reshape(-1, 1)}
    cols["mc_return"] = O.segmented_returns(cols["reward"], cols["episode_done"], 0.98).reshape(-1, 1)
    return cols, lengths, starts_ep, ends, ep_of


@pytest.mark.parametrize("T,G,max_len", [(1, 16, 40), (2, 16, 130), (5, 3, 70), (50, 64, 200), (2, 100, 33), (33, 8, 90)])
def test_sample_time_relabel_vs_oracle(R, fdql, T, G, max_len):
    """Windows with hindsight flags vs oracle.sample_time_relabel (defined through the reference's write-time rows):
    windows straddling episode ends, episodes longer than 32/64/128 rows, goal widths that are not multiples of 4."""
    rng = np.random.default_rng(T * 100 + G)
    cols, lengths, starts_ep, ends, ep_of = _synthetic(rng, 60, max_len, G)
    N = int(lengths.sum())
    ring = R.ReplayMemory(N + 3, 16, T)
    ring.set_reward_op(fdql.RewardOp.bitflip(), 0.98)
    ring.add_rows(cols, episode_lengths=lengths)
    es, ee = ring.episode_extents()
    np.testing.assert_array_equal(npy(es)[:N], starts_ep[ep_of])
    np.testing.assert_array_equal(npy(ee)[:N], ends[ep_of])
    B = 700
    starts = rng.integers(0, N - T, B)
    flags = rng.random(B) < 0.8
    goal_rows = np.array([rng.integers(s, ends[ep_of[s]] + 1) for s in starts])  # "future" incl. the row itself
    want = O.sample_time_relabel(cols, starts, T, flags, goal_rows, starts_ep[ep_of], ends[ep_of], O.reward_bitflip, 0.98)
    got = ring.temporal_sample(starts=starts, flags=flags.astype(np.uint8), goal_rows=goal_rows, exact_episode_step=True,
                               aux=T > 1, length=N)
    for k in cols:
        if k in ("reward", "mc_return"):
            np.testing.assert_allclose(npy(got[k]), want[k], rtol=1e-5, atol=1e-6, err_msg=k)
        else:
            np.testing.assert_array_equal(npy(got[k]), want[k], err_msg=k)
    if T > 1:
        mask, contig = O.learner_preprocess(want["task_done"], want["episode_step"])
        np.testing.assert_array_equal(npy(got["mask"]), mask.astype(np.float32))
        np.testing.assert_array_equal(npy(got["is_contiguous"]), contig.astype(np.float32))
        np.testing.assert_allclose(npy(got["loss_weight"]), O.upstream_weight(contig, T), rtol=1e-6, atol=1e-12)
        # default (non-exact) mode: episode_step may differ, the learner-visible mask / is_contiguous may not
        fast = ring.temporal_sample(starts=starts, flags=flags.astype(np.uint8), goal_rows=goal_rows, aux=True, length=N)
        np.testing.assert_array_equal(npy(fast["mask"]), mask.astype(np.float32))
        np.testing.assert_array_equal(npy(fast["is_contiguous"]), contig.astype(np.float32))
        for k in ("desired_goal", "task_done", "reward", "mc_return", "obs_1d"):
            np.testing.assert_array_equal(npy(fast[k]), npy(got[k]), err_msg=k)


def test_ring_wrap_and_stale_rows(R, fdql):
    """Episodes that wrap around the end of the ring keep valid extents; relabelling across the wrap matches the oracle
    evaluated on the unrolled episode."""
    rng = np.random.default_rng(5)
    cols, lengths, starts_ep, ends, ep_of = _synthetic(rng, 12, 0, 8, fixed_len=25)
    N = int(lengths.sum())  # 300 rows into a ring of 128: wraps twice
    cap = 128
    ring = R.ReplayMemory(cap, 8, 2)
    ring.set_reward_op(fdql.RewardOp.bitflip(), 0.98)
    for e in range(len(lengths)):
        ring.add_rows({k: v[starts_ep[e]:ends[e] + 1] for k, v in cols.items()}, episode_lengths=[25])
    assert ring._top == N % cap and len(ring) == cap - 1
    # the last 5 episodes are fully resident; episode 10 (rows 250..274 -> 122..127, 0..18) wraps the end of the ring
    e = 10
    phys = (np.arange(starts_ep[e], ends[e] + 1)) % cap
    assert phys[0] > phys[-1]
    es, ee = ring.episode_extents()
    assert int(es[phys[3]]) == phys[0] and int(ee[phys[3]]) == phys[-1]
    sub = {k: v[starts_ep[e]:ends[e] + 1] for k, v in cols.items()}
    j = 2
    want = O.sample_time_relabel(sub, np.array([j]), 1, np.array([True]), np.array([20]), np.zeros(25, int), np.full(25, 24),
                                 O.reward_bitflip, 0.98)
    got = ring.temporal_sample(starts=np.array([phys[j]]), flags=np.array([1], np.uint8), goal_rows=np.array([phys[20]]),
                               exact_episode_step=True, n=1)
    for k in ("desired_goal", "task_done", "episode_step"):
        np.testing.assert_array_equal(npy(got[k])[0], want[k][0], err_msg=k)
    np.testing.assert_allclose(npy(got["mc_return"])[0], want["mc_return"][0], rtol=1e-5)


def test_device_streams_properties(R, fdql):
    from fastdeepqlearning_b200 import _lib as L
    rng = np.random.default_rng(9)
    cols, lengths, starts_ep, ends, ep_of = _synthetic(rng, 300, 50, 4)
    N = int(lengths.sum())
    ring = R.ReplayMemory(N + 1, 64, 2)
    ring.set_reward_op(fdql.RewardOp.bitflip(), 0.98)
    ring.add_rows(cols, episode_lengths=lengths)
    for mode in (L.GOAL_FINAL, L.GOAL_RANDOM, L.GOAL_FUTURE):
        s, f, gr = ring.draw_streams(20000, goal_mode=mode, relabel_prob=0.8)
        s, f, gr = npy(s), npy(f), npy(gr)
        assert s.min() >= 0 and s.max() < N - 2
        assert abs(f.mean() - 0.8) < 0.02
        on = f == 1
        assert (gr[on] >= starts_ep[ep_of[s[on]]]).all() and (gr[on] <= ends[ep_of[s[on]]]).all()
        if mode == L.GOAL_FINAL:
            assert (gr[on] == ends[ep_of[s[on]]]).all()
        if mode == L.GOAL_FUTURE:
            last = s == ends[ep_of[s]]
            assert (gr[on & ~last] > s[on & ~last]).all() and (gr[on & last] == s[on & last]).all()
    a = ring.draw_streams(100, relabel_prob=0.5)
    ring._rng_counter -= 1
    b = ring.draw_streams(100, relabel_prob=0.5)
    assert all(bool((x == y).all()) for x, y in zip(a, b))  # counter-based: same (seed, counter) -> same streams


def test_full_size_properties(R, fdql):
    """BASELINE.json shapes (obs 64, act 8, goal 16, B 4096, T 2) on a 2e6-row ring: properties that do not need the oracle
    at that size -- gather == torch index of the arena views, relabelled goal == achieved_goal[goal_row], done mask
    consistent with the reward, return recurrence G_t = r_t + gamma*G_{t+1} inside each window."""
    import torch
    torch.manual_seed(0)
    Lep, n_eps = 128, 15625
    N = Lep * n_eps
    dev = "cuda"
    ag = (torch.rand(N, 16, device=dev) < 0.5).float()
    dg = (torch.rand(n_eps, 16, device=dev) < 0.5).float().repeat_interleave(Lep, 0)
    match = (ag == dg).all(-1, keepdim=True)
    step = torch.arange(N, device=dev).remainder(Lep).float().unsqueeze(-1)
    cols = {"obs_1d": torch.randn(N, 64, device=dev), "action": torch.rand(N, 8, device=dev) * 2 - 1, "achieved_goal": ag,
            "desired_goal": dg, "reward": match.float() - 1, "task_done": match.float(), "episode_done": (step == Lep - 1).float(),
            "episode_step": step, "mc_return": torch.zeros(N, 1, device=dev)}
    ring = R.ReplayMemory(N + 1, 4096, 2)
    ring.set_reward_op(fdql.RewardOp.bitflip(), 0.99)
    ring.add_rows(cols, episode_lengths=torch.full((n_eps,), Lep), with_returns=True)
    starts, flags, goals = ring.draw_streams(4096, relabel_prob=0.8)
    out = ring.temporal_sample(starts=starts, flags=flags, goal_rows=goals, aux=True)
    mem = ring.memory
    idx = torch.stack([starts, starts + 1])
    for k in ("obs_1d", "action", "achieved_goal", "episode_done"):
        assert torch.equal(out[k], mem[k][idx]), k
    f = flags.bool()
    same_ep = (idx // Lep == (starts // Lep)[None])  # rows of the window inside the start row's episode
    rel = same_ep & f[None]
    gstar = mem["achieved_goal"][goals]
    assert torch.equal(out["desired_goal"][rel], gstar[None].expand(2, -1, -1)[rel])
    assert torch.equal(out["desired_goal"][~rel], mem["desired_goal"][idx][~rel])
    hit = (out["achieved_goal"] == out["desired_goal"]).all(-1, keepdim=True)
    assert torch.equal(out["task_done"][rel], hit[rel].float())
    assert torch.equal(out["reward"][rel], hit[rel].float() - 1)
    both = rel[0] & rel[1]
    g0, g1, r0 = out["mc_return"][0, both, 0], out["mc_return"][1, both, 0], out["reward"][0, both, 0]
    torch.testing.assert_close(g0, r0 + 0.99 * g1, rtol=1e-5, atol=1e-5)
    unrel = ~f
    assert torch.equal(out["mc_return"][:, unrel], mem["mc_return"][idx][:, unrel])
    assert torch.equal(out["mask"], 1 - out["task_done"])


@pytest.mark.parametrize("T,G,max_len", [(2, 16, 130), (50, 64, 200), (5, 3, 70)])
def test_descriptor_walking_gather_kernel(R, fdql, T, G, max_len):
    """Rows wider than 128 float4 take the descriptor-walking kernel; force it on ordinary shapes and repeat the parity run."""
    lib = fdql.lib()
    old = lib.fdql_debug_force_generic_gather(1)
    try:
        test_sample_time_relabel_vs_oracle(R, fdql, T, G, max_len)
    finally:
        lib.fdql_debug_force_generic_gather(old)


@pytest.mark.parametrize("T,G,max_len", [(2, 16, 130), (50, 64, 200), (5, 3, 70), (2, 100, 33)])
def test_full_vector_relabel_scan_with_bitflip(R, fdql, T, G, max_len):
    """The bitflip functor normally takes the hash-assisted scan (16 B scan record per tail row, hash matches verified on the
    full vectors); force the full-vector scan that the other functors use and repeat the parity run."""
    lib = fdql.lib()
    old = lib.fdql_debug_force_generic_gather(2)
    try:
        test_sample_time_relabel_vs_oracle(R, fdql, T, G, max_len)
    finally:
        lib.fdql_debug_force_generic_gather(old)


@pytest.mark.parametrize("T,G,max_len", [(2, 16, 130), (5, 3, 70), (33, 8, 90)])
def test_hash_scan_with_per_pass_suffix_scan(R, fdql, T, G, max_len):
    """The hash-assisted scan has two return recomputes (scan-free Horner form for T <= 32, per-pass suffix scan otherwise);
    force the second on small windows too."""
    lib = fdql.lib()
    old = lib.fdql_debug_force_generic_gather(8 | 4)
    try:
        test_sample_time_relabel_vs_oracle(R, fdql, T, G, max_len)
        test_hash_scan_verifies_matches_and_nans(R, fdql)
    finally:
        lib.fdql_debug_force_generic_gather(old)


@pytest.mark.parametrize("T,G,max_len", [(1, 16, 40), (2, 16, 130), (5, 3, 70), (50, 64, 200)])
def test_warp_per_window_kernels(R, fdql, T, G, max_len):
    """Plain and bitflip gathers normally take the tile kernel (thread-per-window scalars + warp-per-window wide keys); route
    them through the warp-per-window kernels (scan-free Horner returns for T <= 32) and repeat the parity runs."""
    lib = fdql.lib()
    old = lib.fdql_debug_force_generic_gather(8)
    try:
        test_sample_time_relabel_vs_oracle(R, fdql, T, G, max_len)
        test_hash_scan_verifies_matches_and_nans(R, fdql)
        test_ring_cursor_and_gather_golden(R)
    finally:
        lib.fdql_debug_force_generic_gather(old)


@pytest.mark.parametrize("links", [True, False])
@pytest.mark.parametrize("tile", [32, 256])
@pytest.mark.parametrize("T,G,max_len", [(1, 16, 40), (2, 16, 130), (5, 3, 70), (50, 64, 200), (2, 100, 33), (32, 8, 90)])
def test_tile_kernel_on_small_batches(R, fdql, tile, T, G, max_len, links):
    """Batches below ~48K windows take the warp-per-window kernels; force the tile kernel (thread-per-window scalar phase,
    warp-per-window wide-key phase) on the small parity cases, with partial and full tiles, with the link records (chain of equal
    achieved goals + goal-agnostic return: O(hits) relabelled returns) and with the tail scan (exact return recurrence)."""
    lib = fdql.lib()
    old = lib.fdql_debug_force_generic_gather((tile << 8) | (0 if links else 16))
    try:
        test_sample_time_relabel_vs_oracle(R, fdql, T, G, max_len)
        test_hash_scan_verifies_matches_and_nans(R, fdql)
        test_ring_cursor_and_gather_golden(R)
    finally:
        lib.fdql_debug_force_generic_gather(old)


def test_hash_scan_verifies_matches_and_nans(R, fdql):
    """Exactness of the hash-assisted scan does not rest on the hash: -0.0 == +0.0 must match, NaN never matches (not even the
    goal row itself), and equal rows elsewhere in the tail are found."""
    L = 40
    ag = np.arange(L * 4, dtype=np.float32).reshape(L, 4)
    ag[10] = [0.0, 1.0, 2.0, 3.0]
    ag[20] = [-0.0, 1.0, 2.0, 3.0]      # equals row 10 as floats, differs in bits
    ag[25] = [0.0, 1.0, 2.0, 3.0]
    ag[30] = [np.nan, 1.0, 2.0, 3.0]    # a goal row holding NaN matches nothing, itself included
    dg = np.tile(np.array([[9.0, 9.0, 9.0, 9.0]], np.float32), (L, 1))
    step = np.arange(L, dtype=np.float32).reshape(-1, 1)
    cols = {"achieved_goal": ag, "desired_goal": dg, "reward": np.full((L, 1), -1, np.float32), "task_done": np.zeros((L, 1), np.float32),
            "episode_done": (step == L - 1).astype(np.float32), "episode_step": step, "mc_return": np.zeros((L, 1), np.float32)}
    ring = R.ReplayMemory(64, 4, 1)
    ring.set_reward_op(fdql.RewardOp.bitflip(), 0.9)
    ring.add_rows(cols, episode_lengths=[L], with_returns=True)
    starts = np.arange(L)
    for goal in (10, 20, 30, 39):
        got = ring.temporal_sample(starts=starts, flags=np.ones(L, np.uint8), goal_rows=np.full(L, goal), exact_episode_step=True, length=L)
        with np.errstate(invalid="ignore"):
            want = O.sample_time_relabel(cols, starts, 1, np.ones(L, bool), np.full(L, goal), np.zeros(L, int), np.full(L, L - 1),
                                         O.reward_bitflip, 0.9)
        for k in ("task_done", "episode_step", "reward"):
            np.testing.assert_array_equal(npy(got[k]), want[k], err_msg=f"{k} goal={goal}")
        np.testing.assert_allclose(npy(got["mc_return"]), want["mc_return"], rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(npy(got["desired_goal"]), want["desired_goal"])
    d = npy(ring.temporal_sample(starts=starts, flags=np.ones(L, np.uint8), goal_rows=np.full(L, 20), length=L)["task_done"]).reshape(-1)
    assert d[0] == 1 and d[10] == 1 and d[20] == 1 and d[25] == 1 and d.sum() == 4  # row 0 is arange(4) = [0,1,2,3] too
    d = npy(ring.temporal_sample(starts=starts, flags=np.ones(L, np.uint8), goal_rows=np.full(L, 30), length=L)["task_done"]).reshape(-1)
    assert d.sum() == 0


def test_wide_rows_take_the_generic_kernel(R):
    rng = np.random.default_rng(2)
    N = 300
    cols = {"obs": rng.standard_normal((N, 600)).astype(np.float32), "reward": rng.standard_normal((N, 1)).astype(np.float32)}
    ring = R.ReplayMemory(N + 1, 8, 3)
    ring.add_rows(cols)
    starts = rng.integers(0, N - 3, 64)
    out = ring.temporal_sample(starts=starts)
    idx = np.arange(3)[:, None] + starts[None]
    np.testing.assert_array_equal(npy(out["obs"]), cols["obs"][idx])
    np.testing.assert_array_equal(npy(out["reward"]), cols["reward"][idx])


def test_snapshot_restore_roundtrip(R, fdql):
    """state_dict / load_state_dict: same rows, cursor and episode table; scan and link records are rebuilt on the device, so a
    relabelled window batch drawn with the same streams is identical (incl. a wrapped ring and write-time hindsight copies)."""
    rng = np.random.default_rng(77)
    cols, lengths, starts_ep, ends, ep_of = _synthetic(rng, 40, 60, 8)
    N = int(lengths.sum())
    ring = R.ReplayMemory(N - 37, 16, 2)  # smaller than the data: the ring wraps and the oldest episodes are partly overwritten
    ring.set_reward_op(fdql.RewardOp.bitflip(), 0.98)
    off = 0
    for L in lengths:  # episode by episode so that extents of overwritten episodes go stale like in real use
        ring.add_rows({k: v[off:off + L] for k, v in cols.items()}, episode_lengths=[int(L)])
        off += L
    sd = ring.state_dict()
    twin = R.ReplayMemory(N - 37, 16, 2)
    twin.load_state_dict(sd)
    assert (twin._top, len(twin)) == (ring._top, len(ring))
    for k in ring.keys:
        np.testing.assert_array_equal(npy(twin.memory[k]), npy(ring.memory[k]), err_msg=k)
    n = len(ring)
    es, ee = (npy(x) for x in ring.episode_extents())
    starts = rng.integers(0, n - 2, 3000)
    ok = es[starts] >= 0
    starts = starts[ok]
    tail = (ee[starts] - starts) % (N - 37)
    goal_rows = (starts + (rng.random(len(starts)) * (tail + 1)).astype(np.int64)) % (N - 37)
    flags = (rng.random(len(starts)) < 0.8).astype(np.uint8)
    a = ring.temporal_sample(starts=starts, flags=flags, goal_rows=goal_rows, aux=True, exact_episode_step=True)
    b = twin.temporal_sample(starts=starts, flags=flags, goal_rows=goal_rows, aux=True, exact_episode_step=True)
    for k in a:
        np.testing.assert_array_equal(npy(a[k]), npy(b[k]), err_msg=k)


@pytest.mark.parametrize("n,T", [(5000, 2), (300, 2), (2048, 5), (1500, 40)])
def test_fused_draw_equals_separate_launches(R, fdql, n, T):
    """fdql_sample_gather_draw (streams drawn inside the tile gather) vs fdql_sample_streams + fdql_sample_gather with the same
    seed and counter: identical streams, identical batch; small batches / long windows take the two-launch fallback."""
    rng = np.random.default_rng(n + T)
    cols, lengths, starts_ep, ends, ep_of = _synthetic(rng, 60, 90, 16)
    ring = R.ReplayMemory(int(lengths.sum()) + 5, 16, T, seed=1234)
    ring.set_reward_op(fdql.RewardOp.bitflip(), 0.98)
    ring.add_rows(cols, episode_lengths=lengths)
    for relabel_prob in (0.8, 0.0):
        c = ring._rng_counter
        s, f, g = ring.draw_streams(n, relabel_prob=relabel_prob)
        want = ring.temporal_sample(starts=s, flags=f, goal_rows=g, aux=True)
        want = {k: v.clone() for k, v in want.items()}
        ring._rng_counter = c
        got = ring.temporal_sample(n=n, relabel_prob=relabel_prob, aux=True)
        assert ring._rng_counter == c + 1
        s2, f2, g2 = ring.last_streams
        np.testing.assert_array_equal(npy(s2), npy(s))
        if relabel_prob > 0:
            np.testing.assert_array_equal(npy(f2), npy(f))
            np.testing.assert_array_equal(npy(g2), npy(g))
            assert 0.5 < float(f.float().mean()) < 0.95
        for k in want:
            np.testing.assert_array_equal(npy(got[k]), npy(want[k]), err_msg=k)
