"""Pins oracle/cpu_restatement.py against outputs of the unmodified reference
(tests/golden/*.npz, produced by oracle/make_goldens.py) and the reference's own KAT
(/root/reference/tests/test_replays.py:16-33).  CPU only."""
import numpy as np
import pytest

from oracle import cpu_restatement as O
from conftest import load_golden


def test_ring_cursor_and_gather_match_reference():
    g = load_golden("ring")
    ring = O.RingOracle(int(g["maxlen"]), int(g["B"]), int(g["T"]))
    for i in range(len(g["in_reward"])):
        ring.add({"obs_1d": g["in_obs"][i], "action": g["in_act"][i], "desired_goal": g["in_goal"][i],
                  "reward": float(g["in_reward"][i]), "task_done": bool(g["in_done"][i]),
                  "episode_step": int(g["in_step"][i])})
        assert ring._top == g["tops"][i] and len(ring) == g["lens"][i]
        if i == 199:
            part = ring[ring.window_indices(g["partial_starts"])]
            assert len(ring) == g["partial_len"]
            for k, v in part.items():
                np.testing.assert_array_equal(v, g[f"partial_{k}"])
    assert len(ring) == int(g["maxlen"]) - 1  # quirk Q1
    win = ring.temporal_sample(starts=g["starts"])
    flat = ring.sample(idx=g["flat_idx"])
    for k in ring.memory:
        np.testing.assert_array_equal(ring.memory[k], g[f"mem_{k}"])
        assert ring.memory[k].dtype == g[f"mem_{k}"].dtype
        np.testing.assert_array_equal(win[k], g[f"win_{k}"])
        np.testing.assert_array_equal(flat[k], g[f"flat_{k}"])
    assert bool(g["oversample_raised"])
    small = O.RingOracle(100, 8, 5)
    for i in range(7):
        small.add({"x": float(i)})
    with pytest.raises(O.OversampleError):
        small.temporal_sample()


def test_nstep_known_answer_from_reference_tests():
    # tests/test_replays.py:16-33: reward 1 at the last of 1000 steps -> gamma^(n-1-step)
    n, disc = 1000, 0.99
    r = np.zeros(n, np.float32)
    r[-1] = 1
    got = O.mc_return_chrono(r, disc)
    assert np.allclose(got, disc ** (n - 1 - np.arange(n)))
    g = load_golden("nstep")
    np.testing.assert_array_equal(got, g["kat_mc_return"].reshape(-1))  # bit-exact vs the reference loop


def test_nstep_random_episodes_bit_exact():
    g = load_golden("nstep")
    got = O.segmented_returns(g["rand_reward"], g["rand_done"], float(g["rand_gamma"]))
    np.testing.assert_array_equal(got, g["rand_mc_return"].reshape(-1))
    f64 = np.concatenate([O.mc_return_chrono_f64(seg, float(g["rand_gamma"]))
                          for seg in np.split(g["rand_reward"], np.cumsum(g["rand_lengths"])[:-1])])
    np.testing.assert_allclose(got, f64, rtol=1e-5, atol=1e-6)


def test_nstep_quirk_q3_row_stream():
    g = load_golden("nstep")
    sink = O.RingOracle(100, 4, 2)
    w = O.NStepOracle(sink, 3, 0.9)
    rs = g["q3_rewards"]
    for t, r in enumerate(rs):
        w.add({"reward": float(r), "episode_done": t == len(rs) - 1, "episode_step": t})
    assert len(sink) == int(g["q3_n_rows"]) == len(rs) + 1
    np.testing.assert_array_equal(sink.memory["mc_return"][:len(sink)], g["q3_mc_return"])
    np.testing.assert_array_equal(sink.memory["episode_step"][:len(sink)], g["q3_step"])


@pytest.mark.parametrize("name", ["bitflip", "all_geq", "first_geq"])
@pytest.mark.parametrize("mode", ["final", "random"])
def test_her_row_stream_matches_reference(name, mode):
    g = load_golden("her")
    gamma = float(g["gamma"])
    fn = O.REWARD_OPS[name]
    lengths, picks = g[f"{name}_lengths"], list(g[f"{name}_picks"])
    sink = O.RingOracle(4096, 8, 2)
    inner = O.NStepOracle(sink, 1000, gamma)
    it = iter(picks)
    her = O.HindsightOracle(inner, fn, mode=mode, goal_picker=lambda L: next(it))
    gdt = np.int64 if name == "bitflip" else np.float64
    off = 0
    for L in lengths:
        for t in range(L):
            i = off + t
            her.add({"obs_1d": g[f"{name}_in_obs"][i], "action": g[f"{name}_in_action"][i],
                     "achieved_goal": g[f"{name}_in_ag"][i].astype(gdt), "desired_goal": g[f"{name}_in_dg"][i].astype(gdt),
                     "reward": float(g[f"{name}_in_reward"][i]), "task_done": bool(g[f"{name}_in_task_done"][i]),
                     "episode_done": t == L - 1, "episode_step": t, "info": {}})
        off += L
    n = len(sink)
    assert n == 2 * lengths.sum()
    for k in ("obs_1d", "action", "achieved_goal", "desired_goal", "task_done", "episode_done", "episode_step"):
        np.testing.assert_array_equal(sink.memory[k][:n], g[f"{name}_{mode}_{k}"], err_msg=k)
    np.testing.assert_allclose(sink.memory["reward"][:n], g[f"{name}_{mode}_reward"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(sink.memory["mc_return"][:n], g[f"{name}_{mode}_mc_return"], rtol=1e-6, atol=1e-7)


def test_quantile_huber_matches_reference_function():
    g = load_golden("tqc")
    for i in range(int(g["qh_cases"])):
        q, s = g[f"qh{i}_q"], g[f"qh{i}_s"]
        loss = O.quantile_huber(q, s)
        grad = O.quantile_huber_grad(q, s)
        np.testing.assert_allclose(loss, g[f"qh{i}_loss"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(grad, g[f"qh{i}_grad"], rtol=1e-4, atol=1e-7)


def test_tqc_q_loss_matches_reference_method():
    g = load_golden("tqc")
    for i in range(int(g["ql_cases"])):
        p = f"ql{i}_"
        loss, grad, summ = O.tqc_q_loss(g[p + "q_pred"], g[p + "next_z"], g[p + "log_pi"], g[p + "reward"],
                                        g[p + "mask"], g[p + "mc_return"], float(g[p + "alpha"]), float(g[p + "gamma"]),
                                        int(g[p + "n_drop"]), bool(g[p + "ment"]), bool(g[p + "lb"]))
        np.testing.assert_allclose(loss, g[p + "loss"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(grad * g[p + "upstream"], g[p + "grad"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(summ["q_pred_mu"], g[p + "q_pred_mu"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(summ["q_pred_var"], g[p + "q_pred_var"], rtol=1e-4)
        if bool(g[p + "lb"]):
            np.testing.assert_allclose(summ["mc_constraint_violations"], g[p + "viol"], rtol=1e-6)
        # fp32 target is bit-identical to sort/slice/affine in torch
        td32 = O.tqc_td_target(g[p + "next_z"], g[p + "log_pi"], g[p + "reward"], g[p + "mask"],
                               float(g[p + "alpha"]), float(g[p + "gamma"]), int(g[p + "n_drop"]), bool(g[p + "ment"]))
        assert td32.shape[-1] == g[p + "next_z"].shape[-1] - int(g[p + "n_drop"])
    with pytest.raises(ValueError):
        O.tqc_td_target(np.zeros((1, 4)), np.zeros((1, 1)), np.zeros((1, 1)), np.ones((1, 1)), 1.0, 0.99, 0)
    assert O.n_atoms_dropped(0.08, 125) == 10


def test_tqc_torch_form_matches_numpy_form():
    import torch
    g = load_golden("tqc")
    p = "ql0_"
    t = lambda k: torch.tensor(g[p + k].astype(np.float32))
    loss = O.tqc_q_loss_torch(t("q_pred"), t("next_z"), t("log_pi"), t("reward"), t("mask"), t("mc_return"),
                              float(g[p + "alpha"]), float(g[p + "gamma"]), int(g[p + "n_drop"]))
    np.testing.assert_allclose(loss.numpy(), g[p + "loss"], rtol=1e-5, atol=1e-6)


def test_sac_min_target_loss_matches_reference_method():
    g = load_golden("tqc")
    for i in range(int(g["sac_cases"])):
        p = f"sac{i}_"
        loss, grad, summ = O.sac_min_target_loss(g[p + "q_pred"], g[p + "target_z"], g[p + "log_pi"], g[p + "reward"],
                                                 g[p + "mask"], g[p + "mc_return"], float(g[p + "alpha"]),
                                                 float(g[p + "gamma"]), bool(g[p + "ment"]), bool(g[p + "lb"]))
        np.testing.assert_allclose(loss, g[p + "loss"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(grad * g[p + "upstream"], g[p + "grad"], rtol=1e-4, atol=1e-7)
        if bool(g[p + "lb"]):
            np.testing.assert_allclose(summ["mc_constraint_violations"], g[p + "viol"], rtol=1e-6)


def test_learner_pre_post_processing():
    g = load_golden("get_losses")
    for ci in range(2):
        p = f"gl{ci}_"
        mask, contig = O.learner_preprocess(g[p + "task_done"], g[p + "episode_step"])
        np.testing.assert_array_equal(mask, g[p + "mask"])
        np.testing.assert_array_equal(contig, g[p + "is_contiguous"])
        w = O.upstream_weight(contig, int(g[p + "T"]))
        np.testing.assert_allclose(w, g[p + "dloss_dq_loss"], rtol=1e-5, atol=1e-9)
    np.testing.assert_array_equal(O.action_onehot(g["gl1_action"], 3)[:-1], g["gl1_onehot"])  # captured from curr_xp
    red = O.loss_reduce(g["red_per"], g["red_contig"], int(g["red_T"]))
    np.testing.assert_allclose(red, float(g["red_loss"]), rtol=1e-5)
    np.testing.assert_allclose(O.pohlen_transform(g["pohlen_in"]), g["pohlen_out"], rtol=1e-12)


def test_sample_time_relabel_equals_write_time_rows():
    """The sample-time definition (oracle.sample_time_relabel) reproduces, row for row, the hindsight
    rows the reference stored at write time (golden her.npz, mode=random with the same goal pick)."""
    g = load_golden("her")
    gamma = float(g["gamma"])
    for name in ("bitflip", "all_geq", "first_geq"):
        fn = O.REWARD_OPS[name]
        lengths, picks = g[f"{name}_lengths"], g[f"{name}_picks"]
        N = int(lengths.sum())
        ends = np.cumsum(lengths) - 1
        starts_ep = ends - lengths + 1
        ep_of = np.repeat(np.arange(len(lengths)), lengths)
        real_mc = O.segmented_returns(g[f"{name}_in_reward"], np.isin(np.arange(N), ends), gamma)
        cols = {"achieved_goal": g[f"{name}_in_ag"].astype(np.float64), "desired_goal": g[f"{name}_in_dg"].astype(np.float64),
                "reward": g[f"{name}_in_reward"].reshape(-1, 1), "task_done": g[f"{name}_in_task_done"].reshape(-1, 1),
                "episode_step": (np.arange(N) - starts_ep[ep_of]).reshape(-1, 1).astype(np.float32),
                "mc_return": real_mc.reshape(-1, 1)}
        starts = np.arange(N)
        out = O.sample_time_relabel(cols, starts, 1, np.ones(N, bool), (starts_ep + picks)[ep_of],
                                    starts_ep[ep_of], ends[ep_of], fn, gamma)
        # stored layout: per episode, L real rows then L hindsight rows
        hs = np.concatenate([np.arange(2 * s + L, 2 * s + 2 * L) for s, L in zip(starts_ep, lengths)])
        for k in ("desired_goal", "task_done", "episode_step"):
            np.testing.assert_array_equal(out[k][0], g[f"{name}_random_{k}"][hs].astype(np.float32), err_msg=k)
        np.testing.assert_allclose(out["reward"][0], g[f"{name}_random_reward"][hs], rtol=1e-6)
        np.testing.assert_allclose(out["mc_return"][0], g[f"{name}_random_mc_return"][hs], rtol=1e-6, atol=1e-7)


def _vmap_case(g, name):
    lengths = g[f"{name}_lengths"]
    V = int(g["V"])
    picks_deque = g[f"{name}_picks_deque"].reshape(len(lengths), V)
    offs = np.concatenate([[0], np.cumsum(lengths)])
    return lengths, V, picks_deque, offs


@pytest.mark.parametrize("name,fn", [("bitflip", O.reward_bitflip), ("all_geq", O.reward_all_geq)])
def test_vmap_write_chain_matches_reference(name, fn):
    """HindsightVmapWrite -> NStepReturnVmap -> ReplayMemory of the reference (executed with a numpy stand-in for jax.vmap,
    oracle/ref_loader.py) vs the restatement: stored virtual goals / rewards / dones / returns, bit for bit."""
    g = load_golden("her_vmap")
    lengths, V, picks_deque, offs = _vmap_case(g, name)
    for e, L in enumerate(lengths):
        sl = slice(offs[e], offs[e + 1])
        cols = {"achieved_goal": g[f"{name}_in_ag"][sl], "desired_goal": g[f"{name}_in_dg"][sl],
                "reward": g[f"{name}_in_reward"][sl], "task_done": g[f"{name}_in_task_done"][sl]}
        picks = L - 1 - picks_deque[e]  # deque index (newest first) -> chronological row
        got = O.vmap_write_episode(cols, picks, fn, gamma=float(g["gamma"]), reference_done_quirk=True)
        for tag in ("ret", "noret"):
            np.testing.assert_array_equal(got["virtual_goals"], g[f"{name}_{tag}_virtual_goals"][sl])
            np.testing.assert_array_equal(got["virtual_rewards"], g[f"{name}_{tag}_virtual_rewards"][sl])
            np.testing.assert_array_equal(got["virtual_dones"], g[f"{name}_{tag}_virtual_dones"][sl])
        np.testing.assert_array_equal(got["virtual_mc_return"], g[f"{name}_ret_virtual_mc_return"][sl])
        assert f"{name}_noret_virtual_mc_return" not in g.files
        # the stored real rows are untouched
        np.testing.assert_array_equal(g[f"{name}_ret_reward"][sl].reshape(-1), np.float32(cols["reward"]))


@pytest.mark.parametrize("name", ["bitflip", "all_geq"])
def test_vmap_read_head_matches_reference(name):
    g = load_golden("her_vmap")
    V = int(g["V"])
    stored = {k[len(name) + 5:]: g[k] for k in g.files if k.startswith(f"{name}_ret_")}
    starts = g[f"{name}_read_starts"]
    idx = np.arange(2)[:, None] + starts[None, :]
    batch = {k: v[idx] for k, v in stored.items()}
    for col in range(V + 1):
        got = O.vmap_read_select(batch, col)
        want = {k[len(f"{name}_read{col}_"):]: g[k] for k in g.files if k.startswith(f"{name}_read{col}_")}
        assert set(got) == set(want)
        for k in want:
            np.testing.assert_array_equal(np.asarray(got[k], np.float64), want[k], err_msg=k)


def test_vmap_returns_sane_mode_stops_at_virtual_terminals():
    r = np.array([[1.0], [2.0], [3.0]], np.float32)
    d = np.array([[0], [1], [0]], bool)
    np.testing.assert_allclose(O.vmap_returns(r, d, 0.5, reference_done_quirk=False).reshape(-1), [2.0, 2.0, 3.0])
    np.testing.assert_allclose(O.vmap_returns(r, d, 0.5, reference_done_quirk=True).reshape(-1), [1.0, 3.5, 3.0])
