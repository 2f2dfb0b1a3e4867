"""Pins the oracle restatements added in round 2 against outputs of the unmodified reference (tests/golden/{her_q3,pnorm,
vmap_pop,sac_bootstrap}.npz, produced by oracle/make_goldens_r2.py).  CPU only."""
import numpy as np
import pytest

from oracle import cpu_restatement as O
from conftest import load_golden


def _feed(her, g, name, gdt):
    off = 0
    for L in g[f"{name}_lengths"]:
        for t in range(L):
            i = off + t
            her.add({"obs_1d": g[f"{name}_in_obs"][i], "action": g[f"{name}_in_action"][i],
                     "achieved_goal": g[f"{name}_in_ag"][i].astype(gdt), "desired_goal": g[f"{name}_in_dg"][i].astype(gdt),
                     "reward": float(g[f"{name}_in_reward"][i]), "task_done": bool(g[f"{name}_in_task_done"][i]),
                     "episode_done": t == L - 1, "episode_step": t, "info": {}})
        off += L


@pytest.mark.parametrize("mode", ["final", "random"])
def test_her_over_short_nstep_row_stream(mode):
    """quirk Q3 under hindsight (her.py:36-46 over nstep_return.py:33-34,50-57): dup, rows, dup', hindsight rows."""
    g = load_golden("her_q3")
    n_step, gamma = int(g["n_step"]), float(g["gamma"])
    lengths = g["bitflip_lengths"]
    sink = O.RingOracle(4096, 8, 2)
    it = iter(list(g["bitflip_picks"]))
    her = O.HindsightOracle(O.NStepOracle(sink, n_step, gamma), O.reward_bitflip, mode=mode, goal_picker=lambda L: next(it))
    _feed(her, g, "bitflip", np.int64)
    n = len(sink)
    assert n == 2 * lengths.sum() + 2 * int((lengths > n_step).sum()) == len(g[f"bitflip_{mode}_reward"])
    for k in ("obs_1d", "action", "achieved_goal", "desired_goal", "task_done", "episode_done", "episode_step"):
        np.testing.assert_array_equal(sink.memory[k][:n], g[f"bitflip_{mode}_{k}"], err_msg=k)
    np.testing.assert_allclose(sink.memory["reward"][:n], g[f"bitflip_{mode}_reward"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(sink.memory["mc_return"][:n], g[f"bitflip_{mode}_mc_return"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("mode", ["final", "random"])
def test_parking_functor_row_stream(mode):
    """Env/eleurent_parking.py:42-55 through write-time hindsight: oracle.make_reward_weighted_pnorm vs the reference rows."""
    g = load_golden("pnorm")
    fn = O.make_reward_weighted_pnorm(g["weights"], float(g["success"]), float(g["p"]))
    sink = O.RingOracle(4096, 8, 2)
    it = iter(list(g["parking_picks"]))
    her = O.HindsightOracle(O.NStepOracle(sink, 1000, float(g["gamma"])), fn, mode=mode, goal_picker=lambda L: next(it))
    _feed(her, g, "parking", np.float64)
    n = len(sink)
    assert n == 2 * g["parking_lengths"].sum()
    for k in ("achieved_goal", "desired_goal", "task_done", "episode_done", "episode_step"):
        np.testing.assert_array_equal(sink.memory[k][:n], g[f"parking_{mode}_{k}"], err_msg=k)
    np.testing.assert_allclose(sink.memory["reward"][:n], g[f"parking_{mode}_reward"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(sink.memory["mc_return"][:n], g[f"parking_{mode}_mc_return"], rtol=1e-5, atol=1e-6)


def test_parking_functor_vmap_columns():
    g = load_golden("pnorm")
    fn = O.make_reward_weighted_pnorm(g["weights"], float(g["success"]), float(g["p"]))
    V = int(g["V"])
    lengths = g["vmap_lengths"]
    picks_deque = g["vmap_picks_deque"].reshape(len(lengths), V)
    offs = np.concatenate([[0], np.cumsum(lengths)])
    for e, L in enumerate(lengths):
        sl = slice(offs[e], offs[e + 1])
        cols = {"achieved_goal": g["vmap_in_ag"][sl], "desired_goal": g["vmap_in_dg"][sl], "reward": g["vmap_in_reward"][sl],
                "task_done": g["vmap_in_task_done"][sl]}
        got = O.vmap_write_episode(cols, L - 1 - picks_deque[e], fn, gamma=float(g["gamma"]), reference_done_quirk=True)
        np.testing.assert_array_equal(got["virtual_goals"], g["vmap_virtual_goals"][sl])
        np.testing.assert_array_equal(got["virtual_dones"], g["vmap_virtual_dones"][sl])
        np.testing.assert_allclose(got["virtual_rewards"], g["vmap_virtual_rewards"][sl], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(got["virtual_mc_return"], g["vmap_virtual_mc_return"][sl], rtol=1e-5, atol=1e-6)


def test_vmap_pop_duplicate_rows():
    """nstep_return_vmap.py:50-57: the ring holds, per episode longer than n_step, the oldest row once more IN FRONT of the episode,
    with the per-column return truncated after n_step rows."""
    g = load_golden("vmap_pop")
    V, n_step, gamma = int(g["V"]), int(g["n_step"]), float(g["gamma"])
    lengths = g["bitflip_lengths"]
    picks_deque = g["bitflip_picks_deque"].reshape(len(lengths), V)
    offs = np.concatenate([[0], np.cumsum(lengths)])
    row = 0
    for e, L in enumerate(lengths):
        sl = slice(offs[e], offs[e + 1])
        cols = {"achieved_goal": g["bitflip_in_ag"][sl], "desired_goal": g["bitflip_in_dg"][sl], "reward": g["bitflip_in_reward"][sl],
                "task_done": g["bitflip_in_task_done"][sl]}
        got = O.vmap_write_episode(cols, L - 1 - picks_deque[e], O.reward_bitflip, gamma=gamma, reference_done_quirk=True)
        if L > n_step:
            np.testing.assert_array_equal(g["stored_virtual_rewards"][row], got["virtual_rewards"][0])
            np.testing.assert_array_equal(g["stored_episode_step"][row].reshape(-1), [0.0])
            np.testing.assert_array_equal(g["stored_virtual_mc_return"][row],
                                          O.vmap_pop_returns(got["virtual_rewards"], got["virtual_dones"], n_step, gamma, True))
            row += 1
        np.testing.assert_array_equal(g["stored_virtual_goals"][row:row + L], got["virtual_goals"])
        np.testing.assert_array_equal(g["stored_virtual_mc_return"][row:row + L], got["virtual_mc_return"])
        row += L
    assert row == int(g["n_rows"])


def test_sac_bootstrap_bound_matches_reference():
    g = load_golden("sac_bootstrap")
    for i in range(int(g["cases"])):
        p = f"sb{i}_"
        bound = O.sac_bootstrap_bound(g[p + "q_pred"], g[p + "target_z"], g[p + "log_pi"], g[p + "reward"], g[p + "mask"],
                                      float(g[p + "alpha"]), float(g[p + "gamma"]), bool(g[p + "ment"]))
        np.testing.assert_allclose(bound, g[p + "bound"], rtol=1e-5, atol=1e-6)
        loss, grad, _ = O.sac_min_target_loss(g[p + "q_pred"], g[p + "target_z"], g[p + "log_pi"], g[p + "reward"], g[p + "mask"],
                                              g[p + "mc_return"], float(g[p + "alpha"]), float(g[p + "gamma"]), bool(g[p + "ment"]), True)
        np.testing.assert_allclose(loss, g[p + "loss"], rtol=1e-5, atol=1e-6)
        want = grad * g[p + "up_q"]
        want[0] -= (bound > 0) * g[p + "up_b"] * g[p + "mask"].astype(np.float64).prod(0)
        np.testing.assert_allclose(want, g[p + "grad"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(float((bound != 0).mean()), float(g[p + "viol"]), rtol=1e-6)
