"""GPU parity through the mirror classes, used exactly like the reference's wrappers: rows are added one dict at a time to
HindsightNStepReplay(NStepReturn(AsyncReplayMemory)) (the composition of franQ/Replay/__init__.py:19-36) and the ring
must hold the row stream the unmodified reference stored (tests/golden/her.npz, nstep.npz)."""
import random
import types

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name", ["bitflip", "all_geq", "first_geq"])
@pytest.mark.parametrize("mode", ["final", "random"])
def test_reference_row_stream_through_make(fdql, name, mode, monkeypatch):
    from fastdeepqlearning_b200 import Replay
    g = load_golden("her")
    conf = types.SimpleNamespace(replay_size=4096, batch_size=8, temporal_len=2, num_instances=1, use_nStep_lowerbounds=True,
                                 nStep_return_steps=1000, gamma=float(g["gamma"]), use_squashed_rewards=False, use_HER=True,
                                 her_mode=mode, training_device="cuda:0")
    read_heads, write_heads = Replay.make(conf, compute_reward=fdql.RewardOp.coerce(name))
    head = write_heads[0]
    assert type(head).__name__ == "HindsightNStepReplay" and type(head.replay_buffer).__name__ == "NStepReturn"
    lengths, picks = g[f"{name}_lengths"], list(g[f"{name}_picks"])
    it = iter(picks)
    # the reference draws random.choice(deque) over the newest-first deque; inject the golden's chronological picks
    monkeypatch.setattr(random, "choice", lambda seq: seq[len(seq) - 1 - next(it)])
    gdt = np.int64 if name == "bitflip" else np.float64
    off = 0
    for L in lengths:
        for t in range(L):
            i = off + t
            head.add({"obs_1d": g[f"{name}_in_obs"][i], "action": g[f"{name}_in_action"][i],
                      "achieved_goal": g[f"{name}_in_ag"][i].astype(gdt), "desired_goal": g[f"{name}_in_dg"][i].astype(gdt),
                      "reward": float(g[f"{name}_in_reward"][i]), "task_done": bool(g[f"{name}_in_task_done"][i]),
                      "episode_done": t == L - 1, "episode_step": t, "info": {}})
        off += L
    ring = read_heads[0]
    n = len(ring)
    assert n == 2 * lengths.sum()
    mem = {k: npy(v)[:n] for k, v in ring.memory.items()}
    assert "info" not in mem
    for k in ("obs_1d", "action", "achieved_goal", "desired_goal", "task_done", "episode_done", "episode_step", "reward", "mc_return"):
        np.testing.assert_array_equal(mem[k], g[f"{name}_{mode}_{k}"].astype(np.float32), err_msg=k)
    xp = ring.temporal_sample()
    assert tuple(xp["obs_1d"].shape) == (2, 8, 3)


def test_nstep_wrapper_known_answer(fdql):
    """reference tests/test_replays.py:16-33 through the mirror classes: ReplayMemory(1001,128,1) + NStepReturn + sample()."""
    from fastdeepqlearning_b200 import Replay
    n_step, discount = 1000, 0.99
    replay = Replay.wrappers.NStepReturn(Replay.ReplayMemory(n_step + 1, 128, 1), n_step, discount)
    for i in range(n_step):
        replay.add({"reward": 1 if i == (n_step - 1) else 0, "episode_done": i == (n_step - 1), "episode_step": i})
    xp = replay.sample()
    assert np.allclose(npy(xp["mc_return"]), discount ** (n_step - 1 - npy(xp["episode_step"])))


def test_nstep_quirk_q3_duplicate_row(fdql):
    """quirk Q3 (nstep_return.py:33-34,50-57): with episode length > n_step the oldest row is stored twice, first with the
    n-step-truncated return; golden = the reference's own row stream for n=3, L=6."""
    from fastdeepqlearning_b200 import Replay
    g = load_golden("nstep")
    ring = Replay.ReplayMemory(100, 4, 2)
    w = Replay.wrappers.NStepReturn(ring, 3, 0.9)
    rs = g["q3_rewards"]
    for t, r in enumerate(rs):
        w.add({"reward": float(r), "episode_done": t == len(rs) - 1, "episode_step": t})
    assert len(ring) == int(g["q3_n_rows"]) == len(rs) + 1
    mem = ring.memory
    np.testing.assert_array_equal(npy(mem["mc_return"])[:len(ring)], g["q3_mc_return"])
    np.testing.assert_array_equal(npy(mem["episode_step"])[:len(ring)], g["q3_step"])


def test_sample_time_future_read_head(fdql):
    """her_mode='future' (BASELINE.json configs[2]): rows stored once, relabelled at sample time by the read head."""
    from fastdeepqlearning_b200 import Replay
    rng = np.random.default_rng(0)
    conf = types.SimpleNamespace(replay_size=5000, batch_size=64, temporal_len=2, num_instances=2, use_nStep_lowerbounds=True,
                                 nStep_return_steps=1000, gamma=0.98, use_squashed_rewards=False, use_HER=True, her_mode="future",
                                 her_relabel_prob=0.8, training_device="cuda:0")
    read_heads, write_heads = Replay.make(conf, compute_reward=fdql.RewardOp.bitflip())
    assert len(read_heads) == len(write_heads) == 2
    for head in write_heads:
        for ep in range(30):
            L = int(rng.integers(2, 40))
            dg = rng.integers(0, 2, 8)
            for t in range(L):
                ag = rng.integers(0, 2, 8)
                hit = bool((ag == dg).all())
                head.add({"obs_1d": rng.standard_normal(5).astype(np.float32), "action": rng.standard_normal(2).astype(np.float32),
                          "achieved_goal": ag, "desired_goal": dg, "reward": 0.0 if hit else -1.0, "task_done": hit,
                          "episode_done": t == L - 1, "episode_step": t, "info": {"is_success": hit}})  # runner.py:185-186 keeps info
    assert "info" not in write_heads[0].keys
    xp = read_heads[0].temporal_sample()
    assert {"mask", "is_contiguous", "loss_weight"} <= set(xp)
    hit = (xp["achieved_goal"] == xp["desired_goal"]).all(-1, keepdim=True).float()
    assert bool((xp["task_done"] == hit).all()) and bool((xp["reward"] == hit - 1).all())
    assert bool((xp["mask"] == 1 - xp["task_done"]).all())
