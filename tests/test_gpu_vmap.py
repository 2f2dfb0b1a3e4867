"""GPU parity of the "vmap" hindsight variant (her_mode="vmap"): HindsightVmapWrite -> NStepReturnVmap -> ring and the
HindsightVmapRead head, vs goldens produced by running the reference's own code (tests/golden/her_vmap.npz) and vs the oracle.
Bit-exact for goals, rewards, dones and (reference arithmetic, quirk Q7) returns."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import cpu_restatement as O

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


def _feed(fdql, g, name, op, with_returns, quirk, monkeypatch):
    from fastdeepqlearning_b200 import Replay
    from fastdeepqlearning_b200.Replay import wrappers as W
    V = int(g["V"])
    lengths = g[f"{name}_lengths"]
    picks = iter(g[f"{name}_picks_deque"].reshape(len(lengths), V))
    shard = Replay.AsyncReplayMemory(4096, 8, 2)
    inner = W.NStepReturnVmap(shard, 1000, float(g["gamma"]), reference_done_quirk=quirk) if with_returns else shard
    her = W.HindsightVmapWrite(inner, op, num_virtual_goals=V)
    monkeypatch.setattr(np.random, "randint", lambda low, high=None, size=None: np.asarray(next(picks)))
    off = 0
    for L in lengths:
        for t in range(L):
            i = off + t
            her.add({"obs_1d": g[f"{name}_in_obs"][i], "action": g[f"{name}_in_action"][i],
                     "achieved_goal": g[f"{name}_in_ag"][i].astype(np.float32), "desired_goal": g[f"{name}_in_dg"][i].astype(np.float32),
                     "reward": float(g[f"{name}_in_reward"][i]), "task_done": bool(g[f"{name}_in_task_done"][i]),
                     "episode_done": t == L - 1, "episode_step": t, "info": {}})
        off += L
    monkeypatch.undo()
    return shard


@pytest.mark.parametrize("name", ["bitflip", "all_geq"])
@pytest.mark.parametrize("with_returns", [True, False])
def test_vmap_write_chain_golden(fdql, name, with_returns, monkeypatch):
    g = load_golden("her_vmap")
    op = fdql.RewardOp.bitflip() if name == "bitflip" else fdql.RewardOp.all_geq()
    shard = _feed(fdql, g, name, op, with_returns, True, monkeypatch)
    tag = "ret" if with_returns else "noret"
    n = int(g[f"{name}_lengths"].sum())
    assert len(shard) == n
    mem = shard.replay.memory
    want_keys = [k[len(f"{name}_{tag}_"):] for k in g.files if k.startswith(f"{name}_{tag}_")]
    assert set(want_keys) == set(mem)
    for k in want_keys:
        want = g[f"{name}_{tag}_{k}"]
        np.testing.assert_array_equal(npy(mem[k])[:n].reshape(n, -1), want.reshape(n, -1).astype(np.float32), err_msg=k)


@pytest.mark.parametrize("name", ["bitflip", "all_geq"])
def test_vmap_read_head_golden(fdql, name, monkeypatch):
    from fastdeepqlearning_b200.Replay import wrappers as W
    g = load_golden("her_vmap")
    op = fdql.RewardOp.bitflip() if name == "bitflip" else fdql.RewardOp.all_geq()
    shard = _feed(fdql, g, name, op, True, True, monkeypatch)
    read = W.HindsightVmapRead(shard, aux=True)
    starts = g[f"{name}_read_starts"]
    for col in range(int(g["V"]) + 1):
        got = read.temporal_sample(column=col, starts=starts)
        want = {k[len(f"{name}_read{col}_"):]: g[k] for k in g.files if k.startswith(f"{name}_read{col}_")}
        assert set(want) <= set(got) and not any(k.startswith("virtual_") for k in got)
        for k, v in want.items():
            np.testing.assert_array_equal(npy(got[k]), v.astype(np.float32), err_msg=f"{k} col {col}")
        mask, contig = O.learner_preprocess(want["task_done"], want["episode_step"])
        np.testing.assert_array_equal(npy(got["mask"]), mask.astype(np.float32))
        np.testing.assert_array_equal(npy(got["is_contiguous"]), contig.astype(np.float32))
        np.testing.assert_allclose(npy(got["loss_weight"]), O.upstream_weight(contig, 2), rtol=1e-6, atol=1e-12)
    # the column is drawn with random.randint(0, V) inclusive when not injected (her_vmap.py:107)
    seen = {float(npy(read.temporal_sample(starts=starts)["reward"]).sum()) for _ in range(40)}
    assert len(seen) > 1


@pytest.mark.parametrize("quirk", [True, False])
def test_vmap_vs_oracle_large(fdql, quirk):
    """V = 32 virtual goals of width 16 (the reference's default num_virtual_goals), episodes up to 150 rows, both return modes."""
    import torch
    from fastdeepqlearning_b200 import Replay
    from fastdeepqlearning_b200.Replay import wrappers as W
    rng = np.random.default_rng(5)
    V, G, gamma = 32, 16, 0.97
    shard = Replay.AsyncReplayMemory(6000, 64, 3)
    inner = W.NStepReturnVmap(shard, 1000, gamma, reference_done_quirk=quirk)
    her = W.HindsightVmapWrite(inner, fdql.RewardOp.bitflip(), num_virtual_goals=V)
    want = {k: [] for k in ("virtual_goals", "virtual_rewards", "virtual_dones", "virtual_mc_return")}
    for e in range(25):
        L = int(rng.integers(1, 151))
        ag = rng.integers(0, 2, (L, G)).astype(np.float32)
        dg = np.tile(rng.integers(0, 2, G).astype(np.float32), (L, 1))
        hit = rng.random(L) < 0.2
        ag[hit] = dg[hit]
        ag[rng.random(L) < 0.3] = ag[0]  # revisited goals: virtual goals that match several rows
        rew, done = O.reward_bitflip(ag, dg)
        cols = {"obs_1d": rng.standard_normal((L, 7)).astype(np.float32), "achieved_goal": ag, "desired_goal": dg,
                "reward": (rew + (rng.random(L) < 0.1) * 0.25).astype(np.float32).reshape(-1, 1),
                "task_done": done.astype(np.float32).reshape(-1, 1),
                "episode_done": (np.arange(L) == L - 1).astype(np.float32).reshape(-1, 1),
                "episode_step": np.arange(L, dtype=np.float32).reshape(-1, 1)}
        picks = rng.integers(0, L, V)
        cols_dev = dict(cols)
        cols_dev["virtual_goals"] = torch.zeros((L, (V + 1) * G), device="cuda")
        cols_dev["virtual_rewards"] = torch.zeros((L, V + 1), device="cuda")
        cols_dev["virtual_dones"] = torch.zeros((L, V + 1), device="cuda")
        inner.add_vmap_rows(cols_dev, L, picks)
        w = O.vmap_write_episode(cols, picks, O.reward_bitflip, gamma=gamma, reference_done_quirk=quirk)
        for k in want:
            want[k].append(np.asarray(w[k], np.float32).reshape(L, -1))
    mem = shard.replay.memory
    for k, v in want.items():
        v = np.concatenate(v)
        np.testing.assert_array_equal(npy(mem[k])[:len(v)], v, err_msg=k)
