"""GPU parity added in round 2, every case against outputs of the unmodified reference (tests/golden/{her_q3,pnorm,vmap_pop,
learner_step,sac_bootstrap}.npz, oracle/make_goldens_r2.py) or against the pinned oracle:
  * hindsight over an NStepReturn shorter than the episode (quirk Q3 under HER), through Replay.make, one dict at a time;
  * the parking reward functor (weighted p-norm) at write time, in the vmap chain and at sample time;
  * vmap episodes longer than nStep_return_steps (NStepReturnVmap._pop);
  * one whole learner update (loss, every gradient, Adam step, target update) on recorded weights, batch and policy noise;
  * the minibatch n-step bootstrap bound of SoftActorCritic.q_loss;
  * link records on episodes of 1000 and of more than 32767 rows; small rings; partly overwritten episodes."""
import random
import types

import numpy as np
import pytest

from conftest import load_golden
from oracle import cpu_restatement as O

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


def _feed(head, g, name, gdt):
    off = 0
    for L in g[f"{name}_lengths"]:
        for t in range(L):
            i = off + t
            head.add({"obs_1d": g[f"{name}_in_obs"][i], "action": g[f"{name}_in_action"][i],
                      "achieved_goal": g[f"{name}_in_ag"][i].astype(gdt), "desired_goal": g[f"{name}_in_dg"][i].astype(gdt),
                      "reward": float(g[f"{name}_in_reward"][i]), "task_done": bool(g[f"{name}_in_task_done"][i]),
                      "episode_done": t == L - 1, "episode_step": t, "info": {}})
        off += L


def _make(fdql, mode, op, n_step, gamma, replay_size=4096):
    from fastdeepqlearning_b200 import Replay
    conf = types.SimpleNamespace(replay_size=replay_size, batch_size=8, temporal_len=2, num_instances=1, use_nStep_lowerbounds=True,
                                 nStep_return_steps=n_step, gamma=gamma, use_squashed_rewards=False, use_HER=True, her_mode=mode,
                                 training_device="cuda:0")
    return Replay.make(conf, compute_reward=op)


# ------------------------------------------------------------------------------------------------ quirk Q3 under hindsight
@pytest.mark.parametrize("mode", ["final", "random"])
def test_her_over_short_nstep_reference_row_stream(fdql, mode, monkeypatch):
    """her.py:36-46 over nstep_return.py:33-34,50-57 with n_step = 4 < episode length: the ring must hold, per long episode,
    dup, the real rows, dup', the hindsight rows -- copied from and counted from the episode's own first row, not the duplicate's."""
    g = load_golden("her_q3")
    read, write = _make(fdql, mode, fdql.RewardOp.bitflip(), int(g["n_step"]), float(g["gamma"]))
    it = iter(list(g["bitflip_picks"]))
    monkeypatch.setattr(random, "choice", lambda seq: seq[len(seq) - 1 - next(it)])
    _feed(write[0], g, "bitflip", np.int64)
    ring = read[0]
    n = len(ring)
    lengths = g["bitflip_lengths"]
    assert n == 2 * lengths.sum() + 2 * int((lengths > int(g["n_step"])).sum())
    mem = {k: npy(v)[:n] for k, v in ring.memory.items()}
    for k in ("obs_1d", "action", "achieved_goal", "desired_goal", "task_done", "episode_done", "episode_step", "reward", "mc_return"):
        np.testing.assert_array_equal(mem[k], g[f"bitflip_{mode}_{k}"].astype(np.float32), err_msg=k)


# ------------------------------------------------------------------------------------------------ parking functor
def _parking_op(fdql, g):
    return fdql.RewardOp.weighted_pnorm(list(g["weights"]), float(g["success"]), float(g["p"]))


@pytest.mark.parametrize("mode", ["final", "random"])
def test_parking_functor_write_time_reference_rows(fdql, mode, monkeypatch):
    """Env/eleurent_parking.py:42-55 as FDQL_REWARD_WEIGHTED_PNORM through fdql_her_flush_episodes vs the reference's stored rows:
    goals, dones and re-based steps bit for bit, rewards and returns to 1e-6 (fp64 functor, fp32 store on both sides)."""
    g = load_golden("pnorm")
    read, write = _make(fdql, mode, _parking_op(fdql, g), 1000, float(g["gamma"]))
    it = iter(list(g["parking_picks"]))
    monkeypatch.setattr(random, "choice", lambda seq: seq[len(seq) - 1 - next(it)])
    _feed(write[0], g, "parking", np.float64)
    ring = read[0]
    n = len(ring)
    assert n == 2 * g["parking_lengths"].sum()
    mem = {k: npy(v)[:n] for k, v in ring.memory.items()}
    for k in ("obs_1d", "action", "achieved_goal", "desired_goal", "task_done", "episode_done", "episode_step"):
        np.testing.assert_array_equal(mem[k], g[f"parking_{mode}_{k}"].astype(np.float32), err_msg=k)
    assert mem["task_done"].sum() > 0
    np.testing.assert_allclose(mem["reward"], g[f"parking_{mode}_reward"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(mem["mc_return"], g[f"parking_{mode}_mc_return"], rtol=1e-5, atol=1e-6)


def test_parking_functor_vmap_chain_reference_rows(fdql, monkeypatch):
    from fastdeepqlearning_b200 import Replay
    from fastdeepqlearning_b200.Replay import wrappers as W
    g = load_golden("pnorm")
    V = int(g["V"])
    lengths = g["vmap_lengths"]
    picks = iter(g["vmap_picks_deque"].reshape(len(lengths), V))
    shard = Replay.AsyncReplayMemory(4096, 8, 2)
    her = W.HindsightVmapWrite(W.NStepReturnVmap(shard, 1000, float(g["gamma"]), reference_done_quirk=True), _parking_op(fdql, g),
                               num_virtual_goals=V)
    monkeypatch.setattr(np.random, "randint", lambda low, high=None, size=None: np.asarray(next(picks)))
    _feed(her, g, "vmap", np.float32)
    monkeypatch.undo()
    n = int(lengths.sum())
    assert len(shard) == n
    mem = {k: npy(v)[:n] for k, v in shard.replay.memory.items()}
    np.testing.assert_array_equal(mem["virtual_goals"].reshape(n, -1), g["vmap_virtual_goals"].reshape(n, -1).astype(np.float32))
    np.testing.assert_array_equal(mem["virtual_dones"], g["vmap_virtual_dones"].astype(np.float32))
    np.testing.assert_allclose(mem["virtual_rewards"], g["vmap_virtual_rewards"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(mem["virtual_mc_return"], g["vmap_virtual_mc_return"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("T", [1, 3])
def test_parking_functor_sample_time_vs_oracle(fdql, T):
    """Sample-time relabelling with a non-equality functor (the full-vector relabel path) vs oracle.sample_time_relabel."""
    from fastdeepqlearning_b200 import Replay
    g = load_golden("pnorm")
    fn = O.make_reward_weighted_pnorm(g["weights"], float(g["success"]), float(g["p"]))
    gamma = float(g["gamma"])
    lengths = g["parking_lengths"]
    N = int(lengths.sum())
    ends = np.cumsum(lengths) - 1
    starts_ep = ends - lengths + 1
    ep_of = np.repeat(np.arange(len(lengths)), lengths)
    real_mc = O.segmented_returns(g["parking_in_reward"], np.isin(np.arange(N), ends), gamma)
    cols = {"obs_1d": g["parking_in_obs"], "achieved_goal": g["parking_in_ag"].astype(np.float32),
            "desired_goal": g["parking_in_dg"].astype(np.float32), "reward": g["parking_in_reward"].reshape(-1, 1).astype(np.float32),
            "task_done": g["parking_in_task_done"].reshape(-1, 1).astype(np.float32),
            "episode_done": np.isin(np.arange(N), ends).reshape(-1, 1).astype(np.float32),
            "episode_step": (np.arange(N) - starts_ep[ep_of]).reshape(-1, 1).astype(np.float32), "mc_return": real_mc.reshape(-1, 1)}
    ring = Replay.ReplayMemory(N + 3, 16, T)
    ring.set_reward_op(_parking_op(fdql, g), gamma)
    ring.add_rows(cols, episode_lengths=lengths)
    rng = np.random.default_rng(3)
    B = 300
    starts = rng.integers(0, N - T, B)
    flags = rng.random(B) < 0.8
    goal_rows = np.array([rng.integers(s, ends[ep_of[s]] + 1) for s in starts])
    want = O.sample_time_relabel(cols, starts, T, flags, goal_rows, starts_ep[ep_of], ends[ep_of], fn, gamma)
    got = ring.temporal_sample(starts=starts, flags=flags.astype(np.uint8), goal_rows=goal_rows, exact_episode_step=True, length=N)
    for k in ("desired_goal", "task_done", "episode_step", "achieved_goal", "obs_1d"):
        np.testing.assert_array_equal(npy(got[k]), want[k], err_msg=k)
    assert want["task_done"].sum() > 0
    np.testing.assert_allclose(npy(got["reward"]), want["reward"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(npy(got["mc_return"]), want["mc_return"], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------ vmap _pop
def test_vmap_pop_duplicate_reference_rows(fdql, monkeypatch):
    """nstep_return_vmap.py:50-57 through HindsightVmapWrite -> NStepReturnVmap(n_step = 3): bit for bit (reference arithmetic, Q7)."""
    from fastdeepqlearning_b200 import Replay
    from fastdeepqlearning_b200.Replay import wrappers as W
    g = load_golden("vmap_pop")
    V, n_step, gamma = int(g["V"]), int(g["n_step"]), float(g["gamma"])
    lengths = g["bitflip_lengths"]
    picks = iter(g["bitflip_picks_deque"].reshape(len(lengths), V))
    shard = Replay.AsyncReplayMemory(4096, 8, 2)
    her = W.HindsightVmapWrite(W.NStepReturnVmap(shard, n_step, gamma, reference_done_quirk=True), fdql.RewardOp.bitflip(),
                               num_virtual_goals=V)
    monkeypatch.setattr(np.random, "randint", lambda low, high=None, size=None: np.asarray(next(picks)))
    _feed(her, g, "bitflip", np.float32)
    monkeypatch.undo()
    n = int(g["n_rows"])
    assert len(shard) == n == lengths.sum() + int((lengths > n_step).sum())
    mem = shard.replay.memory
    want_keys = [k[len("stored_"):] for k in g.files if k.startswith("stored_")]
    assert set(want_keys) == set(mem)
    for k in want_keys:
        np.testing.assert_array_equal(npy(mem[k])[:n].reshape(n, -1), g[f"stored_{k}"].reshape(n, -1).astype(np.float32), err_msg=k)


def test_vmap_pop_sane_mode_vs_oracle(fdql):
    """Default (1 - done) recurrence with episodes longer than n_step, rows that already carry the virtual columns (add())."""
    from fastdeepqlearning_b200 import Replay
    from fastdeepqlearning_b200.Replay import wrappers as W
    rng = np.random.default_rng(9)
    V1, n_step, gamma = 4, 5, 0.93
    shard = Replay.AsyncReplayMemory(512, 8, 2)
    head = W.NStepReturnVmap(shard, n_step, gamma)
    want = []
    for L in (3, 9, 5, 12):
        vr = rng.standard_normal((L, V1)).astype(np.float32)
        vd = rng.random((L, V1)) < 0.25
        for t in range(L):
            head.add({"x": float(t), "virtual_goals": np.zeros((V1, 2), np.float32), "virtual_rewards": vr[t],
                      "virtual_dones": vd[t].astype(np.float32), "episode_done": t == L - 1})
        full = O.vmap_returns(vr, vd, gamma, reference_done_quirk=False)
        if L > n_step:
            want.append(O.vmap_pop_returns(vr, vd, n_step, gamma, reference_done_quirk=False)[None])
        want.append(full)
    want = np.concatenate(want)
    np.testing.assert_array_equal(npy(shard.replay.memory["virtual_mc_return"])[:len(want)], want)


# ------------------------------------------------------------------------------------------------ one whole learner update
def _strip(g, prefix):
    return {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}


def _mlp_grads(mlp, prefix):
    out = {}
    for i, lin in enumerate(mlp.hidden):
        out[f"{prefix}feature_extractor.{i}.0.weight"], out[f"{prefix}feature_extractor.{i}.0.bias"] = lin.weight.grad, lin.bias.grad
    out[f"{prefix}head.weight"], out[f"{prefix}head.bias"] = mlp.head.weight.grad, mlp.head.bias.grad
    return out


def _ensemble_tensors(critic, prefix, grad):
    """per-member tensors of a critic ensemble under the reference's parameter names (values or gradients)"""
    from fastdeepqlearning_b200.Agent.components import models
    pick = (lambda p: p.grad) if grad else (lambda p: p.detach())
    out = {}
    if isinstance(critic, models.BatchedMLPEnsemble):
        for e in range(critic.E):
            for l in range(len(critic.hidden_w)):
                out[f"{prefix}nets.{e}.feature_extractor.{l}.0.weight"] = pick(critic.hidden_w[l])[e].t()
                out[f"{prefix}nets.{e}.feature_extractor.{l}.0.bias"] = pick(critic.hidden_b[l])[e, 0]
            out[f"{prefix}nets.{e}.head.weight"], out[f"{prefix}nets.{e}.head.bias"] = pick(critic.head_w)[e].t(), pick(critic.head_b)[e, 0]
    else:
        for e, net in enumerate(critic.nets):
            for i, lin in enumerate(net.hidden):
                out[f"{prefix}nets.{e}.feature_extractor.{i}.0.weight"], out[f"{prefix}nets.{e}.feature_extractor.{i}.0.bias"] = \
                    pick(lin.weight), pick(lin.bias)
            out[f"{prefix}nets.{e}.head.weight"], out[f"{prefix}nets.{e}.head.bias"] = pick(net.head.weight), pick(net.head.bias)
    return out


def _mlp_values(mlp, prefix):
    return {k: v for k, v in mlp.reference_state_dict(prefix).items()}


@pytest.mark.parametrize("case", [0, 1])
@pytest.mark.parametrize("batched", [True, False])
@pytest.mark.parametrize("fold", [False, True])
def test_learner_update_matches_reference(fdql, case, batched, fold, monkeypatch):
    """deepQlearning.py:105-127,198-249: the reference's weights, one injected [T, B] batch and its policy noise go into the mirror;
    loss, every gradient, the Adam step and the target update must come out the same (<= 1e-5 of each tensor's scale).
    `fold`: the loss-reduce weights enter the loss kernel as grad_scale (what the gather kernel's aux output feeds)."""
    import torch
    from fastdeepqlearning_b200 import Agent
    from fastdeepqlearning_b200.Agent.components import models
    g = load_golden("learner_step")
    tag = f"ls{case}"
    T, B, C, Q = int(g[f"{tag}_T"]), int(g[f"{tag}_B"]), int(g[f"{tag}_C"]), int(g[f"{tag}_Q"])
    latent = int(g[f"{tag}_latent"])
    conf = Agent.LearnerConf(training_device="cuda:0", obs_space={"obs_1d": 6, "achieved_goal": 3, "desired_goal": 3},
                             action_space=types.SimpleNamespace(shape=(2,)), num_critics=C, num_q_predictions=Q,
                             top_quantiles_to_drop=float(g[f"{tag}_drop"]), batch_size=B, temporal_len=T, pi_hidden_dims=(16,),
                             critic_hidden_dims=(16, 16), init_log_alpha=-0.3, learning_rate=float(g[f"{tag}_lr"]), tau=float(g[f"{tag}_tau"]),
                             gamma=float(g[f"{tag}_gamma"]), use_distributional_sac=bool(g[f"{tag}_distributional"]), batched_critics=batched,
                             fold_loss_reduce=fold)
    enc = models.FeedForwardEncoder(12, latent, hidden_features=10, obs_1d_hidden_dims=(8,), joint_hidden_dims=(8,)).to("cuda")
    learner = Agent.Learner(conf, encoder=enc, state_dim=latent)
    ac = learner.actor_critic
    w0 = _strip(g, f"{tag}_w0_")
    enc.load_reference_state_dict(w0, "encoder.")
    ac.actor.load_reference_state_dict(w0, "actor.")
    ac.actor_target.load_reference_state_dict(w0, "actor_target.")
    ac.critic.load_reference_state_dict(w0, "critic.")
    ac.critic_target.load_reference_state_dict(w0, "critic_target.")
    with torch.no_grad():
        ac.log_alpha.copy_(torch.as_tensor(w0["log_alpha"]))
        ac.curr_alpha.fill_(float(np.exp(w0["log_alpha"])))
    xp = {k: torch.as_tensor(v, dtype=torch.float32, device="cuda") for k, v in _strip(g, f"{tag}_xp_").items()}
    if fold:  # what the gather kernel emits with FDQL_OPT_EMIT_LEARNER_AUX (checked against the oracle in test_gpu_replay.py)
        mask, contig = O.learner_preprocess(npy(xp["task_done"]), npy(xp["episode_step"]))
        xp["mask"] = torch.as_tensor(mask.astype(np.float32), device="cuda")
        xp["is_contiguous"] = torch.as_tensor(contig.astype(np.float32), device="cuda")
        xp["loss_weight"] = torch.as_tensor(O.upstream_weight(contig, T), device="cuda")
    eps = [torch.as_tensor(g[f"{tag}_eps{i}"], device="cuda") for i in range(2)]  # actor_target(next state), actor(curr state)
    tape = iter(eps)
    monkeypatch.setattr(torch, "randn_like", lambda t, **kw: next(tape))
    learner.optimizer.zero_grad(set_to_none=True)
    loss = learner.get_losses(xp)
    monkeypatch.undo()
    loss.backward()
    np.testing.assert_allclose(float(loss), float(g[f"{tag}_loss"]), rtol=1e-5)
    got = {}
    got.update(_mlp_grads(enc.obs_1d, "encoder.visible_layer_encoders.obs_1d."))
    got.update(_mlp_grads(enc.joiner, "encoder.joiner."))
    got.update(_mlp_grads(ac.actor, "actor_critic.actor."))
    got.update(_ensemble_tensors(ac.critic, "actor_critic.critic.", grad=True))
    got["actor_critic.log_alpha"] = ac.log_alpha.grad
    want = _strip(g, f"{tag}_grad_")
    assert set(want) == set(got)
    for k, w in want.items():
        scale = max(np.abs(w).max(), 1e-12)
        np.testing.assert_allclose(npy(got[k]), w, rtol=1e-4, atol=1e-5 * scale, err_msg=k)
    # Adam step + target update (deepQlearning.py:123-124, soft_actor_critic.py:53-60)
    learner.optimizer.step()
    ac.update_target()
    w1 = _strip(g, f"{tag}_w1_")
    after = {}
    after.update(_mlp_values(enc.obs_1d, "encoder.visible_layer_encoders.obs_1d."))
    after.update(_mlp_values(enc.joiner, "encoder.joiner."))
    after.update(_mlp_values(ac.actor, "actor."))
    after.update(_mlp_values(ac.actor_target, "actor_target."))
    after.update(_ensemble_tensors(ac.critic, "critic.", grad=False))
    after.update(_ensemble_tensors(ac.critic_target, "critic_target.", grad=False))
    after["log_alpha"] = ac.log_alpha.detach()
    assert set(w1) == set(after)
    lr = float(g[f"{tag}_lr"])
    for k, w in w1.items():
        gk = k if k.startswith("encoder.") else "actor_critic." + k
        # Adam's first step moves a weight by lr * g / (|g| + 1e-8): where the reference's gradient is round-off noise around zero
        # (|g| < 1e-6) the step is noise of up to lr, on both sides; everywhere else the weights must agree
        atol = np.where(np.abs(want[gk]) < 1e-6, 1.01 * lr, 2e-5) if gk in want else 2e-5
        assert (np.abs(npy(after[k]) - w) <= atol + 1e-4 * np.abs(w)).all(), k
    np.testing.assert_allclose(float(ac.curr_alpha), float(g[f"{tag}_curr_alpha_after"]), rtol=1e-6)
    # critic_frozen aliases the critic's storage: equal after the optimizer step without any copy (soft_actor_critic.py:142)
    for pf, p in zip(ac.critic_frozen.parameters(), ac.critic.parameters()):
        assert pf.data_ptr() == p.data_ptr() and not pf.requires_grad


# ------------------------------------------------------------------------------------------------ bootstrap bound
class _Fixed:
    def __init__(self, value):
        self.value = value

    def __call__(self, *_):
        return self.value


def test_sac_bootstrap_bound_matches_reference(fdql):
    """soft_actor_critic.py:102-132: q_loss returns the bound as its second value; loss, bound and d/d q_pred vs the reference."""
    import torch
    from fastdeepqlearning_b200 import Agent
    g = load_golden("sac_bootstrap")
    for i in range(int(g["cases"])):
        p = f"sb{i}_"
        CQ = g[p + "q_pred"].shape[-1]
        conf = Agent.LearnerConf(training_device="cuda:0", obs_space={"obs_1d": 4}, obs_keys=("obs_1d",),
                                 action_space=types.SimpleNamespace(shape=(2,)), num_critics=CQ, num_q_predictions=1,
                                 use_distributional_sac=False, use_bootstrap_minibatch_nstep=True, use_max_entropy_q=bool(g[p + "ment"]),
                                 gamma=float(g[p + "gamma"]), temporal_len=int(g[p + "T"]), pi_hidden_dims=(8,), critic_hidden_dims=(8,))
        ac = Agent.SoftActorCritic(conf, 4).to("cuda")
        t = lambda k: torch.as_tensor(g[p + k].astype(np.float32), device="cuda")
        q_pred = t("q_pred").requires_grad_(True)
        ac._critic_io = lambda c, n: (q_pred, t("target_z"), t("log_pi"))  # the three MLPs replaced by the golden's fixed outputs
        ac.curr_alpha.fill_(float(g[p + "alpha"]))
        nxt = {"state": None, "reward": t("reward"), "mask": t("mask"), "mc_return": t("mc_return")}
        loss, bound, summ = ac.q_loss({"state": None}, nxt)
        np.testing.assert_allclose(npy(loss), g[p + "loss"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(npy(bound), g[p + "bound"], rtol=1e-5, atol=1e-6)
        ((loss * t("up_q")).sum() + (bound * t("up_b")).sum()).backward()
        want = g[p + "grad"]
        np.testing.assert_allclose(npy(q_pred.grad), want, rtol=1e-5, atol=1e-5 * np.abs(want).max())
        np.testing.assert_allclose(float(summ["bootstrap_minibatch_nstep_violations"]), float(g[p + "viol"]), rtol=1e-6)


# ------------------------------------------------------------------------------------------------ long episodes / link records
@pytest.mark.parametrize("L,tile", [(1000, 0), (1000, 64), (40000, 64)])
def test_link_records_on_long_episodes(fdql, L, tile):
    """Reference default nStep_return_steps = 1000 (conf.py:46): episodes of 1000 rows keep a chain of equal achieved goals; above
    32767 rows the chain is not built (bit 31 of the link record) and the tile kernel falls back to the verified tail scan.
    Both must equal oracle.sample_time_relabel (goals, dones, exact episode_step bit for bit; returns to 1e-5)."""
    from fastdeepqlearning_b200 import Replay
    rng = np.random.default_rng(L)
    G, gamma, T = 3, 0.999, 2
    ag = rng.integers(0, 2, (L, G)).astype(np.float32)       # 8 distinct goals: every row has many equal successors
    dg = np.tile(rng.integers(0, 2, G).astype(np.float32), (L, 1))
    hit = (ag == dg).all(-1)
    step = np.arange(L, dtype=np.float32).reshape(-1, 1)
    cols = {"obs_1d": rng.standard_normal((L, 5)).astype(np.float32), "achieved_goal": ag, "desired_goal": dg,
            "reward": (hit.astype(np.float32) - 1).reshape(-1, 1), "task_done": hit.astype(np.float32).reshape(-1, 1),
            "episode_done": (step == L - 1).astype(np.float32), "episode_step": step, "mc_return": np.zeros((L, 1), np.float32)}
    cols["reward"][0] = 0.0
    ring = Replay.ReplayMemory(L + 8, 16, T)
    ring.set_reward_op(fdql.RewardOp.bitflip(), gamma)
    ring.add_rows(cols, episode_lengths=[L], with_returns=True)
    cols["mc_return"] = npy(ring.memory["mc_return"])[:L]
    B = 48 if L > 5000 else 160
    starts = rng.integers(0, L - T, B)
    starts[:4] = [0, 1, L - T - 1, L // 2]
    goal_rows = np.array([rng.integers(s, L) for s in starts])
    flags = np.ones(B, bool)
    flags[::5] = False
    want = O.sample_time_relabel(cols, starts, T, flags, goal_rows, np.zeros(L, int), np.full(L, L - 1), O.reward_bitflip, gamma)
    lib = fdql.lib()
    old = lib.fdql_debug_force_generic_gather(tile << 8)
    try:
        got = ring.temporal_sample(starts=starts, flags=flags.astype(np.uint8), goal_rows=goal_rows, exact_episode_step=True, aux=True,
                                   length=L)
    finally:
        lib.fdql_debug_force_generic_gather(old)
    for k in ("desired_goal", "task_done", "episode_step", "reward", "obs_1d"):
        np.testing.assert_array_equal(npy(got[k]), want[k], err_msg=k)
    np.testing.assert_allclose(npy(got["mc_return"]), want["mc_return"], rtol=1e-5, atol=1e-5)
    mask, contig = O.learner_preprocess(want["task_done"], want["episode_step"])
    np.testing.assert_array_equal(npy(got["is_contiguous"]), contig.astype(np.float32))


# ------------------------------------------------------------------------------------------------ small rings, overwritten episodes
def test_small_ring_survives_more_adds_than_rows(fdql):
    """replay_memory.py:38-46 overwrites the ring in place for any maxlen: a ring smaller than the staging block must take more
    than maxlen add() calls between two reads (the staged rows are flushed in blocks of at most maxlen rows)."""
    from fastdeepqlearning_b200 import Replay
    ring = Replay.AsyncReplayMemory(50, 4, 2)
    ref = O.RingOracle(50, 4, 2)
    for i in range(120):
        row = {"x": float(i), "v": np.array([i, -i], np.float32)}
        ring.add(row)
        ref.add(row)
    starts = np.array([0, 10, 30, 47])
    got = ring.replay.temporal_sample(starts=starts)
    want = ref.temporal_sample(starts=starts)
    for k in ("x", "v"):
        np.testing.assert_array_equal(npy(got[k]), want[k].astype(np.float32), err_msg=k)
    np.testing.assert_array_equal(npy(ring.replay.memory["x"]).reshape(-1), ref.memory["x"].reshape(-1))


def test_partly_overwritten_episode_is_not_relabelled(fdql):
    """Once the write head has entered an episode, its surviving rows still name the old first row -- which now belongs to a
    newer episode.  They become uncommitted: never relabelled (flags ignored), gathered verbatim; intact episodes are untouched."""
    from fastdeepqlearning_b200 import Replay, _lib as L
    rng = np.random.default_rng(4)
    cap, Lep, G = 100, 20, 4

    def episode():
        ag = rng.integers(0, 2, (Lep, G)).astype(np.float32)
        step = np.arange(Lep, dtype=np.float32).reshape(-1, 1)
        return {"achieved_goal": ag, "desired_goal": np.ones((Lep, G), np.float32), "reward": -np.ones((Lep, 1), np.float32),
                "task_done": np.zeros((Lep, 1), np.float32), "episode_done": (step == Lep - 1).astype(np.float32), "episode_step": step,
                "mc_return": np.zeros((Lep, 1), np.float32)}
    ring = Replay.ReplayMemory(cap, 4, 2)
    ring.set_reward_op(fdql.RewardOp.bitflip(), 0.9)
    for _ in range(5):
        ring.add_rows(episode(), episode_lengths=[Lep], with_returns=True)   # rows 0..99: five whole episodes
    short = {k: v[:7] for k, v in episode().items()}
    short["episode_done"][-1] = 1.0
    ring.add_rows(short, episode_lengths=[7], with_returns=True)              # overwrites rows 0..6 of the first episode
    es, ee = (npy(t) for t in ring.episode_extents())
    assert (es[:7] == 0).all() and (ee[:7] == 6).all()
    assert (es[7:20] == -1).all() and (ee[7:20] == -1).all()                 # survivors of the first episode: uncommitted
    assert (es[20:40] == 20).all() and (ee[20:40] == 39).all()               # the next episode is intact
    mem = {k: npy(v) for k, v in ring.memory.items()}
    starts = np.array([8, 12, 25])
    got = ring.temporal_sample(starts=starts, flags=np.ones(3, np.uint8), goal_rows=np.array([15, 19, 30]), exact_episode_step=True,
                               length=cap - 1)
    for k in ("desired_goal", "reward", "task_done", "episode_step", "mc_return"):
        np.testing.assert_array_equal(npy(got[k])[:, :2], mem[k][np.arange(2)[:, None] + starts[None, :2]], err_msg=k)
    np.testing.assert_array_equal(npy(got["desired_goal"])[0, 2], mem["achieved_goal"][30])  # the intact episode is relabelled
    # device-drawn streams never flag a survivor (RANDOM mode used to draw goal rows from the overwritten part)
    for _ in range(20):
        s, f, gr = ring.draw_streams(256, 2, goal_mode=L.GOAL_RANDOM, relabel_prob=1.0)
        s, f = npy(s), npy(f)
        assert not f[(s >= 7) & (s < 20)].any() and f[(s >= 20) & (s < 98)].all()


# ------------------------------------------------------------------------------------------------ lean (cp.async + bulk write-back) kernel
@pytest.mark.parametrize("links", [True, False])
@pytest.mark.parametrize("T,G,obs,act,B", [(2, 16, 64, 8, 700), (1, 4, 8, 4, 33), (5, 8, 12, 4, 1000), (50, 16, 64, 8, 90), (2, 16, 64, 8, 4096 + 17)])
def test_lean_kernel_vs_oracle_and_tile_kernel(fdql, T, G, obs, act, B, links):
    """fdql_sample_gather through the lean kernel (cp.async staging of 16-window stages, one bulk shared->global copy per key and
    stage; what FDQL_OPT_CORESIDENT selects): partial chunks, partial stages, windows straddling episode ends and the ring length,
    relabelled and plain windows mixed -- equal to oracle.sample_time_relabel and, bit for bit, to the tile kernel."""
    from fastdeepqlearning_b200 import Replay, _lib as L
    from test_gpu_replay import _synthetic
    rng = np.random.default_rng(T * 1000 + G + B)
    cols, lengths, starts_ep, ends, ep_of = _synthetic(rng, 60, 130, G, obs=obs, act=act)
    N = int(lengths.sum())
    ring = Replay.ReplayMemory(N + 3, 16, T)
    ring.set_reward_op(fdql.RewardOp.bitflip(), 0.98)
    ring.add_rows(cols, episode_lengths=lengths)
    starts = rng.integers(0, N - T, B)
    starts[:3] = [0, N - T - 1, N - T - 1]
    flags = rng.random(B) < 0.8
    goal_rows = np.array([rng.integers(s, ends[ep_of[s]] + 1) for s in starts])
    want = O.sample_time_relabel(cols, starts, T, flags, goal_rows, starts_ep[ep_of], ends[ep_of], O.reward_bitflip, 0.98)
    lib = fdql.lib()
    kw = dict(starts=starts, flags=flags.astype(np.uint8), goal_rows=goal_rows, exact_episode_step=True, aux=T > 1, length=N)
    old = lib.fdql_debug_force_generic_gather(32 | (0 if links else 16))
    try:
        got = {k: v.clone() for k, v in ring.temporal_sample(**kw).items()}
        plain = {k: v.clone() for k, v in ring.temporal_sample(starts=starts, length=N).items()}
        lib.fdql_debug_force_generic_gather((64 << 8) | (0 if links else 16))
        tile = ring.temporal_sample(**kw)
    finally:
        lib.fdql_debug_force_generic_gather(old)
    for k in cols:
        if k in ("reward", "mc_return"):
            np.testing.assert_allclose(npy(got[k]), want[k], rtol=1e-5, atol=1e-6, err_msg=k)
        else:
            np.testing.assert_array_equal(npy(got[k]), want[k], err_msg=k)
    for k in tile:
        if k == "mc_return" and links and T > 32:
            # beyond 32 window rows the tile kernel recomputes the returns from the tail scan, the lean kernel still from the link
            # records (64-bit hit mask): both within tolerance of the oracle (checked above), not bit-identical to each other
            np.testing.assert_allclose(npy(tile[k]), npy(got[k]), rtol=1e-5, atol=1e-6, err_msg=k)
        else:
            assert torch_equal(tile[k], got[k]), k
    idx = np.arange(T)[:, None] + starts[None]
    for k in cols:
        np.testing.assert_array_equal(npy(plain[k]), cols[k][idx], err_msg=k)


def torch_equal(a, b):
    import torch
    return torch.equal(a, b)


@pytest.mark.parametrize("obs,act,G,T,n", [(64, 8, 16, 2, 8192), (12, 4, 8, 5, 1000), (8, 4, 4, 1, 33), (200, 8, 16, 2, 600), (20, 12, 24, 3, 4099)])
def test_coresident_option_selects_the_lean_kernel_and_matches(fdql, obs, act, G, T, n):
    """FDQL_OPT_CORESIDENT through the fused-draw entry point: identical streams and identical batch to the default kernels.  The shapes
    cover keys narrower than one copy round (3, 1 and 2 float4 per row against 4 parts per window), several rounds with a partial last
    one (5 and 6 float4), partial chunks and stages, and a key wider than the lean plan serves (50 float4: the request falls back to the
    tile kernel and must still match)."""
    import ctypes as C
    import torch
    from fastdeepqlearning_b200 import Replay, _lib as L
    from test_gpu_replay import _synthetic
    rng = np.random.default_rng(77 + obs)
    cols, lengths, starts_ep, ends, ep_of = _synthetic(rng, 400, 0, G, obs=obs, act=act, fixed_len=64)
    N = int(lengths.sum())
    ring = Replay.ReplayMemory(N + 1, 4096, T)
    ring.set_reward_op(fdql.RewardOp.bitflip(), 0.99)
    ring.add_rows(cols, episode_lengths=lengths, with_returns=True)
    lib = fdql.lib()
    res = []
    for opt in (0, L.OPT_CORESIDENT):
        out = {k: torch.full((T, n, w), -7.0, device="cuda") for k, w in zip(ring._keys, ring._widths)}
        aux = [torch.empty(T, n, device="cuda"), torch.empty(max(T - 1, 1), n, device="cuda"), torch.empty(max(T - 1, 1), n, device="cuda")]
        st, fl, go = (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda"),
                      torch.empty(n, dtype=torch.int64, device="cuda"))
        params, n_params = ring.reward_op.c_params()
        p = lambda t: C.c_void_p(t.data_ptr())
        L.check(lib.fdql_sample_gather_draw(ring._h, n, T, L.GOAL_FUTURE, 0.8, 5, 3, None, p(st), p(fl), p(go), ring.reward_op.op, params,
                                            n_params, 0.99, (L.OPT_EMIT_LEARNER_AUX if T > 1 else 0) | L.OPT_EXACT_EPISODE_STEP | opt, 4096,
                                            L.ptr_array([out[k].data_ptr() for k in ring._keys]), *[p(t) for t in aux],
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        res.append((out, aux, st, fl, go))
    (o0, a0, s0, f0, g0), (o1, a1, s1, f1, g1) = res
    assert torch.equal(s0, s1) and torch.equal(f0, f1) and torch.equal(g0, g1) and float(f0.float().mean()) > 0.7
    for k in o0:
        if k == "mc_return":  # small batches take the warp-per-window kernel by default, whose relabelled returns come from the tail
            np.testing.assert_allclose(npy(o0[k]), npy(o1[k]), rtol=1e-5, atol=1e-6, err_msg=k)  # scan, not from the link records
        else:
            assert torch.equal(o0[k], o1[k]), k
    if T > 1:
        for x, y in zip(a0, a1):
            assert torch.equal(x, y)


# ------------------------------------------------------------------------------------------------ one launch per pass
@pytest.mark.parametrize("n,T,n_atoms,n_drop,with_lb,with_stats,repeat,ep_len,permute", [
    (24576, 2, 125, 10, True, True, 2, 64, False),    # the fused kernel (16 loss warps + 8 gather warps per SM), headline shape, two passes
    (24576, 2, 125, 10, True, True, 1, 40000, False),  # episodes without a chain of equal goals (> 32767 rows): the T = 2 build's out-of-line tail scan
    (20000, 2, 125, 10, False, False, 1, 64, False),  # fused kernel, flavour without lower bound / summaries, ragged last group and chunk
    (24576, 2, 100, 8, True, False, 1, 64, False),    # fused kernel, 100 atoms (4 x 25)
    (6001, 5, 125, 10, True, True, 1, 64, False),     # fused kernel, five-row windows (link records), ragged
    (1024, 50, 125, 10, True, True, 2, 64, False),    # fused kernel at the reference's default temporal_len (tail scan: windows past the hit mask)
    (3000, 2, 125, 10, True, True, 1, 64, False),     # batch too small for the one-block-per-SM form: the two separate launches
    (24576, 2, 50, 4, True, True, 1, 64, False),      # 64-entry loss tables: separate launches
    (24576, 2, 125, 10, True, True, 1, 64, True),  # scalar keys in another order: not the compiled record columns -> the general build
])
def test_fused_pass_equals_the_two_launches(fdql, n, T, n_atoms, n_drop, with_lb, with_stats, repeat, ep_len, permute):
    """fdql_fused_pass (loss of batch k + gather of batch k+1 in one warp-specialised launch) against fdql_sample_gather_draw and
    fdql_tqc_loss as separate launches on the same arguments: bit-identical batch, loss, gradient; summaries to fp64 rounding.  The
    loss half reads the PREVIOUS gather's reward / mask / mc_return / weight (other buffers), as the learner loop does."""
    import ctypes as C
    import torch
    from fastdeepqlearning_b200 import Replay, _lib as L
    from test_gpu_replay import _synthetic
    G = 16
    rng = np.random.default_rng(5 + n)
    cols, lengths, starts_ep, ends, ep_of = _synthetic(rng, 600 * 64 // ep_len + 1, 0, G, obs=64, act=8, fixed_len=ep_len)
    N = int(lengths.sum())
    if permute:
        order = ["obs_1d", "episode_step", "action", "mc_return", "achieved_goal", "task_done", "desired_goal", "reward", "episode_done"]
        cols = {k: cols[k] for k in order}
    ring = Replay.ReplayMemory(N + 1, 4096, T)
    ring.set_reward_op(fdql.RewardOp.bitflip(), 0.99)
    ring.add_rows(cols, episode_lengths=lengths, with_returns=True)
    lib = fdql.lib()
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    params, n_params = ring.reward_op.c_params()
    opts = L.OPT_EMIT_LEARNER_AUX | L.OPT_EXACT_EPISODE_STEP
    M = (T - 1) * n
    gen = torch.Generator(device="cuda").manual_seed(n)
    z = torch.randn(M, n_atoms, device="cuda", generator=gen) * 3
    q = torch.randn(M, n_atoms, device="cuda", generator=gen) * 3
    lp = torch.randn(M, device="cuda", generator=gen)

    def buf():
        out = {k: torch.full((T, n, w), -7.0, device="cuda") for k, w in zip(ring._keys, ring._widths)}
        return {"out": out, "outp": L.ptr_array([out[k].data_ptr() for k in ring._keys]), "mask": torch.empty(T, n, device="cuda"),
                "contig": torch.empty(T - 1, n, device="cuda"), "weight": torch.empty(T - 1, n, device="cuda"),
                "st": torch.empty(n, dtype=torch.int64, device="cuda"), "fl": torch.empty(n, dtype=torch.uint8, device="cuda"),
                "go": torch.empty(n, dtype=torch.int64, device="cuda")}

    def g_half(b, ctr, n_w=n):
        return (ring._h, n_w, T, L.GOAL_FUTURE, 0.8, 11, ctr, None, p(b["st"]), p(b["fl"]), p(b["go"]), ring.reward_op.op, params, n_params,
                0.99, opts, 4096, b["outp"], p(b["mask"]), p(b["contig"]), p(b["weight"]))

    def t_half(b, m, loss, grad, stats):
        return (m, n_atoms, n_drop, p(z), p(q), p(lp), p(b["out"]["reward"][1:]), p(b["mask"][1:]),
                p(b["out"]["mc_return"][1:]) if with_lb else None, p(b["weight"]), 0.2, 0.99, p(loss), p(grad),
                p(stats) if with_stats else None)
    # --- separate launches: gather(0), then per pass loss(k) and gather(k+1)
    ref = [buf() for _ in range(repeat + 1)]
    ref_loss = []
    for k in range(repeat + 1):
        a = g_half(ref[k], k)  # the fused entry point always asks for the co-resident gather: same kernel on both sides
        L.check(lib.fdql_sample_gather_draw(*a[:15], opts | L.OPT_CORESIDENT, *a[16:], sp))
    for k in range(repeat):
        loss, grad, stats = torch.empty(M, device="cuda"), torch.empty(M, n_atoms, device="cuda"), torch.zeros(4, dtype=torch.float64, device="cuda")
        a = t_half(ref[k], M, loss, grad, stats)
        L.check(lib.fdql_tqc_loss(*a[:14], None, a[14], sp))
        ref_loss.append((loss, grad, stats))
    # --- fused: gather alone (M = 0), then `repeat` fused passes
    got = [buf() for _ in range(repeat + 1)]
    dummy = torch.empty(1, device="cuda")
    L.check(lib.fdql_fused_pass(*g_half(got[0], 0), *t_half(got[0], 0, dummy, dummy, None), sp))
    for k in range(repeat):
        loss, grad, stats = torch.empty(M, device="cuda"), torch.empty(M, n_atoms, device="cuda"), torch.zeros(4, dtype=torch.float64, device="cuda")
        L.check(lib.fdql_fused_pass(*g_half(got[k + 1], k + 1), *t_half(got[k], M, loss, grad, stats), sp))
        torch.cuda.synchronize()
        rl, rg, rs = ref_loss[k]
        assert torch.equal(loss, rl) and torch.equal(grad, rg), k
        if with_stats:
            np.testing.assert_allclose(npy(stats), npy(rs), rtol=1e-12)
    # the loss alone through the same entry point (n_windows = 0)
    loss, grad, stats = torch.empty(M, device="cuda"), torch.empty(M, n_atoms, device="cuda"), torch.zeros(4, dtype=torch.float64, device="cuda")
    L.check(lib.fdql_fused_pass(*g_half(got[0], 0, 0), *t_half(got[0], M, loss, grad, stats), sp))
    torch.cuda.synchronize()
    assert torch.equal(loss, ref_loss[0][0]) and torch.equal(grad, ref_loss[0][1])
    for a, b in zip(ref, got):
        assert torch.equal(a["st"], b["st"]) and torch.equal(a["fl"], b["fl"]) and torch.equal(a["go"], b["go"])
        for k in a["out"]:
            assert torch.equal(a["out"][k], b["out"][k]), k
        for k in ("mask", "contig", "weight"):
            assert torch.equal(a[k], b[k]), k


# ------------------------------------------------------------------------------------------------ SquashRewards on the device
@pytest.mark.parametrize("with_nstep", [True, False])
def test_squash_rewards_in_the_append_kernel(fdql, with_nstep):
    """Replay.make with use_squashed_rewards (Replay/__init__.py:28-29: SquashRewards over NStepReturn, no HER): the stored reward is
    the Pohlen transform of the env reward (golden pohlen_in/out from the reference function) and mc_return is the reference
    recurrence over the SQUASHED rewards."""
    from fastdeepqlearning_b200 import Replay
    g = load_golden("get_losses")
    x, want = g["pohlen_in"], g["pohlen_out"]
    x32 = x.astype(np.float32).astype(np.float64)  # rewards must be fp32-representable (replay_memory.py:31-32)
    want32 = O.pohlen_transform(x32)
    np.testing.assert_allclose(O.pohlen_transform(x), want, rtol=1e-12)
    conf = types.SimpleNamespace(replay_size=512, batch_size=8, temporal_len=2, num_instances=1, use_nStep_lowerbounds=with_nstep,
                                 nStep_return_steps=1000, gamma=0.9, use_squashed_rewards=True, use_HER=False, training_device="cuda:0")
    read, write = Replay.make(conf)
    assert type(write[0]).__name__ == "SquashRewards"
    lengths = [7, 1, 20, 36]
    i = 0
    for L in lengths:
        for t in range(L):
            write[0].add({"obs_1d": np.full(3, i, np.float32), "reward": float(x32[i]), "episode_done": t == L - 1, "episode_step": t})
            i += 1
    n = sum(lengths)
    mem = {k: npy(v)[:n] for k, v in read[0].memory.items()}
    np.testing.assert_array_equal(mem["reward"].reshape(-1), want32[:n].astype(np.float32))
    np.testing.assert_array_equal(mem["obs_1d"][:, 0], np.arange(n, dtype=np.float32))
    if with_nstep:
        done = np.zeros(n, bool)
        done[np.cumsum(lengths) - 1] = True
        np.testing.assert_array_equal(mem["mc_return"].reshape(-1), O.segmented_returns(want32[:n].astype(np.float32), done, 0.9))


# ------------------------------------------------------------------------------------------------ BASELINE.json configs[0..2] shapes
@pytest.mark.parametrize("name", ["config0_pendulum_sac", "config1_cartpole_discrete", "config2_her_bitflip64"])
def test_baseline_config_shapes_vs_oracle(fdql, name):
    """The shapes bench.py times in its `extra` sections (Pendulum: obs 3 / act 1 / SAC min-target head; CartPole: obs 4, discrete
    n = 2, 5 x 10 atoms; 64-bit HER: obs 64, goals 64 + 64, discrete n = 64, 5 x 25 atoms), at test size, against the oracle:
    gathered and relabelled batch, one-hot, loss and gradient."""
    import torch
    import bench
    from fastdeepqlearning_b200 import Replay, ops
    sp = bench.EXTRA_SPECS[name]
    rng = np.random.default_rng(len(name))
    T, B, Lep, n_eps, gamma = 2, 600, 40, 30, 0.99
    N = Lep * n_eps
    step = (np.arange(N) % Lep).astype(np.float32).reshape(-1, 1)
    ends = np.arange(n_eps) * Lep + Lep - 1
    ep_of = np.arange(N) // Lep
    cols = {"obs_1d": rng.standard_normal((N, sp["obs"])).astype(np.float32)}
    cols["action"] = (rng.integers(0, sp["discrete_n"], (N, 1)).astype(np.float32) if sp["discrete_n"]
                      else rng.uniform(-1, 1, (N, sp["act"])).astype(np.float32))
    if sp["goal"]:
        ag = rng.integers(0, 2, (N, sp["goal"])).astype(np.float32)
        dg = rng.integers(0, 2, (n_eps, sp["goal"])).astype(np.float32)[ep_of]
        ag[rng.random(N) < 0.2] = dg[0]  # revisited goals
        hit = (ag == dg).all(-1, keepdims=True).astype(np.float32)
        cols.update(achieved_goal=ag, desired_goal=dg, reward=hit - 1, task_done=hit)
    else:
        cols.update(reward=-rng.random((N, 1)).astype(np.float32), task_done=(step == Lep - 1).astype(np.float32))
    cols.update(episode_done=(step == Lep - 1).astype(np.float32), episode_step=step)
    cols["mc_return"] = O.segmented_returns(cols["reward"], cols["episode_done"], gamma).reshape(-1, 1)
    ring = Replay.ReplayMemory(N + 1, B, T)
    if sp["goal"]:
        ring.set_reward_op(fdql.RewardOp.bitflip(), gamma)
    ring.add_rows(cols, episode_lengths=[Lep] * n_eps, with_returns=True)
    np.testing.assert_array_equal(npy(ring.memory["mc_return"])[:N], cols["mc_return"])
    starts = rng.integers(0, N - T, B)
    if sp["her"]:
        flags = rng.random(B) < 0.8
        goal_rows = np.array([rng.integers(s, ends[ep_of[s]] + 1) for s in starts])
        want = O.sample_time_relabel(cols, starts, T, flags, goal_rows, ep_of * Lep, ends[ep_of], O.reward_bitflip, gamma)
        xp = ring.temporal_sample(starts=starts, flags=flags.astype(np.uint8), goal_rows=goal_rows, exact_episode_step=True, aux=True, length=N)
    else:
        idx = np.arange(T)[:, None] + starts[None]
        want = {k: v[idx] for k, v in cols.items()}
        xp = ring.temporal_sample(starts=starts, aux=True, length=N)
    for k in cols:
        if k in ("reward", "mc_return"):
            np.testing.assert_allclose(npy(xp[k]), want[k], rtol=1e-5, atol=1e-6, err_msg=k)
        else:
            np.testing.assert_array_equal(npy(xp[k]), want[k], err_msg=k)
    mask, contig = O.learner_preprocess(want["task_done"], want["episode_step"])
    np.testing.assert_array_equal(npy(xp["mask"]), mask.astype(np.float32))
    np.testing.assert_array_equal(npy(xp["is_contiguous"]), contig.astype(np.float32))
    if sp["discrete_n"]:
        np.testing.assert_array_equal(npy(ops.action_onehot(xp["action"], sp["discrete_n"])), O.action_onehot(want["action"], sp["discrete_n"]))
    CQ = sp["critics"] * sp["atoms"]
    g = torch.Generator(device="cuda").manual_seed(1)
    z, q = torch.randn(T - 1, B, CQ, device="cuda", generator=g) * 3, torch.randn(T - 1, B, CQ, device="cuda", generator=g) * 3
    lp = torch.randn(T - 1, B, 1, device="cuda", generator=g)
    f64 = lambda t: npy(t).astype(np.float64)
    if sp["loss"] == "sac":
        got = ops.sac_min_target_loss(q, z, lp, xp["reward"][1:], xp["mask"][1:], xp["mc_return"][1:], 0.7, gamma)
        wl, wg, _ = O.sac_min_target_loss(f64(q), f64(z), f64(lp), want["reward"][1:], mask[1:], want["mc_return"][1:], 0.7, gamma)
    else:
        got = ops.tqc_loss(q, z, lp, xp["reward"][1:], xp["mask"][1:], xp["mc_return"][1:], 0.7, gamma, sp["n_drop"])
        wl, wg, _ = O.tqc_q_loss(f64(q), f64(z), f64(lp), want["reward"][1:].astype(np.float64), mask[1:].astype(np.float64),
                                 want["mc_return"][1:].astype(np.float64), 0.7, gamma, sp["n_drop"])
    np.testing.assert_allclose(npy(got["loss"]), wl, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(npy(got["grad"]), wg, rtol=1e-5, atol=1e-5 * np.abs(wg).max())


@pytest.mark.parametrize("n_atoms", [1, 5, 12, 16])
def test_sac_thread_kernel_equals_warp_kernel(fdql, n_atoms):
    """Narrow heads (<= 16 atoms) take one thread per transition; same results as the warp-per-transition kernel and the oracle."""
    import torch
    from fastdeepqlearning_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n_atoms)
    M = 3001
    z, q = torch.randn(M, n_atoms, device="cuda", generator=g) * 2, torch.randn(M, n_atoms, device="cuda", generator=g) * 2
    lp, rew, mc = (torch.randn(M, 1, device="cuda", generator=g) for _ in range(3))
    mask = (torch.rand(M, 1, device="cuda", generator=g) > 0.2).float()
    w = torch.rand(M, 1, device="cuda", generator=g)
    a = ops.sac_min_target_loss(q, z, lp, rew, mask, mc, 0.6, 0.97, grad_scale=w, want_stats=True)
    lib = fdql.lib()
    old = lib.fdql_debug_tqc_warp_kernel(1)
    try:
        b = ops.sac_min_target_loss(q, z, lp, rew, mask, mc, 0.6, 0.97, grad_scale=w, want_stats=True)
    finally:
        lib.fdql_debug_tqc_warp_kernel(old)
    torch.testing.assert_close(a["loss"], b["loss"], rtol=1e-6, atol=1e-7)
    assert torch.equal(a["grad"], b["grad"])
    torch.testing.assert_close(a["stats"], b["stats"], rtol=1e-6, atol=1e-6)  # per-row sums in fp32, different orders
    f64 = lambda t: npy(t).astype(np.float64)
    wl, wg, _ = O.sac_min_target_loss(f64(q), f64(z), f64(lp), f64(rew), f64(mask), f64(mc), 0.6, 0.97)
    np.testing.assert_allclose(npy(a["loss"]), wl, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(npy(a["grad"]), wg * f64(w), rtol=1e-5, atol=1e-6)
