"""Generate tests/golden/*.npz by EXECUTING the unmodified reference -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):  ``python -m oracle.make_goldens``.
The GPU box only ever sees the committed .npz files.  Every golden records the inputs that
were injected (index streams, goal picks, MLP outputs) and what the reference returned.
"""
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def bitflip_reward(ag, dg):
    # restated from franQ/Env/bitflip.py:143-152 without its debug print (gym is absent, so the
    # env class itself cannot be constructed here)
    m = (np.asarray(ag) == np.asarray(dg)).all()
    r = 0.0 * m + ((1.0 - m) * -1.0)
    return r, r == 0


def all_geq_reward(ag, dg):
    # restated from franQ/Env/classic_control_goal/classic_goal.py:88-93
    c = (np.asarray(ag) >= np.asarray(dg)).all()
    r = 0.0 * c + ((-1.0) * (1 - c))
    return r, r == 0


def first_geq_reward(ag, dg):
    # restated from classic_goal.py:306-311
    d = bool(ag[0] >= dg[0])
    return float(d) - 1.0, d


def golden_ring(ref):
    rng = np.random.default_rng(1)
    maxlen, B, T = 257, 32, 5
    mem = ref.ReplayMemory(maxlen, B, T)
    n_rows = 700
    obs = rng.standard_normal((n_rows, 6)).astype(np.float32)
    act = rng.uniform(-1, 1, (n_rows, 2)).astype(np.float32)
    goal = rng.integers(0, 2, (n_rows, 4)).astype(np.int64)
    reward = rng.integers(-3, 4, n_rows).astype(np.float64) * 0.5
    done = rng.random(n_rows) < 0.1
    step = np.arange(n_rows) % 37
    tops, lens = [], []
    snap = {}
    for i in range(n_rows):
        mem.add({"obs_1d": obs[i], "action": act[i], "desired_goal": goal[i], "reward": float(reward[i]),
                 "task_done": bool(done[i]), "episode_step": int(step[i])})
        tops.append(mem._top)
        lens.append(len(mem))
        if i == 199:  # partially filled snapshot
            starts_a = rng.integers(0, len(mem) - T, B)
            got = mem._temporal_sample_idxes(starts_a, len(mem))
            snap.update({f"partial_{k}": v for k, v in got.items()})
            snap["partial_starts"] = starts_a
            snap["partial_len"] = len(mem)
    starts = rng.integers(0, len(mem) - T, B)
    win = mem._temporal_sample_idxes(starts, len(mem))
    flat_idx = rng.integers(0, len(mem), B)
    flat = mem[flat_idx]
    out = dict(maxlen=maxlen, B=B, T=T, in_obs=obs, in_act=act, in_goal=goal, in_reward=reward, in_done=done,
               in_step=step, tops=np.array(tops), lens=np.array(lens), starts=starts, flat_idx=flat_idx, **snap)
    out.update({f"win_{k}": v for k, v in win.items()})
    out.update({f"flat_{k}": v for k, v in flat.items()})
    out.update({f"mem_{k}": v for k, v in mem.memory.items()})
    # error behaviour: under-filled ring must raise OversampleError (replay_memory.py:50,57-58)
    small = ref.ReplayMemory(100, 8, 5)
    for i in range(7):
        small.add({"x": float(i)})
    try:
        small.temporal_sample()
        out["oversample_raised"] = False
    except ref.OversampleError:
        out["oversample_raised"] = True
    np.savez_compressed(os.path.join(GOLDEN, "ring.npz"), **out)


def golden_nstep(ref):
    out = {}
    # (a) the reference's own KAT, tests/test_replays.py:16-33
    disc, n = 0.99, 1000
    mem = ref.ReplayMemory(1001, 128, 1)
    w = ref.NStepReturn(mem, n_step=n, discount=disc)
    for i in range(n):
        w.add({"reward": float(i == n - 1), "episode_done": i == n - 1, "step": i})
    out["kat_mc_return"] = mem.memory["mc_return"][:n].copy()
    out["kat_step"] = mem.memory["step"][:n].copy()
    assert np.allclose(out["kat_mc_return"], disc ** (n - 1 - out["kat_step"]))
    # (b) random rewards, several episodes of ragged length incl. length 1
    rng = np.random.default_rng(2)
    lengths = [1, 2, 7, 33, 128, 5, 64, 1, 31]
    mem = ref.ReplayMemory(1000, 8, 2)
    w = ref.NStepReturn(mem, n_step=1000, discount=0.97)
    rewards, dones = [], []
    for L in lengths:
        for t in range(L):
            r = float(np.float32(rng.standard_normal()))
            d = t == L - 1
            w.add({"reward": r, "episode_done": d, "episode_step": t})
            rewards.append(r)
            dones.append(d)
    tot = sum(lengths)
    out.update(rand_gamma=0.97, rand_lengths=np.array(lengths), rand_reward=np.array(rewards, np.float32),
               rand_done=np.array(dones), rand_mc_return=mem.memory["mc_return"][:tot].copy(),
               rand_stored_reward=mem.memory["reward"][:tot].copy())
    # (c) quirk Q3: n_step < episode length -> oldest row stored twice (nstep_return.py:33-34,50-57)
    mem = ref.ReplayMemory(100, 4, 2)
    w = ref.NStepReturn(mem, n_step=3, discount=0.9)
    rs = [1.0, 2.0, -1.0, 0.5, 4.0, -2.0]
    for t, r in enumerate(rs):
        w.add({"reward": r, "episode_done": t == len(rs) - 1, "episode_step": t})
    out.update(q3_rewards=np.array(rs, np.float32), q3_n_rows=len(mem),
               q3_mc_return=mem.memory["mc_return"][:len(mem)].copy(),
               q3_step=mem.memory["episode_step"][:len(mem)].copy())
    np.savez_compressed(os.path.join(GOLDEN, "nstep.npz"), **out)


def _run_her(ref, mode, reward_fn, episodes, picks, gamma, goal_dtype):
    mem = ref.ReplayMemory(4096, 8, 2)
    inner = ref.NStepReturn(mem, 1000, gamma)
    her = ref.HindsightNStepReplay(inner, reward_fn, mode=mode)
    pick_iter = iter(picks)
    real_choice = random.choice

    def injected_choice(seq):  # deque is newest-first (her.py:29-31): chronological t <-> len-1-t
        t = next(pick_iter)
        return seq[len(seq) - 1 - t]

    random.choice = injected_choice
    try:
        for ep in episodes:
            L = len(ep["reward"])
            for t in range(L):
                her.add({"obs_1d": ep["obs"][t], "action": ep["action"][t],
                         "achieved_goal": ep["ag"][t].astype(goal_dtype), "desired_goal": ep["dg"][t].astype(goal_dtype),
                         "reward": float(ep["reward"][t]), "task_done": bool(ep["task_done"][t]),
                         "episode_done": t == L - 1, "episode_step": t, "info": {}})
    finally:
        random.choice = real_choice
    n = len(mem)
    return {k: v[:n].copy() for k, v in mem.memory.items()}


def _bitflip_episodes(rng, n_eps, n_bits, max_len):
    eps = []
    for _ in range(n_eps):
        L = int(rng.integers(1, max_len + 1))
        dg = rng.integers(0, 2, n_bits)
        state = rng.integers(0, 2, n_bits)
        ag, rew, td, obs, act = [], [], [], [], []
        for t in range(L):
            if t > 0:
                a = int(rng.integers(0, n_bits))
                state = state.copy()
                state[a] = 1 - state[a]
            else:
                a = 0
            hit = bool((state == dg).all())
            ag.append(state.copy())
            # env_handler.py:38: the reset row carries reward 0.0; later rows carry the env reward
            rew.append(0.0 if t == 0 else (0.0 if hit else -1.0))
            td.append(hit and t > 0)
            obs.append(state.astype(np.float32))
            act.append(np.array([a], np.float32))
            if hit and t > 0:
                break
        L = len(rew)
        eps.append(dict(obs=np.stack(obs), action=np.stack(act), ag=np.stack(ag), dg=np.tile(dg, (L, 1)),
                        reward=np.array(rew), task_done=np.array(td)))
    return eps


def _float_goal_episodes(rng, n_eps, g, max_len):
    eps = []
    for _ in range(n_eps):
        L = int(rng.integers(2, max_len + 1))
        ag = np.round(rng.standard_normal((L, g)) * 2) / 2  # coarse grid so >= comparisons hit sometimes
        dg = np.tile(np.round(rng.standard_normal(g) * 2) / 2, (L, 1))
        dg[L // 2:] += 0.5 * (rng.random() < 0.3)  # goal that changes mid-episode exercises per-row dg (Q6)
        rew = rng.integers(-2, 2, L).astype(np.float64) * 0.25
        eps.append(dict(obs=rng.standard_normal((L, 3)).astype(np.float32),
                        action=rng.uniform(-1, 1, (L, 1)).astype(np.float32), ag=ag, dg=dg, reward=rew,
                        task_done=np.zeros(L, bool)))
    return eps


def golden_her(ref):
    out = {}
    rng = np.random.default_rng(3)
    cases = [("bitflip", bitflip_reward, _bitflip_episodes(rng, 24, 3, 12), np.int64),
             ("all_geq", all_geq_reward, _float_goal_episodes(rng, 12, 2, 10), np.float64),
             ("first_geq", first_geq_reward, _float_goal_episodes(rng, 12, 2, 10), np.float64)]
    for name, fn, eps, gdt in cases:
        picks = [int(rng.integers(0, len(e["reward"]))) for e in eps]
        out[f"{name}_n_eps"] = len(eps)
        out[f"{name}_picks"] = np.array(picks)
        out[f"{name}_lengths"] = np.array([len(e["reward"]) for e in eps])
        for k in ("obs", "action", "ag", "dg", "reward", "task_done"):
            out[f"{name}_in_{k}"] = np.concatenate([e[k] for e in eps])
        for mode in ("final", "random"):
            stored = _run_her(ref, mode, fn, eps, picks, 0.98, gdt)
            for k, v in stored.items():
                out[f"{name}_{mode}_{k}"] = v
    out["gamma"] = 0.98
    np.savez_compressed(os.path.join(GOLDEN, "her.npz"), **out)


def _run_her_vmap(ref, reward_fn, episodes, picks, gamma, goal_dtype, V, with_returns):
    """The reference's vmap write chain HindsightVmapWrite -> NStepReturnVmap -> ReplayMemory (Replay/__init__.py:21-23,30-31),
    executed with the numpy stand-in for jax.vmap (oracle/ref_loader.py); goal picks injected through np.random.randint."""
    mem = ref.ReplayMemory(4096, 8, 2)
    inner = ref.NStepReturnVmap(mem, 1000, gamma) if with_returns else mem
    her = ref.HindsightVmapWrite(inner, reward_fn, num_virtual_goals=V)
    pick_iter = iter(picks)
    real_randint = np.random.randint

    def injected_randint(low, high=None, size=None, **kw):  # her_vmap.py:75 indexes the newest-first deque
        p = np.asarray(next(pick_iter))
        assert p.shape == (size,) and p.min() >= low and p.max() < high
        return p

    np.random.randint = injected_randint
    try:
        for ep in episodes:
            L = len(ep["reward"])
            for t in range(L):
                her.add({"obs_1d": ep["obs"][t], "action": ep["action"][t],
                         "achieved_goal": ep["ag"][t].astype(goal_dtype), "desired_goal": ep["dg"][t].astype(goal_dtype),
                         "reward": float(ep["reward"][t]), "task_done": bool(ep["task_done"][t]),
                         "episode_done": t == L - 1, "episode_step": t, "info": {}})
    finally:
        np.random.randint = real_randint
    n = len(mem)
    return mem, {k: np.asarray(v[:n]).copy() for k, v in mem.memory.items()}


def golden_her_vmap(ref):
    """her_vmap.py / nstep_return_vmap.py: stored rows of the write chain and what the read head returns for every column."""
    out = {}
    rng = np.random.default_rng(11)
    V = 5
    # episodes of a single row are left out: the reference's calculate_montecarlo_return squeezes [1, V+1] rewards to [V+1],
    # walks the goal columns as if they were time steps and then fails with an IndexError (nstep_return_vmap.py:62-66,43-44)
    cases = [("bitflip", bitflip_reward, [e for e in _bitflip_episodes(rng, 20, 3, 12) if len(e["reward"]) > 1], np.int64),
             ("all_geq", all_geq_reward, _float_goal_episodes(rng, 10, 2, 10), np.float64)]
    for name, fn, eps, gdt in cases:
        # deque index (newest first) of every virtual goal, as np.random.randint(0, L, V) would return it
        picks = [rng.integers(0, len(e["reward"]), V) for e in eps]
        out[f"{name}_n_eps"] = len(eps)
        out[f"{name}_picks_deque"] = np.concatenate(picks)
        out[f"{name}_lengths"] = np.array([len(e["reward"]) for e in eps])
        for k in ("obs", "action", "ag", "dg", "reward", "task_done"):
            out[f"{name}_in_{k}"] = np.concatenate([e[k] for e in eps])
        for tag, with_returns in (("ret", True), ("noret", False)):
            mem, stored = _run_her_vmap(ref, fn, eps, picks, 0.98, gdt, V, with_returns)
            for k, v in stored.items():
                out[f"{name}_{tag}_{k}"] = v.astype(np.float64) if v.dtype == bool else v
            if with_returns:  # read head: inject the window starts and the column (random.randint(0, V), her_vmap.py:107)
                read = ref.HindsightVmapRead(mem)
                starts = rng.integers(0, len(mem) - 2, 8)
                out[f"{name}_read_starts"] = starts
                real_ts, real_ri = mem.temporal_sample, random.randint
                mem.temporal_sample = lambda: mem._temporal_sample_idxes(starts, len(mem))
                try:
                    for col in range(V + 1):
                        random.randint = lambda a, b, _c=col: _c
                        got = read.temporal_sample()
                        assert "virtual_goals" not in got
                        for k, v in got.items():
                            out[f"{name}_read{col}_{k}"] = np.asarray(v).astype(np.float64)
                finally:
                    mem.temporal_sample, random.randint = real_ts, real_ri
    out["gamma"] = 0.98
    out["V"] = V
    np.savez_compressed(os.path.join(GOLDEN, "her_vmap.npz"), **out)


class _Fixed(torch.nn.Module):
    def __init__(self, value):
        super().__init__()
        self.value = value

    def forward(self, *_):
        return self.value


def _conf(ref, tmp, C, Q, drop, discrete=False, distributional=True, lower_bound=True, max_ent=True):
    conf = ref.AgentConf()
    A = ref.AttrDict
    conf.obs_space = A(spaces={"obs_1d": A(shape=(6,)), "achieved_goal": A(shape=(3,)), "desired_goal": A(shape=(3,))})
    conf.action_space = A(n=3) if discrete else A(shape=(2,))
    conf.discrete = discrete
    conf.log_dir = tmp
    conf.training_device = conf.inference_device = "cpu"
    conf.num_critics, conf.num_q_predictions, conf.top_quantiles_to_drop = C, Q, drop
    conf.use_distributional_sac = distributional
    conf.use_nStep_lowerbounds = lower_bound
    conf.use_max_entropy_q = max_ent
    conf.latent_state_dim = 16
    conf.pi_hidden_dims, conf.critic_hidden_dims = [8], [8, 8]
    conf.encoder_conf.hidden_features = 8
    conf.encoder_conf.joint_hidden_dims = (8,)
    conf.encoder_conf.obs_1d_hidden_dims = (8,)
    conf.init_log_alpha = -0.5
    return conf


def golden_tqc(ref):
    out = {}
    torch.manual_seed(4)
    rng = np.random.default_rng(4)
    # (a) the free function on a spread of magnitudes (offsets stress the prefix-sum form)
    for i, (scale, off, n, k) in enumerate([(3.0, 0.0, 125, 115), (0.3, -50.0, 125, 115), (1.0, 5.0, 50, 40),
                                            (0.05, 0.0, 20, 16), (10.0, 100.0, 125, 115), (1.0, 0.0, 7, 3)]):
        q = (torch.randn(3, 9, n) * scale + off).requires_grad_(True)
        s = torch.randn(3, 9, k) * scale + off
        loss = ref.quantile_huber_loss_f(q, s)
        loss.sum().backward()
        out.update({f"qh{i}_q": q.detach().numpy(), f"qh{i}_s": s.numpy(), f"qh{i}_loss": loss.detach().numpy(),
                    f"qh{i}_grad": q.grad.numpy()})
    out["qh_cases"] = 6
    # (b) the reference q_loss method itself, with the three MLPs replaced by fixed outputs
    with tempfile.TemporaryDirectory() as tmp:
        ci = 0
        for (C, Q, drop, lb, ment, Tm1, B) in [(5, 25, 0.08, True, True, 1, 64), (5, 25, 0.08, True, True, 3, 17),
                                                 (2, 10, 0.2, True, True, 2, 33), (5, 25, 0.08, False, False, 1, 40),
                                                 (3, 7, 0.1, True, False, 1, 5)]:
            conf = _conf(ref, tmp, C, Q, drop, lower_bound=lb, max_ent=ment)
            ac = ref.DistributionalSoftActorCritic(conf, conf.latent_state_dim)
            CQ = C * Q
            next_z = torch.randn(Tm1, B, CQ) * 3
            next_z[0, 0, :4] = next_z[0, 0, 4]  # ties in the sort
            q_pred = (torch.randn(Tm1, B, CQ) * 3).requires_grad_(True)
            log_pi = torch.randn(Tm1, B, 1)
            state = torch.randn(Tm1, B, conf.latent_state_dim)
            action = torch.randn(Tm1, B, 2)
            reward = torch.tensor(rng.integers(-1, 1, (Tm1, B, 1)).astype(np.float32))
            task_done = torch.tensor((rng.random((Tm1, B, 1)) < 0.2).astype(np.float32))
            mask = torch.logical_not(task_done)
            mc = torch.randn(Tm1, B, 1) * 3
            ac.actor_target = _Fixed((action, log_pi, None))
            ac.critic_target = _Fixed(next_z)
            ac.critic = _Fixed(q_pred)
            ac.curr_alpha = float(np.exp(-0.5))
            curr = {"state": state, "action": action}
            nxt = {"state": state, "reward": reward, "mask": mask, "mc_return": mc}
            q_loss, _, summ = ac.q_loss(curr, nxt)
            upstream = torch.rand_like(q_loss)
            (q_loss * upstream).sum().backward()
            out.update({f"ql{ci}_next_z": next_z.numpy(), f"ql{ci}_q_pred": q_pred.detach().numpy(),
                        f"ql{ci}_log_pi": log_pi.numpy(), f"ql{ci}_reward": reward.numpy(),
                        f"ql{ci}_mask": mask.numpy(), f"ql{ci}_mc_return": mc.numpy(),
                        f"ql{ci}_alpha": ac.curr_alpha, f"ql{ci}_gamma": conf.gamma,
                        f"ql{ci}_n_drop": int(drop * CQ), f"ql{ci}_lb": lb, f"ql{ci}_ment": ment,
                        f"ql{ci}_loss": q_loss.detach().numpy(), f"ql{ci}_upstream": upstream.numpy(),
                        f"ql{ci}_grad": q_pred.grad.numpy(),
                        f"ql{ci}_q_pred_mu": float(summ["q_pred_mu"]), f"ql{ci}_q_pred_var": float(summ["q_pred_var"]),
                        f"ql{ci}_viol": float(summ.get("mc_constraint_violations", -1.0))})
            ci += 1
        out["ql_cases"] = ci
        # (c) non-distributional SAC q_loss (soft_actor_critic.py:63-134)
        si = 0
        for (C, Q, lb, ment) in [(5, 1, True, True), (2, 10, True, False), (3, 4, False, True)]:
            conf = _conf(ref, tmp, C, Q, 0.2, distributional=False, lower_bound=lb, max_ent=ment)
            ac = ref.SoftActorCritic(conf, conf.latent_state_dim)
            CQ, Tm1, B = C * Q, 2, 29
            tz = torch.randn(Tm1, B, CQ) * 2
            q_pred = (torch.randn(Tm1, B, CQ) * 2).requires_grad_(True)
            log_pi = torch.randn(Tm1, B, 1)
            state = torch.randn(Tm1, B, conf.latent_state_dim)
            action = torch.randn(Tm1, B, 2)
            reward = torch.randn(Tm1, B, 1)
            mask = torch.tensor(rng.random((Tm1, B, 1)) > 0.2)
            mc = torch.randn(Tm1, B, 1) * 2
            ac.actor_target, ac.critic_target, ac.critic = _Fixed((action, log_pi, None)), _Fixed(tz), _Fixed(q_pred)
            ac.curr_alpha = 0.7
            q_loss, _, summ = ac.q_loss({"state": state, "action": action},
                                        {"state": state, "reward": reward, "mask": mask, "mc_return": mc})
            upstream = torch.rand_like(q_loss)
            (q_loss * upstream).sum().backward()
            out.update({f"sac{si}_target_z": tz.numpy(), f"sac{si}_q_pred": q_pred.detach().numpy(),
                        f"sac{si}_log_pi": log_pi.numpy(), f"sac{si}_reward": reward.numpy(), f"sac{si}_mask": mask.numpy(),
                        f"sac{si}_mc_return": mc.numpy(), f"sac{si}_alpha": 0.7, f"sac{si}_gamma": conf.gamma,
                        f"sac{si}_lb": lb, f"sac{si}_ment": ment, f"sac{si}_loss": q_loss.detach().numpy(),
                        f"sac{si}_upstream": upstream.numpy(), f"sac{si}_grad": q_pred.grad.numpy(),
                        f"sac{si}_viol": float(summ.get("mc_constraint_violations", -1.0))})
            si += 1
        out["sac_cases"] = si
    np.savez_compressed(os.path.join(GOLDEN, "tqc.npz"), **out)


def golden_get_losses(ref):
    """DeepQLearning.get_losses pre/post-processing (deepQlearning.py:198-258): record mask,
    is_contiguous, the q_loss the critic head returned, the final scalar and d loss / d q_loss."""
    out = {}
    torch.manual_seed(5)
    rng = np.random.default_rng(5)
    with tempfile.TemporaryDirectory() as tmp:
        for ci, discrete in enumerate([False, True]):
            conf = _conf(ref, tmp, 2, 5, 0.2, discrete=discrete)
            conf.temporal_len = T = 6
            B = 11
            agent = ref.DeepQLearning(conf)
            step = np.zeros((T, B, 1), np.float32)
            start = rng.integers(0, 20, B)
            for b in range(B):
                s = np.arange(T) + start[b]
                if b % 3 == 0:
                    cut = int(rng.integers(1, T))
                    s[cut:] = np.arange(T - cut)  # episode boundary inside the window
                step[:, b, 0] = s
            task_done = (rng.random((T, B, 1)) < 0.15).astype(np.float32)
            xp = {"obs_1d": torch.randn(T, B, 6), "achieved_goal": torch.randn(T, B, 3),
                  "desired_goal": torch.randn(T, B, 3),
                  "action": torch.tensor(rng.integers(0, 3, (T, B, 1)).astype(np.float32)) if discrete else torch.rand(T, B, 2) * 2 - 1,
                  "reward": torch.randn(T, B, 1), "task_done": torch.tensor(task_done),
                  "episode_step": torch.tensor(step), "mc_return": torch.randn(T, B, 1)}
            captured = {}
            real_q_loss = agent.actor_critic.q_loss

            def spy(curr, nxt, _real=real_q_loss, _cap=captured):
                q_loss, lbv, summ = _real(curr, nxt)
                q_loss.retain_grad()
                _cap["q_loss"] = q_loss
                if discrete:
                    _cap["onehot"] = curr["action_onehot"].detach().clone()
                return q_loss, lbv, summ

            agent.actor_critic.q_loss = spy
            inputs = {k: v.clone().numpy() for k, v in xp.items()}
            loss = agent.get_losses(xp)
            loss.backward()
            out.update({f"gl{ci}_{k}": v for k, v in inputs.items()})
            out.update({f"gl{ci}_mask": xp["mask"].numpy(), f"gl{ci}_is_contiguous": xp["is_contiguous"].numpy(),
                        f"gl{ci}_q_loss": captured["q_loss"].detach().numpy(),
                        f"gl{ci}_dloss_dq_loss": captured["q_loss"].grad.numpy(), f"gl{ci}_loss": loss.item(),
                        f"gl{ci}_T": T})
            if discrete:
                out[f"gl{ci}_onehot"] = captured["onehot"].numpy()
    # loss reduce alone on arbitrary per-step losses (deepQlearning.py:222-225,249)
    T, B = 7, 13
    per = torch.randn(T - 1, B, 1)
    contig = torch.tensor(rng.random((T - 1, B, 1)) < 0.7)
    red = ((per * contig).sum(0) / (contig.float().sum(0) + 1e-4)).mean() / T
    out.update(red_per=per.numpy(), red_contig=contig.numpy(), red_T=T, red_loss=red.item())
    x = rng.standard_normal(64) * 5
    out.update(pohlen_in=x, pohlen_out=ref.pohlen_transform(x))
    np.savez_compressed(os.path.join(GOLDEN, "get_losses.npz"), **out)


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle.ref_loader import load_reference
    os.makedirs(GOLDEN, exist_ok=True)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as scratch:
        os.chdir(scratch)  # the reference writes tensorboard logs relative to cwd
        try:
            ref = load_reference()
            golden_ring(ref)
            golden_nstep(ref)
            golden_her(ref)
            golden_her_vmap(ref)
            golden_tqc(ref)
            golden_get_losses(ref)
        finally:
            os.chdir(cwd)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
