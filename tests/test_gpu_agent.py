"""GPU parity of the Agent mirror: q_loss of DistributionalSoftActorCritic / SoftActorCritic (fused kernels) vs the
reference's operator sequence (oracle.tqc_q_loss_torch / sac restatement) on the same MLP outputs, forward and backward,
and one Learner.train_step end to end."""
import types

import numpy as np
import pytest

from oracle import cpu_restatement as O

pytestmark = pytest.mark.gpu


def make_conf(Agent, **kw):
    d = dict(training_device="cuda:0", obs_space={"obs_1d": 64, "achieved_goal": 16, "desired_goal": 16},
             action_space=types.SimpleNamespace(shape=(8,)), num_critics=5, num_q_predictions=25, top_quantiles_to_drop=0.08,
             batch_size=64, temporal_len=3, pi_hidden_dims=(32,), critic_hidden_dims=(32, 32))
    d.update(kw)
    return Agent.LearnerConf(**d)


def random_xp(torch, T, B, dev="cuda"):
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    step = torch.arange(T, device=dev).float().view(T, 1, 1).expand(T, B, 1).clone()
    step[:, ::7] = 0  # some non-contiguous windows
    done = (torch.rand(T, B, 1, device=dev, generator=g) < 0.1).float()
    return {"obs_1d": r(T, B, 64), "achieved_goal": r(T, B, 16), "desired_goal": r(T, B, 16), "action": torch.tanh(r(T, B, 8)),
            "reward": r(T, B, 1), "task_done": done, "episode_step": step, "mc_return": r(T, B, 1) * 3}


@pytest.mark.parametrize("distributional", [True, False])
def test_q_loss_matches_reference_operator_sequence(fdql, distributional):
    import torch
    from fastdeepqlearning_b200 import Agent
    torch.manual_seed(0)
    conf = make_conf(Agent, use_distributional_sac=distributional)
    L = Agent.Learner(conf)
    ac = L.actor_critic
    xp = random_xp(torch, 3, 64)
    xp["mask"] = 1 - xp["task_done"]
    xp["state"] = L.encoder.forward_train(xp)
    curr, nxt = L._temporal_difference_shift(xp)
    captured = {}
    orig = ac._critic_io

    def spy(c, n):
        out = orig(c, n)
        captured["io"] = out
        return out
    ac._critic_io = spy
    q_loss, extra, summ = ac.q_loss(curr, nxt)
    assert extra is None and tuple(q_loss.shape) == (2, 64, 1)
    q_pred, next_z, next_log_pi = captured["io"]
    w = torch.rand_like(q_loss)
    (gq,) = torch.autograd.grad((q_loss * w).sum(), q_pred)
    c = lambda t: t.detach().cpu()
    qp = c(q_pred).clone().requires_grad_(True)
    if distributional:
        ref = O.tqc_q_loss_torch(qp, c(next_z), c(next_log_pi), c(nxt["reward"]), c(nxt["mask"]), c(nxt["mc_return"]),
                                 float(ac.curr_alpha), conf.gamma, int(0.08 * 125))
    else:
        lo, gr, _ = O.sac_min_target_loss(c(q_pred).numpy(), c(next_z).numpy(), c(next_log_pi).numpy(), c(nxt["reward"]).numpy(),
                                          c(nxt["mask"]).numpy(), c(nxt["mc_return"]).numpy(), float(ac.curr_alpha), conf.gamma)
        np.testing.assert_allclose(c(q_loss).numpy(), lo, rtol=1e-5, atol=1e-6)
        want = gr * c(w).numpy()
        np.testing.assert_allclose(c(gq).numpy(), want, rtol=1e-5, atol=1e-5 * np.abs(want).max())
        return
    np.testing.assert_allclose(c(q_loss).numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    (gref,) = torch.autograd.grad((ref * c(w)).sum(), qp)
    np.testing.assert_allclose(c(gq).numpy(), gref.numpy(), rtol=1e-5, atol=1e-5 * float(gref.abs().max()))
    np.testing.assert_allclose(float(summ["q_pred_mu"]), float(q_pred.mean()), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(float(summ["q_pred_var"]), float(q_pred.var(-1).mean()), rtol=1e-4)
    lb = (nxt["mc_return"] - q_pred).relu()
    np.testing.assert_allclose(float(summ["mc_constraint_violations"]), float((lb > 0).float().mean()), rtol=1e-5)


def test_get_losses_uses_kernel_aux_and_matches_torch_preprocessing(fdql):
    """mask / is_contiguous emitted by the gather kernel == the torch formulas of deepQlearning.py:201-203."""
    import torch
    from fastdeepqlearning_b200 import Agent
    torch.manual_seed(1)
    conf = make_conf(Agent)
    xp = random_xp(torch, 3, 64)
    a, b = dict(xp), dict(xp)
    mask = 1 - xp["task_done"]
    b["mask"] = mask
    b["is_contiguous"] = ((xp["episode_step"][1:] == xp["episode_step"][:-1] + 1) & (mask[:-1] != 0)).float()
    L = Agent.Learner(conf)
    st = torch.cuda.get_rng_state()
    la = L.get_losses(a)
    torch.cuda.set_rng_state(st)
    lb = L.get_losses(b)
    assert torch.allclose(la, lb, rtol=1e-6)
    assert "state" in a and "mask" in a  # the sampled dict is mutated like in the reference


def test_train_step_end_to_end(fdql):
    import torch
    from fastdeepqlearning_b200 import Agent, Replay
    torch.manual_seed(0)
    conf = make_conf(Agent, replay_size=20000, use_HER=True, her_mode="future", num_instances=1, temporal_len=2, batch_size=256)
    read, write = Replay.make(conf, compute_reward=fdql.RewardOp.bitflip())
    rng = np.random.default_rng(0)
    Lep, n_eps = 32, 200
    N = Lep * n_eps
    ag = rng.integers(0, 2, (N, 16)).astype(np.float32)
    dg = np.repeat(rng.integers(0, 2, (n_eps, 16)).astype(np.float32), Lep, 0)
    hit = (ag == dg).all(-1, keepdims=True).astype(np.float32)
    step = (np.arange(N) % Lep).astype(np.float32).reshape(-1, 1)
    write[0].add_rows({"obs_1d": rng.standard_normal((N, 64)).astype(np.float32), "action": rng.uniform(-1, 1, (N, 8)).astype(np.float32),
                       "achieved_goal": ag, "desired_goal": dg, "reward": hit - 1, "task_done": hit,
                       "episode_done": (step == Lep - 1).astype(np.float32), "episode_step": step}, episode_lengths=[Lep] * n_eps)
    learner = Agent.Learner(conf, read)
    before = [p.detach().clone() for p in learner.params]
    tgt_before = [p.detach().clone() for p in learner.actor_critic.critic_target.parameters()]
    losses = [float(learner.train_step()) for _ in range(5)]
    assert all(np.isfinite(losses))
    assert any(not torch.equal(a, b) for a, b in zip(before, learner.params))
    assert any(not torch.equal(a, b) for a, b in zip(tgt_before, learner.actor_critic.critic_target.parameters()))
    assert learner.train_steps == 5


def test_cuda_graph_step_matches_eager_and_draws_fresh_streams(fdql):
    """The whole learner step captured in a CUDA graph: same update as eager from the same state, and every replay draws new
    windows (device-side draw counter) with the current temperature (device-side alpha)."""
    import copy
    import torch
    from fastdeepqlearning_b200 import Agent, Replay
    rng = np.random.default_rng(0)
    Lep, n_eps = 32, 400
    N = Lep * n_eps
    ag = rng.integers(0, 2, (N, 16)).astype(np.float32)
    dg = np.repeat(rng.integers(0, 2, (n_eps, 16)).astype(np.float32), Lep, 0)
    hit = (ag == dg).all(-1, keepdims=True).astype(np.float32)
    step = (np.arange(N) % Lep).astype(np.float32).reshape(-1, 1)
    cols = {"obs_1d": rng.standard_normal((N, 64)).astype(np.float32), "action": rng.uniform(-1, 1, (N, 8)).astype(np.float32),
            "achieved_goal": ag, "desired_goal": dg, "reward": hit - 1, "task_done": hit,
            "episode_done": (step == Lep - 1).astype(np.float32), "episode_step": step}

    def build(graph):
        torch.manual_seed(0)
        conf = make_conf(Agent, replay_size=N + 1, use_HER=True, her_mode="future", num_instances=1, temporal_len=2, batch_size=512,
                         use_cuda_graph=graph)
        read, write = Replay.make(conf, compute_reward=fdql.RewardOp.bitflip())
        write[0].add_rows(cols, episode_lengths=[Lep] * n_eps)
        return Agent.Learner(conf, read), read[0]

    lg, head = build(True)
    losses = [float(lg.train_step()) for _ in range(6)]
    assert all(np.isfinite(losses)) and len(set(losses)) > 1, losses   # fresh windows each replay
    ring = head.replay_buffer.replay
    assert int(ring._rng_counter_dev[0]) >= 6 and int(ring._rng_counter_dev[1]) == 0
    a0 = float(lg.actor_critic.curr_alpha)
    # temperature follows log_alpha inside the graph (it is refreshed before the optimizer step, so it lags by one Adam step)
    assert abs(a0 - float(torch.exp(lg.actor_critic.log_alpha.detach()))) < 2e-3 and a0 != 1.0
    le, _ = build(False)
    le_losses = [float(le.train_step()) for _ in range(3)]
    assert all(np.isfinite(le_losses))
    # the captured step IS the eager step: the first train_step() of a graphed learner runs three eager warm-up updates and one replay,
    # with draw counters 0..3 (device side) and the generator's next four noise draws -- four eager steps from the same state must
    # leave the same weights, targets and temperature
    lg2, _ = build(True)
    lg2.train_step()
    le2, _ = build(False)
    for _ in range(4):
        le2.train_step()
    for (n1, p1), (n2, p2) in zip(lg2.actor_critic.state_dict().items(), le2.actor_critic.state_dict().items()):
        assert n1 == n2
        torch.testing.assert_close(p1, p2, rtol=1e-5, atol=1e-6, msg=lambda m, n1=n1: f"{n1}: {m}")


def test_graphed_learner_is_captured_once_while_the_ring_fills(fdql):
    """The sampling range of the captured launch comes from device memory (counter_dev[2]): one capture serves a ring that grows
    by far more than 5 %, and the new rows are sampled."""
    import torch
    from fastdeepqlearning_b200 import Agent, Replay
    rng = np.random.default_rng(1)
    Lep = 32

    def episodes(n_eps):
        N = Lep * n_eps
        ag = rng.integers(0, 2, (N, 16)).astype(np.float32)
        dg = np.repeat(rng.integers(0, 2, (n_eps, 16)).astype(np.float32), Lep, 0)
        hit = (ag == dg).all(-1, keepdims=True).astype(np.float32)
        step = (np.arange(N) % Lep).astype(np.float32).reshape(-1, 1)
        return {"obs_1d": rng.standard_normal((N, 64)).astype(np.float32), "action": rng.uniform(-1, 1, (N, 8)).astype(np.float32),
                "achieved_goal": ag, "desired_goal": dg, "reward": hit - 1, "task_done": hit,
                "episode_done": (step == Lep - 1).astype(np.float32), "episode_step": step}
    torch.manual_seed(0)
    conf = make_conf(Agent, replay_size=40000, use_HER=True, her_mode="future", num_instances=1, temporal_len=2, batch_size=512,
                     use_cuda_graph=True)
    read, write = Replay.make(conf, compute_reward=fdql.RewardOp.bitflip())
    write[0].add_rows(episodes(40), episode_lengths=[Lep] * 40)          # 1280 rows
    learner = Agent.Learner(conf, read)
    learner.train_step()
    graph0 = learner._graphs[0][0]
    ring = read[0].replay_buffer.replay
    assert int(ring.last_streams[0].max()) < 1280
    write[0].add_rows(episodes(900), episode_lengths=[Lep] * 900)        # 30080 rows: 23x the ring the graph was captured on
    seen = 0
    for _ in range(4):
        assert np.isfinite(float(learner.train_step()))
        seen = max(seen, int(ring.last_streams[0].max()))
    assert learner._graphs[0][0] is graph0, "the step must not be re-captured when the ring grows"
    assert seen > 20000, "the captured launch must sample the rows added after the capture"


def test_split_backward_is_the_same_update(fdql):
    """Learner._one_update under torch.distributed backpropagates the critic part first, so that the critic bucket's all-reduce runs
    under the actor's backward (two backward calls on disjoint parameters, two buckets).  Forced on one GPU: the weights after four
    updates equal those of the single-backward learner from the same state, eagerly and as a captured graph."""
    import torch
    from fastdeepqlearning_b200 import Agent, Replay
    rng = np.random.default_rng(1)
    Lep, n_eps = 32, 300
    N = Lep * n_eps
    ag = rng.integers(0, 2, (N, 16)).astype(np.float32)
    dg = np.repeat(rng.integers(0, 2, (n_eps, 16)).astype(np.float32), Lep, 0)
    hit = (ag == dg).all(-1, keepdims=True).astype(np.float32)
    step = (np.arange(N) % Lep).astype(np.float32).reshape(-1, 1)
    cols = {"obs_1d": rng.standard_normal((N, 64)).astype(np.float32), "action": rng.uniform(-1, 1, (N, 8)).astype(np.float32),
            "achieved_goal": ag, "desired_goal": dg, "reward": hit - 1, "task_done": hit,
            "episode_done": (step == Lep - 1).astype(np.float32), "episode_step": step}

    def run(split, graph):
        torch.manual_seed(0)
        conf = make_conf(Agent, replay_size=N + 1, use_HER=True, her_mode="future", num_instances=1, temporal_len=2, batch_size=256,
                         use_cuda_graph=graph, split_backward=split)
        read, write = Replay.make(conf, compute_reward=fdql.RewardOp.bitflip())
        write[0].add_rows(cols, episode_lengths=[Lep] * n_eps)
        learner = Agent.Learner(conf, read)
        losses = [float(learner.train_step()) for _ in range(1 if graph else 4)]  # (a graphed learner's first step = 3 eager + 1 replay)
        return learner, losses

    base, l0 = run(False, False)
    for split, graph in (("force", False), ("force", True)):
        other, l1 = run(split, graph)
        assert other._buckets is not None and len(other._buckets[0]) > 0 and len(other._buckets[1]) > 0
        if not graph:
            np.testing.assert_allclose(l1, l0, rtol=1e-5)
        for (n1, p1), (n2, p2) in zip(base.actor_critic.state_dict().items(), other.actor_critic.state_dict().items()):
            assert n1 == n2
            np.testing.assert_allclose(p1.detach().float().cpu().numpy(), p2.detach().float().cpu().numpy(), rtol=2e-5, atol=2e-6, err_msg=n1)


def test_data_parallel_gradients_equal_the_concatenated_batch(fdql, monkeypatch):
    """DP over N shards steps on the AVERAGE of the ranks' gradients; that equals one learner on the
    concatenation of the ranks' injected batches (the loss is a mean over the batch), <= 1e-5 of each gradient's scale.  Two
    'ranks' are emulated on one GPU: same weights, disjoint batches, the rank-mean formed as NCCL's all-reduce(sum) / world does."""
    import torch
    from fastdeepqlearning_b200 import Agent
    torch.manual_seed(3)
    conf = make_conf(Agent, batch_size=48, temporal_len=3)
    learner = Agent.Learner(conf)
    T, Bh = 3, 48
    xps = [random_xp(torch, T, Bh) for _ in range(2)]
    g = torch.Generator(device="cuda").manual_seed(7)
    for xp in xps:  # random_xp is seeded: make the two shards differ
        for k in ("obs_1d", "achieved_goal", "desired_goal", "reward", "mc_return"):
            xp[k] = xp[k] + torch.randn(xp[k].shape, device="cuda", generator=g)
    noise = [[torch.randn(T - 1, Bh, 8, device="cuda", generator=g) for _ in range(2)] for _ in range(2)]  # (actor_target, actor) per shard

    def grads_of(xp, eps):
        tape = iter(eps)
        monkeypatch.setattr(torch, "randn_like", lambda t, **kw: next(tape))
        learner.optimizer.zero_grad(set_to_none=True)
        learner.get_losses(dict(xp)).backward()
        monkeypatch.undo()
        return [p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p) for p in learner.params]
    per_rank = [grads_of(xps[r], noise[r]) for r in range(2)]
    mean = [(a + b) / 2 for a, b in zip(*per_rank)]
    cat = {k: torch.cat([xps[0][k], xps[1][k]], dim=1) for k in xps[0]}
    whole = grads_of(cat, [torch.cat([noise[0][i], noise[1][i]], dim=1) for i in range(2)])
    for m, w in zip(mean, whole):
        scale = float(w.abs().max()) + 1e-12
        assert float((m - w).abs().max()) <= 1e-5 * scale + 1e-9


def test_action_onehot_kernel_golden_and_random(fdql):
    """deepQlearning.py:206-210: eye(n)[action.long()] -- the one-hot the reference's get_losses built for the discrete golden case,
    and torch's own indexing on random actions."""
    import torch
    from conftest import load_golden
    from fastdeepqlearning_b200 import ops
    g = load_golden("get_losses")
    a = torch.as_tensor(g["gl1_action"], dtype=torch.float32, device="cuda")
    got = ops.action_onehot(a, 3)
    np.testing.assert_array_equal(got.cpu().numpy()[:-1], g["gl1_onehot"])
    gen = torch.Generator(device="cuda").manual_seed(0)
    act = torch.randint(0, 64, (5, 4096, 1), device="cuda", generator=gen).float()
    want = torch.eye(64, device="cuda")[act.view(5, 4096).long()]
    assert torch.equal(ops.action_onehot(act, 64), want)


def test_discrete_learner_config(fdql):
    """BASELINE.json configs[1]: discrete Gumbel-softmax SAC on CartPole-shaped rows (obs 4, 2 actions, no goals) with n-step
    returns: rows go through Replay.make's NStepReturn head, the learner one-hots the stored action index and steps."""
    import torch
    from fastdeepqlearning_b200 import Agent, Replay
    torch.manual_seed(0)
    conf = make_conf(Agent, obs_space={"obs_1d": 4}, obs_keys=("obs_1d",), action_space=types.SimpleNamespace(n=2, shape=()), discrete=True,
                     replay_size=5000, use_HER=False, num_instances=1, temporal_len=2, batch_size=256, num_q_predictions=10,
                     top_quantiles_to_drop=0.2, nStep_return_steps=1000)
    read, write = Replay.make(conf)
    rng = np.random.default_rng(0)
    for ep in range(40):
        L = int(rng.integers(8, 60))
        for t in range(L):
            write[0].add({"obs_1d": rng.standard_normal(4).astype(np.float32), "action": float(rng.integers(0, 2)), "reward": 1.0,
                          "task_done": t == L - 1, "episode_done": t == L - 1, "episode_step": t})
    learner = Agent.Learner(conf, read)
    xp = read[0].temporal_sample()
    assert tuple(xp["action"].shape) == (2, 256, 1) and "mc_return" in xp
    loss = learner.get_losses(dict(xp))
    assert torch.isfinite(loss)
    losses = [float(learner.train_step()) for _ in range(3)]
    assert all(np.isfinite(losses))
    a, lp, logits = learner.actor_critic.actor(torch.randn(7, 4, device="cuda"))
    assert tuple(a.shape) == (7, 2) and bool(((a.detach() == 0) | (a.detach() == 1)).all()) and bool((a.detach().sum(-1) == 1).all())
    assert bool((lp <= 0).all()) and tuple(lp.shape) == (7, 1)
