"""Runner integration: ReplayHandler = Runner._replay_handler (runner.py:177-191) feeding the device
write heads from per-stream threads while the learner samples; ParamPublisher = DeepQLearning._push_params / _pull_params
(deepQlearning.py:136-148) over pinned double buffers."""
import types

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


def test_replay_handler_threads_store_the_reference_row_stream(fdql):
    """Two actor streams, each fed through its own handler thread, one dict at a time with `info` attached (runner.py:185-186 drops
    it unless HER is on), while this thread keeps sampling: every shard ends up holding exactly the rows the reference's
    HindsightNStepReplay(NStepReturn(...)) stored for the same stream (tests/golden/her.npz)."""
    import random
    from fastdeepqlearning_b200 import Replay, Runner
    g = load_golden("her")
    name, mode = "bitflip", "final"  # "final" needs no goal pick, so the two threads share no injected state
    conf = types.SimpleNamespace(replay_size=4096, batch_size=8, temporal_len=2, num_instances=2, use_nStep_lowerbounds=True,
                                 nStep_return_steps=1000, gamma=float(g["gamma"]), use_squashed_rewards=False, use_HER=True,
                                 her_mode=mode, training_device="cuda:0")
    read, write = Replay.make(conf, compute_reward=fdql.RewardOp.bitflip())
    handler = Runner.ReplayHandler(write, use_HER=True)
    lengths = g[f"{name}_lengths"]
    off = 0
    sampled = 0

    def sample_all():
        k = 0
        for r in read:
            if len(r) >= 16:
                batch = r.temporal_sample()
                assert tuple(batch["obs_1d"].shape) == (2, 8, 3)
                k += 1
        return k
    for L in lengths:
        for t in range(L):
            i = off + t
            xp = {"obs_1d": g[f"{name}_in_obs"][i], "action": g[f"{name}_in_action"][i], "achieved_goal": g[f"{name}_in_ag"][i].astype(np.int64),
                  "desired_goal": g[f"{name}_in_dg"][i].astype(np.int64), "reward": float(g[f"{name}_in_reward"][i]),
                  "task_done": bool(g[f"{name}_in_task_done"][i]), "episode_done": t == L - 1, "episode_step": t, "info": {"idx": i}}
            for shard in range(2):
                handler.put(shard, xp)
            xp["reward"] = 99.0  # the handler owns a copy (runner.py:161)
        off += L
        sampled += sample_all()  # the learner thread reads between episodes, whatever the handler threads have stored by then
    while handler.pending():  # ... and keeps reading while they drain their queues
        sampled += sample_all()
    handler.join()
    handler.close()
    sampled += sample_all()
    assert sampled >= 2  # (how many reads overlap the writes depends on thread timing; at least the final ones see every row)
    n = 2 * int(lengths.sum())
    for r in read:
        assert len(r) == n
        mem = {k: npy(v)[:n] for k, v in r.memory.items()}
        assert "info" not in mem
        for k in ("obs_1d", "action", "achieved_goal", "desired_goal", "task_done", "episode_done", "episode_step", "reward", "mc_return"):
            np.testing.assert_array_equal(mem[k], g[f"{name}_{mode}_{k}"].astype(np.float32), err_msg=k)


def test_param_publisher_round_trip(fdql):
    """The trainer's weights reach an inference-side copy (a CPU policy, like the reference's inference_device="cpu" presets):
    snapshots are consistent (taken between optimizer steps), the newest completed one wins, and the training stream never
    blocks on the host."""
    import torch
    from fastdeepqlearning_b200 import Agent, Runner
    from fastdeepqlearning_b200.Agent.components import models
    torch.manual_seed(0)
    conf = Agent.LearnerConf(training_device="cuda:0", obs_space={"obs_1d": 6}, obs_keys=("obs_1d",),
                             action_space=types.SimpleNamespace(shape=(2,)), num_critics=2, num_q_predictions=5, pi_hidden_dims=(16,),
                             critic_hidden_dims=(16,), batch_size=8, temporal_len=2)
    learner = Agent.Learner(conf)
    pub = Runner.ParamPublisher({"actor": learner.actor_critic.actor}, interval=2)
    assert pub.pull() is None
    inference = models.GaussianPolicy(6, 2, (16,))  # the inference copy lives on the CPU
    versions = []
    for step in range(1, 7):
        with torch.no_grad():
            for p in learner.actor_critic.actor.parameters():
                p.add_(0.01 * step)  # stands in for an optimizer step
        want = {k: v.detach().cpu().clone() for k, v in learner.actor_critic.actor.state_dict().items()}
        if pub.maybe_push(step):
            got, version = pub.pull(wait=True)
            versions.append(version)
            assert set(got) == {"actor"} and set(got["actor"]) == set(want)
            for k in want:
                assert torch.equal(got["actor"][k], want[k]), k
            inference.load_state_dict(got["actor"])
    assert versions == [1, 2, 3]
    x = torch.randn(5, 6)
    a_cpu = inference(x)[2]
    a_gpu = learner.actor_critic.actor(x.cuda())[2].cpu()
    torch.testing.assert_close(a_cpu, a_gpu, rtol=1e-5, atol=1e-6)
