"""Write-path throughput: rows/s through Replay.make write heads, one dict per env step like franQ's Runner (runner.py:177-191),
and through the batched protocol (add_rows: whole episodes as [n, w] arrays, what a vectorised actor hands over)."""
import sys, os, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastdeepqlearning_b200 as fdql
from fastdeepqlearning_b200 import Agent, Replay
rng = np.random.default_rng(0)
G, L, n_eps = 16, 128, 200
for mode in ("plain", "final", "future", "vmap"):
    conf = Agent.LearnerConf(training_device="cuda:0", replay_size=200_000, batch_size=256, temporal_len=2, num_instances=1,
                             use_HER=mode != "plain", her_mode=mode if mode != "plain" else "final", gamma=0.99)
    read, write = Replay.make(conf, compute_reward=fdql.RewardOp.bitflip())
    rows = []
    for e in range(n_eps):
        dg = rng.integers(0, 2, G).astype(np.float32)
        for t in range(L):
            ag = rng.integers(0, 2, G).astype(np.float32)
            rows.append({"obs_1d": rng.standard_normal(64).astype(np.float32), "action": rng.uniform(-1, 1, 8).astype(np.float32),
                         "achieved_goal": ag, "desired_goal": dg, "reward": -1.0, "task_done": False, "episode_done": t == L - 1,
                         "episode_step": t})
            if mode != "plain":
                rows[-1]["info"] = {}  # the Runner drops `info` unless HER is on (runner.py:185-186)
    w = write[0]
    for r in rows[:L]:
        w.add(r)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in rows[L:]:
        w.add(r)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{mode:7s} {len(rows) - L} env rows in {dt:.2f} s = {(len(rows) - L) / dt:9.0f} env rows/s  (ring holds {len(read[0])} rows)", flush=True)

# batched protocol: episodes arrive as arrays (pinned host memory), 64 episodes per call
conf = Agent.LearnerConf(training_device="cuda:0", replay_size=4_000_000, batch_size=256, temporal_len=2, num_instances=1, use_HER=True,
                         her_mode="future", gamma=0.99)
read, write = Replay.make(conf, compute_reward=fdql.RewardOp.bitflip())
ne = 64
n = ne * L
step = (np.arange(n) % L).astype(np.float32).reshape(-1, 1)
cols = {"obs_1d": rng.standard_normal((n, 64)).astype(np.float32), "action": rng.uniform(-1, 1, (n, 8)).astype(np.float32),
        "achieved_goal": rng.integers(0, 2, (n, G)).astype(np.float32), "desired_goal": rng.integers(0, 2, (n, G)).astype(np.float32),
        "reward": -np.ones((n, 1), np.float32), "task_done": np.zeros((n, 1), np.float32), "episode_done": (step == L - 1).astype(np.float32),
        "episode_step": step, "mc_return": np.zeros((n, 1), np.float32)}
cols = {k: torch.from_numpy(v).pin_memory() for k, v in cols.items()}
ring = read[0].replay_buffer
for _ in range(3):
    ring.add_rows(cols, episode_lengths=[L] * ne, with_returns=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 100
for _ in range(reps):
    ring.add_rows(cols, episode_lengths=[L] * ne, with_returns=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"batched add_rows ({ne} episodes of {L} rows per call, returns + extents + link records on the device): "
      f"{reps * n / dt:9.0f} env rows/s", flush=True)
