"""Write-path throughput: rows/s through Replay.make write heads, one dict per env step like franQ's Runner (runner.py:177-191)."""
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
