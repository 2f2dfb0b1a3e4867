"""Read-head throughput of the vmap hindsight variant: windows/s of HindsightVmapRead on a ring whose rows carry 32 virtual goals."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastdeepqlearning_b200 as fdql
from fastdeepqlearning_b200 import Replay
from fastdeepqlearning_b200.Replay import wrappers as W
rng = np.random.default_rng(0)
V, G, L, n_eps, B = 32, 16, 128, 2000, 65536
shard = Replay.AsyncReplayMemory(L * n_eps + 1, B, 2)
inner = W.NStepReturnVmap(shard, 1000, 0.99)
her = W.HindsightVmapWrite(inner, fdql.RewardOp.bitflip(), num_virtual_goals=V)
t0 = time.perf_counter()
for e in range(n_eps):
    ag = rng.integers(0, 2, (L, G)).astype(np.float32)
    dg = np.tile(rng.integers(0, 2, G).astype(np.float32), (L, 1))
    cols = {"obs_1d": rng.standard_normal((L, 64)).astype(np.float32), "action": rng.uniform(-1, 1, (L, 8)).astype(np.float32),
            "achieved_goal": ag, "desired_goal": dg, "reward": -np.ones((L, 1), np.float32), "task_done": np.zeros((L, 1), np.float32),
            "episode_done": (np.arange(L) == L - 1).astype(np.float32).reshape(-1, 1), "episode_step": np.arange(L, dtype=np.float32).reshape(-1, 1)}
    cols["virtual_goals"] = torch.zeros((L, (V + 1) * G), device="cuda")
    cols["virtual_rewards"] = torch.zeros((L, V + 1), device="cuda")
    cols["virtual_dones"] = torch.zeros((L, V + 1), device="cuda")
    inner.add_vmap_rows(cols, L, rng.integers(0, L, V))
torch.cuda.synchronize()
print(f"write: {n_eps * L / (time.perf_counter() - t0):.0f} rows/s (episode-batched, {V} virtual goals per row)")
read = W.HindsightVmapRead(shard, aux=True)
for _ in range(3):
    read.temporal_sample(column=5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    read.temporal_sample(column=i % (V + 1))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"read: {ms:.3f} ms per {B} windows = {B / ms / 1e3:.1f} M windows/s (T=2)")
