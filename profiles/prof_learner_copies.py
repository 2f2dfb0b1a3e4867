"""Which tensor copies does one learner step launch?  (torch profiler, grouped by input shape and by Python stack)"""
import sys, types, torch, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastdeepqlearning_b200 as pkg
from fastdeepqlearning_b200 import Agent, Replay
from fastdeepqlearning_b200.Replay.wrappers import SampleTimeHindsight
import bench
ring = bench.build_ring(torch, pkg, Replay, 1_000_000, torch.device('cuda:0'), 1)
conf = Agent.LearnerConf(training_device='cuda:0', obs_space={"obs_1d": 64, "achieved_goal": 16, "desired_goal": 16},
                         action_space=types.SimpleNamespace(shape=(8,)), num_critics=5, num_q_predictions=25,
                         top_quantiles_to_drop=10/125+1e-9, batch_size=4096, temporal_len=2)
L = Agent.Learner(conf, [SampleTimeHindsight(ring, relabel_prob=0.8)])
for _ in range(5): L.train_step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    for _ in range(5): L.train_step()
    torch.cuda.synchronize()
ka = prof.key_averages(group_by_input_shape=True)
rows = [e for e in ka if e.key in ("aten::copy_", "aten::cat", "aten::sum", "aten::mul", "aten::contiguous", "aten::clone", "aten::add", "aten::add_", "aten::fill_", "aten::zero_")]
rows.sort(key=lambda e: -e.self_device_time_total)
for e in rows[:40]:
    print(f"{e.key:18s} n={e.count:4d} cuda_us={e.self_device_time_total:9.1f} shapes={str(e.input_shapes)[:110]}")
ka2 = prof.key_averages(group_by_stack_n=6)
rows = [e for e in ka2 if e.key == "aten::copy_"]
rows.sort(key=lambda e: -e.self_device_time_total)
for e in rows[:12]:
    print(f"copy_ n={e.count} cuda_us={e.self_device_time_total:.1f}")
    for fr in e.stack[:6]: print("     ", fr[:150])
