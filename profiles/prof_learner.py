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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): L.train_step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=18, max_name_column_width=50))
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=50))
