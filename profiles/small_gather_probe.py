"""Per-launch time of the fused gather for one 4096-window batch: warp-per-window kernels vs the tile kernel (with link records)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastdeepqlearning_b200 as pkg
from fastdeepqlearning_b200 import Replay
import bench
dev = torch.device("cuda:0")
ring = bench.build_ring(torch, pkg, Replay, 2_000_000, dev, 1)
lib = pkg.lib()
for B in (1024, 4096, 16384):
    starts, flags, goals = ring.draw_streams(B, relabel_prob=0.8, goal_mode=None)
    for name, flag in (("warp-per-window", 0), ("tile32", 32 << 8), ("tile64", 64 << 8), ("tile128", 128 << 8), ("tile256", 256 << 8),
                       ("tile64-scan", (64 << 8) | 16)):
        lib.fdql_debug_force_generic_gather(flag)
        for _ in range(5):
            ring.temporal_sample(starts=starts, flags=flags, goal_rows=goals, aux=True, reuse_outputs=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            ring.temporal_sample(starts=starts, flags=flags, goal_rows=goals, aux=True, reuse_outputs=True)
        e1.record()
        torch.cuda.synchronize()
        print(f"B={B:6d} {name:16s} {e0.elapsed_time(e1) / 200 * 1e3:8.1f} us per call (incl. Python launch overhead)")
lib.fdql_debug_force_generic_gather(0)
