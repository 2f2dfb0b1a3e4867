"""probe: time 4096-window sample_gather launches for different tile sizes of the tile kernel (development aid)"""
import os, sys, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastdeepqlearning_b200 as pkg
from fastdeepqlearning_b200 import Replay
import bench
dev = torch.device('cuda:0')
ring = bench.build_ring(torch, pkg, Replay, 10_000_000, dev, 1)
lib = pkg.lib()
for tile in (0, 32, 64, 128, 256):
    lib.fdql_debug_force_generic_gather(tile << 8)
    for n in (4096, 32768):
        for _ in range(5):
            ring.temporal_sample(n=n, relabel_prob=0.8, aux=True, reuse_outputs=True)
        s, f, gl = ring.draw_streams(n, relabel_prob=0.8)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            ring.temporal_sample(starts=s, flags=f, goal_rows=gl, aux=True, reuse_outputs=True)
        e1.record(); torch.cuda.synchronize()
        print(f"tile={tile} n={n}: {e0.elapsed_time(e1)/reps*1e3:.1f} us per launch (incl. python call overhead)")
lib.fdql_debug_force_generic_gather(8)
for n in (4096, 32768):
    s, f, gl = ring.draw_streams(n, relabel_prob=0.8)
    for _ in range(5): ring.temporal_sample(starts=s, flags=f, goal_rows=gl, aux=True, reuse_outputs=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): ring.temporal_sample(starts=s, flags=f, goal_rows=gl, aux=True, reuse_outputs=True)
    e1.record(); torch.cuda.synchronize()
    print(f"warp-per-window kernel n={n}: {e0.elapsed_time(e1)/50*1e3:.1f} us")
