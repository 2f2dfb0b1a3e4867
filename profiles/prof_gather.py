"""Launch the gather kernels alone at the bench's shape (262144 windows, T = 2, 1e7-row ring by default) for ncu captures.
usage: python profiles/prof_gather.py [lean|lean2|tile] [ring_rows]   (lean2 = the co-resident form bench.py runs: FDQL_OPT_CORESIDENT)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import fastdeepqlearning_b200 as pkg  # noqa: E402
from fastdeepqlearning_b200 import Replay, _lib as L  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "lean"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
lib = pkg.lib()
dev = torch.device("cuda:0")
B, T = bench.B, bench.T
n = 64 * B
ring = bench.build_ring(torch, pkg, Replay, rows, dev, seed=1)
keys, h = ring.keys, ring._h
out = {k: torch.empty((T, n, w), device=dev) for k, w in zip(keys, ring._widths)}
outp = L.ptr_array([out[k].data_ptr() for k in keys])
aux = [torch.empty(T, n, device=dev), torch.empty(T - 1, n, device=dev), torch.empty(T - 1, n, device=dev)]
st, fl, go = (torch.empty(n, dtype=torch.int64, device=dev), torch.empty(n, dtype=torch.uint8, device=dev),
              torch.empty(n, dtype=torch.int64, device=dev))
p = lambda t: C.c_void_p(t.data_ptr())
params, n_params = ring.reward_op.c_params()
opts = L.OPT_EMIT_LEARNER_AUX | L.OPT_EXACT_EPISODE_STEP | (L.OPT_CORESIDENT if which == "lean2" else 0)
lib.fdql_debug_force_generic_gather(32 if which == "lean" else 0)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(i):
    L.check(lib.fdql_sample_gather_draw(h, n, T, L.GOAL_FUTURE, 0.8, 7, i, None, p(st), p(fl), p(go), ring.reward_op.op, params, n_params,
                                        bench.GAMMA, opts, B, outp, *[p(t) for t in aux], sp))


for i in range(3):
    run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    run(3 + i)
e1.record()
torch.cuda.synchronize()
print(which, "ms per launch", e0.elapsed_time(e1) / 10, "relabelled", float(fl.float().mean()))
