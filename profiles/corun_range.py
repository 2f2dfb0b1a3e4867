"""Range capture of the co-run (ncu --replay-mode range): the lean gather of pass k+1 under the TQC loss of pass k, against the two
kernels back to back.  Whole-range SM metrics (issue slots, pipes, shared-memory wavefronts, stall reasons) of both schedules.
usage: ncu --replay-mode range --section ... python profiles/corun_range.py [serial|corun|tqc|gather]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import fastdeepqlearning_b200 as pkg  # noqa: E402
from fastdeepqlearning_b200 import Replay, _lib as L  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "corun"
lib = pkg.lib()
dev = torch.device("cuda:0")
B, T, CQ = bench.B, bench.T, bench.CQ
n = 64 * B
M = n
ring = bench.build_ring(torch, pkg, Replay, 2_000_000, dev, seed=1)
keys, h = ring.keys, ring._h
g = torch.Generator(device=dev).manual_seed(0)
z = torch.randn(M, CQ, device=dev, generator=g) * 3
q = torch.randn(M, CQ, device=dev, generator=g) * 3
lp = torch.randn(M, device=dev, generator=g)
p = lambda t: C.c_void_p(t.data_ptr())
params, n_params = ring.reward_op.c_params()
opts = L.OPT_EMIT_LEARNER_AUX | L.OPT_EXACT_EPISODE_STEP


def make_buf():
    out = {k: torch.empty((T, n, w), device=dev) for k, w in zip(keys, ring._widths)}
    return {"out": out, "outp": L.ptr_array([out[k].data_ptr() for k in keys]), "mask": torch.empty(T, n, device=dev),
            "contig": torch.empty(T - 1, n, device=dev), "weight": torch.empty(T - 1, n, device=dev),
            "starts": torch.empty(n, dtype=torch.int64, device=dev), "flags": torch.empty(n, dtype=torch.uint8, device=dev),
            "goals": torch.empty(n, dtype=torch.int64, device=dev)}


bufs = [make_buf(), make_buf()]
loss, grad = torch.empty(M, device=dev), torch.empty(M, CQ, device=dev)
stats = torch.zeros(4, dtype=torch.float64, device=dev)
ctr = [0]
sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
co = mode in ("corun", "gather_lean")
lib.fdql_set_coresident(1 if co else 0)
gopt = L.OPT_CORESIDENT if co else 0


def gather(b, st):
    L.check(lib.fdql_sample_gather_draw(h, n, T, L.GOAL_FUTURE, 0.8, 7, ctr[0], None, p(b["starts"]), p(b["flags"]), p(b["goals"]),
                                        ring.reward_op.op, params, n_params, bench.GAMMA, opts | gopt, B, b["outp"], p(b["mask"]),
                                        p(b["contig"]), p(b["weight"]), C.c_void_p(st.cuda_stream)))
    ctr[0] += 1


def tqc(b, st):
    L.check(lib.fdql_tqc_loss(M, CQ, 10, p(z), p(q), p(lp), p(b["out"]["reward"][1:]), p(b["mask"][1:]), p(b["out"]["mc_return"][1:]),
                              p(b["weight"]), 1.0, bench.GAMMA, p(loss), p(grad), None, p(stats), C.c_void_p(st.cuda_stream)))


K = 4


def body():
    if mode == "serial":
        for _ in range(K):
            gather(bufs[0], sa)
            tqc(bufs[0], sa)
    elif mode == "tqc":
        for _ in range(K):
            tqc(bufs[0], sa)
    elif mode in ("gather", "gather_lean"):
        for _ in range(K):
            gather(bufs[0], sa)
    else:  # K gathers on one stream next to K losses on the other (no dependencies: the co-run throughput itself)
        for _ in range(K):
            gather(bufs[1], sb)
            tqc(bufs[0], sa)


gather(bufs[0], sa)
gather(bufs[1], sa)
body()
torch.cuda.synchronize()
torch.cuda.profiler.start()
body()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", mode)
