"""Probe: where does the co-run slowdown come from?  Lean gather (FDQL_OPT_CORESIDENT) under the TQC loss with the gather's phases
switched off one at a time (GatherArgs.dbg: 1 = no wide-key phase, 2 = no scalar phase, 4 = FMA loop instead of the scalar phase).
usage: python profiles/overlap_probe2.py [ring_rows]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import fastdeepqlearning_b200 as pkg  # noqa: E402
from fastdeepqlearning_b200 import Replay, _lib as L  # noqa: E402

lib = pkg.lib()
dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
B, T, CQ = bench.B, bench.T, bench.CQ
n = 64 * B
M = n
ring = bench.build_ring(torch, pkg, Replay, rows, dev, seed=1)
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


NB = 3
bufs = [make_buf() for _ in range(NB)]
loss, grad = torch.empty(M, device=dev), torch.empty(M, CQ, device=dev)
stats = torch.zeros(4, dtype=torch.float64, device=dev)
ctr = [0]
gopt = [0]
sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
K = 40


def gather(b, st):
    L.check(lib.fdql_sample_gather_draw(h, n, T, L.GOAL_FUTURE, 0.8, 7, ctr[0], None, p(b["starts"]), p(b["flags"]), p(b["goals"]),
                                        ring.reward_op.op, params, n_params, bench.GAMMA, opts | gopt[0], B, b["outp"], p(b["mask"]),
                                        p(b["contig"]), p(b["weight"]), C.c_void_p(st.cuda_stream)))
    ctr[0] += 1


def tqc(b, st):
    L.check(lib.fdql_tqc_loss(M, CQ, 10, p(z), p(q), p(lp), p(b["out"]["reward"][1:]), p(b["mask"][1:]), p(b["out"]["mc_return"][1:]),
                              p(b["weight"]), 1.0, bench.GAMMA, p(loss), p(grad), None, p(stats), C.c_void_p(st.cuda_stream)))


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream())
    sa.wait_event(e0)
    sb.wait_event(e0)
    fn()
    ea, eb = torch.cuda.Event(), torch.cuda.Event()
    ea.record(sa)
    eb.record(sb)
    torch.cuda.current_stream().wait_event(ea)
    torch.cuda.current_stream().wait_event(eb)
    e1.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


def only_gather():
    for i in range(K):
        gather(bufs[0], sa)


def only_tqc():
    for i in range(K):
        tqc(bufs[0], sa)


def overlapped(nb=2):
    """stream A: tqc(k); stream B: gather(k+1).. tqc(k) waits for gather(k); gather(k+nb) waits for tqc(k) (buffer reuse)."""
    done_g = [None] * nb
    done_t = [None] * nb
    gather(bufs[0], sb)
    done_g[0] = torch.cuda.Event()
    done_g[0].record(sb)
    for k in range(K):
        cur = k % nb
        for j in range(1, nb):  # keep nb - 1 gathers ahead
            kk = k + j
            if kk < K and (j == nb - 1 or k == 0):
                nxt = kk % nb
                if done_t[nxt] is not None:
                    sb.wait_event(done_t[nxt])
                gather(bufs[nxt], sb)
                done_g[nxt] = torch.cuda.Event()
                done_g[nxt].record(sb)
        sa.wait_event(done_g[cur])
        tqc(bufs[cur], sa)
        done_t[cur] = torch.cuda.Event()
        done_t[cur].record(sa)


def free_running():
    """no dependencies at all: K gathers on one stream, K losses on the other (upper bound of what overlap can give)"""
    for i in range(K):
        gather(bufs[1], sb)
        tqc(bufs[0], sa)


lib.fdql_debug_force_generic_gather(0)
lib.fdql_set_coresident(0)
print(f"tile gather alone {timed(only_gather):.4f}  tqc alone (20 warps) {timed(only_tqc):.4f}", flush=True)
quick = len(sys.argv) > 2 and sys.argv[2] == "quick"
for w in ((16,) if quick else (16, 14, 12)):
    lib.fdql_debug_tqc_warp_kernel(w << 8)
    lib.fdql_set_coresident(1)
    gopt[0] = L.OPT_CORESIDENT
    tt = timed(only_tqc)
    for ctas in ((2,) if quick else (1, 2)):
        for dbg in ((0, 1, 2, 2 | 8, 2 | 16, 2 | 8 | 16) if quick else (0, 1, 2, 4)):
            lib.fdql_debug_force_generic_gather((ctas << 20) | (dbg << 6))
            ga = timed(only_gather)
            o2 = timed(lambda: overlapped(2))
            o3 = timed(lambda: overlapped(3)) if dbg == 0 else float("nan")
            fr = timed(free_running)
            print(f"tqc warps {w} (alone {tt:.4f}) lean blocks/SM {ctas} dbg {dbg}: gather alone {ga:.4f} overlapped2 {o2:.4f} overlapped3 {o3:.4f} "
                  f"free-running {fr:.4f}", flush=True)
lib.fdql_debug_force_generic_gather(0)
lib.fdql_debug_tqc_warp_kernel(0)
