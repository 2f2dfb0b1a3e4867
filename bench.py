#!/usr/bin/env python
"""bench.py -- sampled+relabelled+targeted transitions/s of the B200 learner hot path (BASELINE.json metric).

One step = one pass of the hot path over `batches_per_step` batches of 4096 sampled windows (T=2, one TD pair per window):
  fdql_sample_gather_draw  (uniform starts, hindsight flag p=0.8 = "future, k=4", goal row drawn in the kernel   [sampled]
                            + window gather + sample-time HER relabel + reward and return recompute;            [relabelled]
                            --separate-streams: fdql_sample_streams + fdql_sample_gather, two launches)
  fdql_tqc_loss        (pool 5x25 target atoms, sort, drop 10, soft target, quantile-Huber fwd+bwd,
                        n-step lower bound) on synthetic critic outputs resident in HBM             [targeted]
on a 1e7-row ring (obs 64, act 8, goal 16+16, 5 scalars; fp32).  `value` is device-timed with everything resident in
HBM; `e2e` is the same pass through the host-buffer C-ABI call fdql_hotpath_step_host (index streams and critic outputs
from pinned host memory, loss and dloss/dq back to host memory).  `--impl reference` times the CPU restatement of the
reference path (oracle/, numpy + torch-CPU, all host cores) on a bounded sample of the same workload.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 under torchrun (one rank per GPU, no data-path
collective: the replay shards by actor stream, scaling = weak)."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

OBS, ACT, GOAL, LEP = 64, 8, 16, 128
B, T, C_CRIT, Q_ATOMS, N_DROP = 4096, 2, 5, 25, 10
CQ = C_CRIT * Q_ATOMS
GAMMA, ALPHA, P_RELABEL = 0.99, 1.0, 0.8
ROW_BYTES = 4 * (OBS + ACT + 2 * GOAL) + 4 * 5  # 436 B, SURVEY.md section 8(d)

# algorithmic bytes per transition (T=2), SURVEY.md section 8(d) / DESIGN.md:
#   A gather: read 2 rows + 8 B start, write 2 rows                         = 4*436 + 8       = 1752
#   B relabel: goal source row 4*G + its scan record 16 + goal_row(8) + flag(1) = 64 + 25      =   89   (x p)
#   C return under relabel: n tail rows x one 16-byte scan record (64-bit goal hash, goal-agnostic reward, NaN flag);
#     mean tail of a uniform start in a 128-row episode = 64.5 rows         = 16 * 64.5       = 1032   (x p)
#     (SURVEY.md 8(d) counts 72*n B for this term, reading the goal vectors themselves: 4644 B.  The hash-assisted scan
#      is exact -- hash matches are verified on the full vectors -- and moves 4.5x fewer bytes; the roofline below uses
#      the bytes THIS algorithm needs, and `survey_bytes_per_transition` reports the SURVEY figure beside it.)
#   D TQC: next_z + q_pred + grad_q (3 x 4*125) + 4 scalars in + loss out   = 1500 + 20       = 1520
BYTES_GATHER = 4 * ROW_BYTES + 8
BYTES_RELABEL_SCAN = P_RELABEL * ((4 * GOAL + 25) + 16 * (LEP + 1) / 2)  # --tail-scan: 16 B scan record per tail row
BYTES_RELABEL = P_RELABEL * ((4 * GOAL + 25) + 16 * T)  # goal row + its link record + streams + one link record per window row
BYTES_RELABEL_SURVEY = P_RELABEL * ((4 * GOAL + 5) + 72 * (LEP + 1) / 2)
BYTES_TQC = 3 * 4 * CQ + 20
BYTES_STREAMS = 32 + 17  # read the start row's record sector, write start/flag/goal


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ring-rows", type=int, default=10_000_000)
    ap.add_argument("--batches-per-step", type=int, default=64)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-batches", type=int, default=0, help="bounded CPU-baseline sample, 0 = auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-small", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the T=50 figure (reference default temporal_len, SURVEY 8)")
    ap.add_argument("--no-updates", action="store_true")
    ap.add_argument("--exact-episode-step", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="learner step launched eagerly instead of as one CUDA graph")
    ap.add_argument("--separate-streams", action="store_true", help="draw the index / goal streams in their own launch (fdql_sample_streams)")
    ap.add_argument("--tail-scan", action="store_true", help="relabelled returns by scanning the episode tail instead of the link records")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: fall back to every sample seen
            sm = [float(x.split(",")[1]) for _, x in self.lines if x.count(",") >= 8] or [0.0]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------------
def build_ring(torch, pkg, Replay, n_rows, device, seed):
    """Synthetic bitflip-like replay: fixed-length episodes, goals in {0,1}^16, reward/done from R(ag, dg)."""
    n_eps = n_rows // LEP
    ring = Replay.ReplayMemory(n_eps * LEP + 1, B, T, device=device, seed=seed)
    ring.set_reward_op(pkg.RewardOp.bitflip(), GAMMA)
    gen = torch.Generator(device=device).manual_seed(seed)
    chunk_eps = 8192
    for e0 in range(0, n_eps, chunk_eps):
        ne = min(chunk_eps, n_eps - e0)
        n = ne * LEP
        ag = (torch.rand(n, GOAL, device=device, generator=gen) < 0.5).float()
        dg = (torch.rand(ne, GOAL, device=device, generator=gen) < 0.5).float().repeat_interleave(LEP, 0)
        hit = (ag == dg).all(-1, keepdim=True).float()
        step = torch.arange(n, device=device).remainder(LEP).float().unsqueeze(-1)
        cols = {"obs_1d": torch.randn(n, OBS, device=device, generator=gen),
                "action": torch.rand(n, ACT, device=device, generator=gen) * 2 - 1,
                "achieved_goal": ag, "desired_goal": dg, "reward": hit - 1, "task_done": hit,
                "episode_done": (step == LEP - 1).float(), "episode_step": step, "mc_return": torch.zeros(n, 1, device=device)}
        ring.add_rows(cols, episode_lengths=torch.full((ne,), LEP), with_returns=True)
    torch.cuda.synchronize(device)
    return ring


def run_ours(args):
    import torch
    import fastdeepqlearning_b200 as pkg
    from fastdeepqlearning_b200 import Replay, _lib as L

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    pkg.lib()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    from fastdeepqlearning_b200 import parallel
    numa_cpus = parallel.bind_to_gpu_numa_node(local) if world > 1 else None  # host buffers next to this rank's GPU
    dist = None
    if world > 1:
        import torch.distributed as dist
        # keep stdout to the one JSON line: NCCL's version / debug lines go to a file unless the caller chose one
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/fdql_nccl.%h.%p.log")
        dist.init_process_group("nccl", device_id=device)
    lib = pkg.lib()
    D = args.batches_per_step
    n = D * B
    M = (T - 1) * n
    ring = build_ring(torch, pkg, Replay, args.ring_rows, device, seed=1234 + rank)
    keys = ring.keys
    h = ring._h
    gen = torch.Generator(device=device).manual_seed(99 + rank)
    z = torch.randn(M, CQ, device=device, generator=gen) * 3
    q = torch.randn(M, CQ, device=device, generator=gen) * 3
    lp = torch.randn(M, device=device, generator=gen)
    out = {k: torch.empty((T, n, w), device=device) for k, w in zip(keys, ring._widths)}
    outp = L.ptr_array([out[k].data_ptr() for k in keys])
    aux_mask = torch.empty(T, n, device=device)
    aux_contig = torch.empty(T - 1, n, device=device)
    aux_weight = torch.empty(T - 1, n, device=device)
    starts = torch.empty(n, dtype=torch.int64, device=device)
    flags = torch.empty(n, dtype=torch.uint8, device=device)
    goals = torch.empty(n, dtype=torch.int64, device=device)
    loss = torch.empty(M, device=device)
    grad = torch.empty(M, CQ, device=device)
    stats = torch.zeros(4, dtype=torch.float64, device=device)
    params, n_params = ring.reward_op.c_params()
    opts = L.OPT_EMIT_LEARNER_AUX | (L.OPT_EXACT_EPISODE_STEP if args.exact_episode_step else 0)
    if args.tail_scan:
        lib.fdql_debug_force_generic_gather(16)  # tile kernel without the link records
    bytes_relabel = BYTES_RELABEL_SCAN if args.tail_scan else BYTES_RELABEL
    stream = torch.cuda.current_stream(device)
    sp = C.c_void_p(stream.cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    rlen = len(ring)
    counter = [0]

    def step(ev=None):
        if ev:
            ev[0].record(stream)
        if args.separate_streams:
            L.check(lib.fdql_sample_streams(h, n, T, L.GOAL_FUTURE, P_RELABEL, 7 + rank, counter[0], None, p(starts), p(flags), p(goals), sp))
        if ev:
            ev[1].record(stream)
        if args.separate_streams:
            L.check(lib.fdql_sample_gather(h, n, T, rlen, p(starts), p(flags), p(goals), ring.reward_op.op, params, n_params, GAMMA,
                                           opts, B, outp, p(aux_mask), p(aux_contig), p(aux_weight), sp))
        else:  # streams drawn inside the gather kernel: one launch
            L.check(lib.fdql_sample_gather_draw(h, n, T, L.GOAL_FUTURE, P_RELABEL, 7 + rank, counter[0], None, p(starts), p(flags), p(goals),
                                                ring.reward_op.op, params, n_params, GAMMA, opts, B, outp, p(aux_mask), p(aux_contig),
                                                p(aux_weight), sp))
        counter[0] += 1
        if ev:
            ev[2].record(stream)
        # the target reads reward / mask / mc_return of the NEXT row (t=1), quirk Q10
        L.check(lib.fdql_tqc_loss(M, CQ, N_DROP, p(z), p(q), p(lp), p(out["reward"][1:]), p(aux_mask[1:]),
                                  p(out["mc_return"][1:]), p(aux_weight), ALPHA, GAMMA, p(loss), p(grad), None, p(stats), sp))
        if ev:
            ev[3].record(stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(device)
    K = args.steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.25)
    if dist:
        dist.barrier()
    torch.cuda.synchronize(device)
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        step(evs[i])
    e1.record(stream)
    torch.cuda.synchronize(device)
    t1 = time.time()
    if dist:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    clk = clocks.stop(t0, t1)
    k_ms = np.array([[e[j].elapsed_time(e[j + 1]) for j in range(3)] for e in evs]).mean(0)  # streams, gather, tqc
    if dist:
        tmax = torch.tensor([ms_total], device=device)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
    ms_step = ms_total / K
    value = world * M / (ms_step * 1e-3)

    # ---- single-batch launches (B=4096 windows per launch): latency-bound figure, reported beside the headline ----
    def small_step(i):
        o = (i % D) * B
        outs = L.ptr_array([out[k].data_ptr() for k in keys])
        # outputs of a B-window launch are laid out [T, B, w] inside the first T*B rows of the big buffers
        L.check(lib.fdql_sample_gather_draw(h, B, T, L.GOAL_FUTURE, P_RELABEL, 7 + rank, 10_000 + i, None, p(starts[o:]), p(flags[o:]),
                                            p(goals[o:]), ring.reward_op.op, params, n_params, GAMMA, opts, B, outs, p(aux_mask),
                                            p(aux_contig), p(aux_weight), sp))
        L.check(lib.fdql_tqc_loss(B, CQ, N_DROP, p(z[o:]), p(q[o:]), p(lp[o:]), p(out["reward"].view(-1)[B:]),
                                  p(aux_mask.view(-1)[B:]), p(out["mc_return"].view(-1)[B:]), p(aux_weight), ALPHA, GAMMA,
                                  p(loss[o:]), p(grad[o:]), None, None, sp))
    small_ms = float("nan")
    if not args.no_small:
        for i in range(10):
            small_step(i)
        torch.cuda.synchronize(device)
        e0.record(stream)
        n_small = 200
        for i in range(n_small):
            small_step(i)
        e1.record(stream)
        torch.cuda.synchronize(device)
        small_ms = e0.elapsed_time(e1) / n_small

    # ---- the same per-batch launches as a CUDA graph of G batches round-robin over four streams (the learner's prefetch
    #      pattern: later batches are sampled/relabelled while earlier ones are targeted).  Draw counters live in device memory.
    graph_ms = float("nan")
    if not args.no_small:
        G = 16
        slots = []
        for i in range(G):
            so = {k: torch.empty((T, B, w), device=device) for k, w in zip(keys, ring._widths)}
            slots.append({"out": so, "outp": L.ptr_array([so[k].data_ptr() for k in keys]), "mask": torch.empty(T, B, device=device),
                          "contig": torch.empty(T - 1, B, device=device), "weight": torch.empty(T - 1, B, device=device),
                          "ctr": torch.zeros(2, dtype=torch.int64, device=device)})

        def graph_batch(i, st):
            sl, o, spx = slots[i], i * B, C.c_void_p(st.cuda_stream)
            L.check(lib.fdql_sample_gather_draw(h, B, T, L.GOAL_FUTURE, P_RELABEL, 1000 + 97 * rank + i, 0, p(sl["ctr"]), p(starts[o:]),
                                                p(flags[o:]), p(goals[o:]), ring.reward_op.op, params, n_params, GAMMA, opts, B,
                                                sl["outp"], p(sl["mask"]), p(sl["contig"]), p(sl["weight"]), spx))
            L.check(lib.fdql_tqc_loss(B, CQ, N_DROP, p(z[o:]), p(q[o:]), p(lp[o:]), p(sl["out"]["reward"][1:]), p(sl["mask"][1:]),
                                      p(sl["out"]["mc_return"][1:]), p(sl["weight"]), ALPHA, GAMMA, p(loss[o:]), p(grad[o:]), None,
                                      None, spx))
        NS = 4
        side = [torch.cuda.Stream(device) for _ in range(NS)]
        for i in range(G):  # warm-up outside capture
            graph_batch(i, stream)
        torch.cuda.synchronize(device)
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            cs = torch.cuda.current_stream(device)
            for st_ in side:
                st_.wait_stream(cs)
            for i in range(G):
                graph_batch(i, side[i % NS])
            for st_ in side:
                cs.wait_stream(st_)
        for _ in range(3):
            cg.replay()
        torch.cuda.synchronize(device)
        reps = 20
        first = starts[:B].clone()
        e0.record(stream)
        for _ in range(reps):
            cg.replay()
        e1.record(stream)
        torch.cuda.synchronize(device)
        assert not torch.equal(first, starts[:B]), "graph replays must draw fresh windows"
        graph_ms = e0.elapsed_time(e1) / (reps * G)

    # ---- secondary figure: the reference's default temporal_len = 50 (conf.py:38): 49 TD pairs per sampled window ------
    secondary = None
    if not args.no_secondary:
        T2, n2 = 50, 4 * B
        M2 = (T2 - 1) * n2
        out2 = {k: torch.empty((T2, n2, w), device=device) for k, w in zip(keys, ring._widths)}
        outp2 = L.ptr_array([out2[k].data_ptr() for k in keys])
        mask2, contig2, weight2 = (torch.empty(T2, n2, device=device), torch.empty(T2 - 1, n2, device=device),
                                   torch.empty(T2 - 1, n2, device=device))
        st2, fl2, go2 = (torch.empty(n2, dtype=torch.int64, device=device), torch.empty(n2, dtype=torch.uint8, device=device),
                         torch.empty(n2, dtype=torch.int64, device=device))
        z2 = torch.randn(M2, CQ, device=device, generator=gen) * 3
        q2 = torch.randn(M2, CQ, device=device, generator=gen) * 3
        lp2 = torch.randn(M2, device=device, generator=gen)
        loss2, grad2 = torch.empty(M2, device=device), torch.empty(M2, CQ, device=device)
        ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

        def step50(i, ev=None):
            if ev:
                ev[0].record(stream)
            L.check(lib.fdql_sample_streams(h, n2, T2, L.GOAL_FUTURE, P_RELABEL, 31 + rank, 50_000 + i, None, p(st2), p(fl2), p(go2), sp))
            L.check(lib.fdql_sample_gather(h, n2, T2, rlen, p(st2), p(fl2), p(go2), ring.reward_op.op, params, n_params, GAMMA,
                                           opts, B, outp2, p(mask2), p(contig2), p(weight2), sp))
            if ev:
                ev[1].record(stream)
            L.check(lib.fdql_tqc_loss(M2, CQ, N_DROP, p(z2), p(q2), p(lp2), p(out2["reward"][1:]), p(mask2[1:]),
                                      p(out2["mc_return"][1:]), p(weight2), ALPHA, GAMMA, p(loss2), p(grad2), None, None, sp))
            if ev:
                ev[2].record(stream)
        for i in range(3):
            step50(i)
        torch.cuda.synchronize(device)
        reps2, acc2 = 10, np.zeros(2)
        for i in range(reps2):
            step50(3 + i, ev2)
            torch.cuda.synchronize(device)
            acc2 += [ev2[0].elapsed_time(ev2[1]), ev2[1].elapsed_time(ev2[2])]
        ms2 = float(acc2.sum() / reps2)
        if dist:
            tm = torch.tensor([ms2], device=device)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms2 = float(tm.item())
        secondary = {"temporal_len": T2, "windows_per_step": n2, "transitions_per_step_per_gpu": M2, "ms_per_step": ms2,
                     "transitions_per_s": world * M2 / (ms2 * 1e-3), "rows_gathered_per_s": world * T2 * n2 / (acc2[0] / reps2 * 1e-3),
                     "gather_ms": float(acc2[0] / reps2), "tqc_ms": float(acc2[1] / reps2), "loss_mean": float(loss2.mean()),
                     "note": "reference default temporal_len (conf.py:38): one window gives 49 TD pairs, so the gather is amortised and "
                             "the loss kernel is the step; streams drawn by fdql_sample_streams, relabelled returns by the tail scan "
                             "(link records serve T <= 32)"}
        del out2, z2, q2, grad2, loss2, lp2

    # ---- e2e: the host-buffer C-ABI call, pinned host inputs, host outputs ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        rng = np.random.default_rng(5 + rank)
        s_np = rng.integers(0, rlen - T, n)
        tail = (LEP - 1) - (s_np % LEP)  # rows after the start row inside its episode
        g_np = np.minimum(s_np + np.where(tail > 0, 1 + (rng.random(n) * tail).astype(np.int64), 0), s_np - s_np % LEP + LEP - 1)
        hs, hg = torch.from_numpy(s_np).pin_memory(), torch.from_numpy(g_np).pin_memory()
        hf = torch.from_numpy((rng.random(n) < P_RELABEL).astype(np.uint8)).pin_memory()
        hz, hq, hlp = z.cpu().pin_memory(), q.cpu().pin_memory(), lp.cpu().pin_memory()
        hloss = torch.empty(M).pin_memory()
        hgrad = torch.empty(M, CQ).pin_memory()
        ph = lambda t: C.c_void_p(t.data_ptr())

        def host_step():
            L.check(lib.fdql_hotpath_step_host(h, n, T, rlen, ph(hs), ph(hf), ph(hg), ring.reward_op.op, params, n_params, GAMMA,
                                               opts & ~L.OPT_EMIT_LEARNER_AUX, outp, CQ, N_DROP, ph(hz), ph(hq), ph(hlp), ALPHA,
                                               ph(hloss), ph(hgrad), sp))
            stream.synchronize()  # the caller reads loss / grad from host memory after every step
        for _ in range(2):
            host_step()
        if dist:
            dist.barrier()
        torch.cuda.synchronize(device)
        tw0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.e2e_steps):
            host_step()
        e1.record(stream)
        torch.cuda.synchronize(device)
        tw1 = time.perf_counter()
        ms_e2e = max(e0.elapsed_time(e1), (tw1 - tw0) * 1e3) / args.e2e_steps  # host sync is part of the call
        if dist:
            tm = torch.tensor([ms_e2e], device=device)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms_e2e = float(tm.item())
        e2e = {"value": world * M / (ms_e2e * 1e-3), "unit": "transitions/s",
               "h2d_bytes_per_step": int(n * 17 + M * CQ * 8 + M * 4), "d2h_bytes_per_step": int(M * 4 + M * CQ * 4),
               "ms_per_step": ms_e2e, "api": "fdql_hotpath_step_host (pinned host streams + critic outputs in, loss + dloss/dq out)",
               "loss_mean": float(hloss.mean())}

    # ---- second metric of BASELINE.json: TQC updates/s = full learner steps (sample+relabel on this ring, PyTorch MLP
    #      forward/backward for 5x25 critics + actor, fused TQC loss, Adam, soft target update; gradient all-reduce if N>1)
    updates = None
    if not args.no_updates:
        import types
        from fastdeepqlearning_b200 import Agent
        from fastdeepqlearning_b200.Replay.wrappers import SampleTimeHindsight
        torch.manual_seed(0)
        lconf = Agent.LearnerConf(training_device=str(device), obs_space={"obs_1d": OBS, "achieved_goal": GOAL, "desired_goal": GOAL},
                                  action_space=types.SimpleNamespace(shape=(ACT,)), num_critics=C_CRIT, num_q_predictions=Q_ATOMS,
                                  top_quantiles_to_drop=N_DROP / CQ + 1e-9, batch_size=B, temporal_len=T, gamma=GAMMA,
                                  use_cuda_graph=not args.no_graph, graph_allreduce=True)
        learner = Agent.Learner(lconf, [SampleTimeHindsight(ring, relabel_prob=P_RELABEL)])
        for _ in range(5):
            learner.train_step()
        if dist:
            dist.barrier()
        torch.cuda.synchronize(device)
        n_upd = 30
        e0.record(stream)
        for _ in range(n_upd):
            last = learner.train_step()
        e1.record(stream)
        torch.cuda.synchronize(device)
        ms_upd = e0.elapsed_time(e1) / n_upd
        if dist:
            tm = torch.tensor([ms_upd], device=device)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms_upd = float(tm.item())
        updates = {"value": 1e3 / ms_upd, "unit": "updates/s (each rank steps on its own 4096-window batch, gradients averaged)",
                   "ms_per_update": ms_upd, "transitions_per_s": world * B * (T - 1) * 1e3 / ms_upd,
                   "params": int(sum(p.numel() for p in learner.params)), "loss": float(last),
                   "cuda_graph": bool(lconf.use_cuda_graph),
                   "note": "policy/critic MLPs are ordinary PyTorch fp32 modules; sample/relabel/target/loss are this repo's CUDA kernels; "
                           "the whole step (kernels + MLP fwd/bwd + gradient all-reduce + Adam + target update) is one captured CUDA graph"}

    # ---- roofline of the dominant kernel (by measured time) ------------------------------------------------------------
    peak, peak_src = peaks()
    kernels = {
        "sample_gather_kernel": {"ms": float(k_ms[1]), "bytes_per_transition": BYTES_GATHER + bytes_relabel + (0 if args.separate_streams else 17),
                                 "symbol": "fdql::sample_gather_tile_kernel<1, true>" + ("" if args.separate_streams else " (draws its own index / goal streams)"),
                                 "limiter": "HBM latency on random 32-256 B segments (ncu r1: long-scoreboard stalls dominate, DRAM traffic = algorithmic bytes)",
                                 "relabelled_returns": "tail scan (16 B per tail row)" if args.tail_scan else
                                 "link records: chain of equal achieved goals + goal-agnostic return, O(hits) per window"},
        "tqc_loss_kernel": {"ms": float(k_ms[2]), "bytes_per_transition": BYTES_TQC, "symbol": "fdql::tqc_loss_group_kernel<128, 7>",
                            "limiter": "instruction issue (81% active, ALU pipe 60%) and shared-memory wavefronts (79% of peak): 128-value sort "
                                       "network + 375 seven-level searches per transition; not HBM (ncu r1, profiles/r1_ncu_summary.md)"},
    }
    if args.separate_streams:
        kernels["sample_streams_kernel"] = {"ms": float(k_ms[0]), "bytes_per_transition": BYTES_STREAMS, "symbol": "fdql::sample_streams_kernel"}
    for kd in kernels.values():
        kd["achieved_gbs"] = kd["bytes_per_transition"] * M / (kd["ms"] * 1e-3) / 1e9
        kd["frac"] = kd["achieved_gbs"] / peak
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    total_bytes = sum(kd["bytes_per_transition"] for kd in kernels.values())
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kernels[dom]["bytes_per_transition"] * M, "launch_ms": kernels[dom]["ms"],
                "survey_bytes_per_transition": {"sample_gather_kernel": BYTES_GATHER + BYTES_RELABEL_SURVEY, "tqc_loss_kernel": BYTES_TQC},
                "kernels": kernels,
                "whole_step": {"bytes_per_transition": total_bytes, "achieved_gbs": total_bytes * M / (ms_step * 1e-3) / 1e9,
                               "frac": total_bytes * M / (ms_step * 1e-3) / 1e9 / peak}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            pj = json.load(open(prof))
            roofline["traffic"] = pj.get(dom)
            # second view for kernels that are not HBM-bound: warp instructions per launch (ncu, profiles/) against the SM issue rate
            # (148 SMs x 4 schedulers x 1 instruction per clock at the sampled SM clock), with this run's launch time
            sm_clk = (clk.get("sm_mhz") or 1965.0) * 1e6
            for name, kd in kernels.items():
                wi = pj.get("warp_instructions", {}).get(name)
                if wi and not args.tail_scan and not args.separate_streams:
                    kd["issue"] = {"warp_instructions_per_launch": wi, "achieved_ginst_s": wi / (kd["ms"] * 1e-3) / 1e9,
                                   "peak_ginst_s": 148 * 4 * sm_clk / 1e9, "frac": wi / (kd["ms"] * 1e-3) / (148 * 4 * sm_clk)}
        except Exception:
            pass

    line = {"metric": "sampled+relabelled+targeted transitions/s", "value": value, "unit": "transitions/s", "n_gpus": world,
            "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "ours",
            "config": {"workload": "HER(future,k=4: relabel p=0.8, return-to-go recomputed over the whole episode tail) + TQC 5x25 drop 10 + "
                                   "n-step lower bound; obs64/act8/goal16; ring %d rows/GPU (L=128 episodes); batch 4096, T=2"
                                   % (len(ring) + 1),
                       "batch": B, "temporal_len": T, "batches_per_step": D, "transitions_per_step_per_gpu": M,
                       "ring_rows_per_gpu": len(ring) + 1,
                       "l2": "inputs larger than L2 (random rows of a %.1f GB arena; %d MB of critic outputs per step)"
                             % ((len(ring) + 1) * (ROW_BYTES + 16) / 1e9, M * CQ * 8 // 2 ** 20),
                       "exact_episode_step": bool(args.exact_episode_step), "parallelism": f"replay shards x{world}, no data-path collective",
                       "numa_bound_cpus": len(numa_cpus) if numa_cpus else None},
            "roofline": roofline, "gpu_launches": (3 if args.separate_streams else 2) * K, "clocks": clk,
            "single_batch_launches": {"windows_per_launch": B, "ms_per_batch": small_ms, "transitions_per_s": world * B / (small_ms * 1e-3),
                                      "note": "2 launches per 4096-window batch from Python (gather with fused draw, loss), launch-latency bound",
                                      "cuda_graph_4_streams": {"ms_per_batch": graph_ms, "transitions_per_s": world * B / (graph_ms * 1e-3),
                                                               "note": "16 batches x 2 launches captured once, round-robin over four streams"}},
            "checks": {"loss_mean": float(loss.mean()), "relabel_frac": float(flags.float().mean()),
                       "violations": float(stats[2] / max(float(stats[3]), 1) / CQ)}}
    if secondary:
        line["secondary_T50"] = secondary
    if e2e:
        line["e2e"] = e2e
    if updates:
        line["tqc_updates"] = updates
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(args, steps=3, warmup=1, quiet=True)["cpu_baseline"]
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist:
        # NCCL kernels captured in the learner's CUDA graph must go before the communicator does; a watchdog ends the process
        # if the teardown still blocks (the JSON line is out by then)
        import gc
        import threading
        threading.Timer(20.0, lambda: os._exit(0)).start()
        if not args.no_updates:
            learner.close()
        learner = None  # noqa: F841
        gc.collect()
        torch.cuda.synchronize(device)
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


# ---------------------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_window_chunk(job):
    """One worker: window gather (numpy fancy index, replay_memory.py:62-70) + hindsight relabel + return recompute for a
    slice of the batch, by the oracle's restatement of her.py:55-95 / nstep_return.py:60-72."""
    from oracle import cpu_restatement as O
    lo, hi = job
    c = _CPU
    return O.sample_time_relabel(c["cols"], c["starts"][lo:hi], T, c["flags"][lo:hi], c["goals"][lo:hi], c["ep_start"],
                                 c["ep_end"], O.reward_bitflip, GAMMA)


class _LazyExtent:
    """ep_start / ep_end of a row for fixed-length episodes without materialising 1e7-entry tables per worker."""

    def __init__(self, last):
        self.last = last

    def __getitem__(self, row):
        base = (int(row) // LEP) * LEP
        return base + LEP - 1 if self.last else base


def cpu_reference(args, steps, warmup, quiet=False):
    """The reference's CPU replay-and-target path (oracle port) on the host cores, same workload, bounded sample."""
    import multiprocessing as mp
    import torch
    from oracle import cpu_restatement as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(0)
    n_rows = (args.ring_rows // LEP) * LEP
    ep_of_n = n_rows // LEP
    ag = (rng.random((n_rows, GOAL), dtype=np.float32) < 0.5).astype(np.float32)
    dg = np.repeat((rng.random((ep_of_n, GOAL), dtype=np.float32) < 0.5).astype(np.float32), LEP, axis=0)
    hit = (ag == dg).all(-1, keepdims=True)
    step = (np.arange(n_rows) % LEP).astype(np.float32).reshape(-1, 1)
    cols = {"obs_1d": rng.standard_normal((n_rows, OBS), dtype=np.float32),
            "action": rng.random((n_rows, ACT), dtype=np.float32) * 2 - 1, "achieved_goal": ag, "desired_goal": dg,
            "reward": hit.astype(np.float32) - 1, "task_done": hit.astype(np.float32),
            "episode_done": (step == LEP - 1).astype(np.float32), "episode_step": step}
    cols["mc_return"] = O.segmented_returns(cols["reward"][:LEP * 64], cols["episode_done"][:LEP * 64], GAMMA).reshape(-1, 1)
    cols["mc_return"] = np.resize(cols["mc_return"], (n_rows, 1))  # values only feed the lower bound; timing-equivalent
    nb = args.cpu_batches or 1
    n = nb * B
    _CPU.update(cols=cols, ep_start=_LazyExtent(False), ep_end=_LazyExtent(True))
    z = torch.randn(n, CQ) * 3
    q = torch.randn(n, CQ) * 3
    lp = torch.randn(n, 1)
    workers = min(cores, 32)
    ctx = mp.get_context("fork")

    def one_step(pool):
        s = rng.integers(0, n_rows - T, n)
        tail = (LEP - 1) - (s % LEP)
        g = s + np.where(tail > 0, 1 + (rng.random(n) * tail).astype(np.int64), 0)
        g = np.minimum(g, s - s % LEP + LEP - 1)
        f = rng.random(n) < P_RELABEL
        _CPU.update(starts=s, flags=f, goals=g)
        per = (n + workers - 1) // workers
        jobs = [(i, min(i + per, n)) for i in range(0, n, per)]
        parts = pool.map(_cpu_window_chunk, jobs) if pool else [_cpu_window_chunk(j) for j in jobs]
        batch = {k: np.concatenate([p[k] for p in parts], axis=1) for k in parts[0]}
        xp = {k: torch.from_numpy(v) for k, v in batch.items()}  # TorchDataLoader cast (already fp32)
        mask = 1.0 - xp["task_done"]
        qp = q.clone().requires_grad_(True)
        loss = O.tqc_q_loss_torch(qp.view(T - 1, n, CQ), z.view(T - 1, n, CQ), lp.view(T - 1, n, 1), xp["reward"][1:], mask[1:],
                                  xp["mc_return"][1:], ALPHA, GAMMA, N_DROP)
        loss.mean().backward()
        return float(loss.mean())

    # fork per step so the workers see this step's streams (the ring itself is shared copy-on-write)
    def timed(k):
        t = 0.0
        for _ in range(k):
            t0 = time.perf_counter()
            if workers > 1:
                # streams must exist before the fork: draw them inside one_step, so fork a fresh pool per step
                one_step_pool(one_step, ctx, workers)
            else:
                one_step(None)
            t += time.perf_counter() - t0
        return t

    def one_step_pool(fn, ctx_, w):
        class _P:
            def map(self, f, jobs):
                with ctx_.Pool(w) as pool:
                    return pool.map(f, jobs)
        return fn(_P())

    timed(warmup)
    tt = timed(steps)
    ms_step = tt / steps * 1e3
    value = (T - 1) * n / (ms_step * 1e-3)
    base = {"value": value, "unit": "transitions/s", "cores": cores, "kind": "port",
            "sample": f"{steps} step(s) of {nb} batch(es) x {B} windows (T={T}) on a {n_rows}-row numpy ring: fancy-index gather + "
                      f"oracle HER relabel/return recompute over {workers} forked workers, then the reference's torch-CPU TQC "
                      f"sort/target/[CQ x K] pairwise quantile-Huber fwd+bwd on {cores} threads"}
    line = {"metric": "sampled+relabelled+targeted transitions/s", "value": value, "unit": "transitions/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "HER(future,k=4: relabel p=0.8, return-to-go recomputed over the whole episode tail) + TQC 5x25 drop 10 + "
                                   "n-step lower bound; obs64/act8/goal16; ring %d rows/GPU (L=128 episodes); batch 4096, T=2" % (n_rows + 1),
                       "batch": B, "temporal_len": T, "batches_per_step": nb},
            "cpu_baseline": base, "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    return line


def main():
    args = parse()
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) != 0:
            return
        print(json.dumps(cpu_reference(args, steps=max(args.steps, 1), warmup=args.warmup)))
        return
    run_ours(args)


if __name__ == "__main__":
    main()
