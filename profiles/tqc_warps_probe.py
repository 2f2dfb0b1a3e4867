"""Probe: TQC group kernel time at M = 262144 for different warps per block (fdql_debug_tqc_warp_kernel bits 8..15)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
import fastdeepqlearning_b200 as pkg  # noqa: E402
from fastdeepqlearning_b200 import _lib as L  # noqa: E402

lib = pkg.lib()
dev = torch.device("cuda:0")
M, CQ = 262144, 125
g = torch.Generator(device=dev).manual_seed(0)
z = torch.randn(M, CQ, device=dev, generator=g) * 3
q = torch.randn(M, CQ, device=dev, generator=g) * 3
lp, rew, mc = (torch.randn(M, device=dev, generator=g) for _ in range(3))
mask = (torch.rand(M, device=dev, generator=g) > 0.1).float()
w = torch.rand(M, device=dev, generator=g)
loss, grad = torch.empty(M, device=dev), torch.empty(M, CQ, device=dev)
stats = torch.zeros(4, dtype=torch.float64, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run():
    L.check(lib.fdql_tqc_loss(M, CQ, 10, p(z), p(q), p(lp), p(rew), p(mask), p(mc), p(w), 1.0, 0.99, p(loss), p(grad), None, p(stats), sp))


ref = None
for warps in [int(x) for x in (sys.argv[1:] or ["16", "18", "20", "12"])]:
    lib.fdql_debug_tqc_warp_kernel(warps << 8)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        run()
    e1.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = (loss.clone(), grad.clone())
    same = torch.equal(ref[0], loss) and torch.equal(ref[1], grad)
    print(f"warps {warps}: {e0.elapsed_time(e1) / 50:.4f} ms  identical to first: {same}", flush=True)
