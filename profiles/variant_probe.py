"""Probe: time the TQC loss kernel (bench shape) of a variant build of libfdql.so and fingerprint its outputs.
usage: python profiles/variant_probe.py <path to .so> [warps ...]"""
import ctypes as C
import hashlib
import sys

import torch

sys.path.insert(0, ".")
from fastdeepqlearning_b200 import _lib as L  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] != "-":
    L.LIB_PATH = sys.argv[1]
import fastdeepqlearning_b200 as pkg  # noqa: E402

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


for warps in [int(x) for x in (sys.argv[2:] or ["0", "16"])]:
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
    fp = hashlib.sha1(loss.cpu().numpy().tobytes() + grad.cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"{sys.argv[1] if len(sys.argv) > 1 else '-'} warps {warps or 'auto'}: {e0.elapsed_time(e1) / 50:.4f} ms  loss mean {float(loss.mean()):.6f} "
          f"fingerprint {fp}", flush=True)
