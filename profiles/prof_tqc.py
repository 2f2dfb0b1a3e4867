"""Launch the TQC loss kernel alone at the bench's shape (262144 transitions x 125 atoms, drop 10) for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastdeepqlearning_b200 import ops

M, n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144, 125
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.randn(M, n, device="cuda", generator=g) * 3
q = torch.randn(M, n, device="cuda", generator=g) * 3
lp = torch.randn(M, 1, device="cuda", generator=g)
r = -torch.rand(M, 1, device="cuda", generator=g)
mask = (torch.rand(M, 1, device="cuda", generator=g) > 0.05).float()
G = torch.randn(M, 1, device="cuda", generator=g) - 8
gs = torch.rand(M, 1, device="cuda", generator=g)
for it in range(3):
    out = ops.tqc_loss(q, z, lp, r, mask, G, 1.0, 0.99, 10, grad_scale=gs, want_stats=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(20):
    out = ops.tqc_loss(q, z, lp, r, mask, G, 1.0, 0.99, 10, grad_scale=gs, want_stats=True)
e1.record()
torch.cuda.synchronize()
print("ms per launch", e0.elapsed_time(e1) / 20, "loss", float(out["loss"].mean()))
