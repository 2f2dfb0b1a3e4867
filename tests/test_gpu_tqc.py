"""GPU parity: TQC target + quantile-Huber loss / SAC min-target loss kernels vs the reference goldens
(tests/golden/tqc.npz, produced by running franQ's own q_loss / quantile_huber_loss_f) and vs the CPU oracle.
Tolerances (BASELINE.json north_star): 1e-5 relative for targets and losses; td_target is required bit-exact."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import cpu_restatement as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(fdql):
    from fastdeepqlearning_b200 import ops as _ops
    return _ops


def dev(x):
    import torch
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float32)), device="cuda")


def grad_close(got, want, scale_ref):
    # elementwise 1e-5 of the gradient's own scale (entries can cancel to ~0, so pure relative is meaningless there)
    tol = 1e-5 * max(float(np.abs(scale_ref).max()), 1e-30)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=tol)


def test_quantile_huber_goldens(ops):
    g = load_golden("tqc")
    for i in range(int(g["qh_cases"])):
        q, s = g[f"qh{i}_q"], g[f"qh{i}_s"]
        loss, grad = ops.quantile_huber(dev(q), dev(s))
        np.testing.assert_allclose(loss.cpu().numpy(), g[f"qh{i}_loss"], rtol=1e-5, atol=1e-7, err_msg=f"case {i}")
        grad_close(grad.cpu().numpy(), g[f"qh{i}_grad"], g[f"qh{i}_grad"])


def test_tqc_q_loss_goldens(ops):
    g = load_golden("tqc")
    for i in range(int(g["ql_cases"])):
        p = f"ql{i}_"
        ment, lb = bool(g[p + "ment"]), bool(g[p + "lb"])
        alpha, gamma, n_drop = float(g[p + "alpha"]), float(g[p + "gamma"]), int(g[p + "n_drop"])
        r = ops.tqc_loss(dev(g[p + "q_pred"]), dev(g[p + "next_z"]), dev(g[p + "log_pi"]) if ment else None,
                         dev(g[p + "reward"]), dev(g[p + "mask"]), dev(g[p + "mc_return"]) if lb else None, alpha, gamma,
                         n_drop, grad_scale=dev(g[p + "upstream"]), want_target=True, want_stats=True)
        np.testing.assert_allclose(r["loss"].cpu().numpy(), g[p + "loss"], rtol=1e-5, atol=1e-6, err_msg=p)
        grad_close(r["grad"].cpu().numpy(), g[p + "grad"], g[p + "grad"])
        td = O.tqc_td_target(g[p + "next_z"], g[p + "log_pi"], g[p + "reward"], g[p + "mask"], alpha, gamma, n_drop, ment)
        np.testing.assert_array_equal(r["td_target"].cpu().numpy(), td)  # bit-exact fp32 target
        st = r["stats"].cpu().numpy()
        M, n = np.prod(g[p + "q_pred"].shape[:-1]), g[p + "q_pred"].shape[-1]
        assert st[3] == M
        np.testing.assert_allclose(st[0] / (M * n), float(g[p + "q_pred_mu"]), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(st[1] / M, float(g[p + "q_pred_var"]), rtol=1e-4)
        if lb:
            np.testing.assert_allclose(st[2] / (M * n), float(g[p + "viol"]), rtol=1e-6)


def test_sac_goldens(ops):
    g = load_golden("tqc")
    for i in range(int(g["sac_cases"])):
        p = f"sac{i}_"
        ment, lb = bool(g[p + "ment"]), bool(g[p + "lb"])
        r = ops.sac_min_target_loss(dev(g[p + "q_pred"]), dev(g[p + "target_z"]), dev(g[p + "log_pi"]) if ment else None,
                                    dev(g[p + "reward"]), dev(g[p + "mask"]), dev(g[p + "mc_return"]) if lb else None,
                                    float(g[p + "alpha"]), float(g[p + "gamma"]), grad_scale=dev(g[p + "upstream"]),
                                    want_stats=True)
        np.testing.assert_allclose(r["loss"].cpu().numpy(), g[p + "loss"], rtol=1e-5, atol=1e-6)
        grad_close(r["grad"].cpu().numpy(), g[p + "grad"], g[p + "grad"])
        if lb:
            st = r["stats"].cpu().numpy()
            np.testing.assert_allclose(st[2] / g[p + "q_pred"].size, float(g[p + "viol"]), rtol=1e-6)


@pytest.mark.parametrize("n,n_drop", [(125, 10), (50, 10), (20, 4), (7, 1), (128, 1), (250, 20), (33, 3), (2, 1)])
@pytest.mark.parametrize("scale,offset", [(3.0, 0.0), (0.3, -500.0), (30.0, 50.0), (0.01, 1.0)])
def test_tqc_vs_oracle_random(ops, n, n_drop, scale, offset):
    rng = np.random.default_rng(n * 1000 + n_drop)
    M = 257
    z = (rng.standard_normal((M, n)) * scale + offset).astype(np.float32)
    q = (rng.standard_normal((M, n)) * scale + offset * 0.99 + rng.standard_normal((M, 1))).astype(np.float32)
    lp = rng.standard_normal((M, 1)).astype(np.float32)
    rew = rng.standard_normal((M, 1)).astype(np.float32)
    mask = (rng.random((M, 1)) > 0.2).astype(np.float32)
    mc = (q.mean(-1, keepdims=True) + rng.standard_normal((M, 1)) * scale).astype(np.float32)
    # the reference forms the target in fp32 (sort, slice, affine: bit-exact below) and the pairwise loss from it;
    # the oracle evaluates that loss by brute force in float64
    td = O.tqc_td_target(z, lp, rew, mask, 0.7, 0.99, n_drop)
    loss = O.quantile_huber(q, td)[..., None]
    grad = O.quantile_huber_grad(q, td)
    lb = np.maximum(mc.astype(np.float64) - q, 0)
    loss = loss + lb.mean(-1, keepdims=True)
    grad = grad - (lb > 0) / n
    r = ops.tqc_loss(dev(q), dev(z), dev(lp), dev(rew), dev(mask), dev(mc), 0.7, 0.99, n_drop, want_target=True)
    np.testing.assert_allclose(r["loss"].cpu().numpy(), loss, rtol=1e-5, atol=1e-6)
    grad_close(r["grad"].cpu().numpy(), grad, grad)
    np.testing.assert_array_equal(r["td_target"].cpu().numpy(), td)


def test_edge_cases(ops):
    import torch
    q = torch.zeros(4, 125, device="cuda")
    z = torch.zeros(4, 125, device="cuda")
    s = torch.zeros(4, 1, device="cuda")
    with pytest.raises(ValueError):  # quirk Q8: int(p*CQ) == 0 -> the reference's empty target; refused
        ops.tqc_loss(q, z, s, s, s + 1, s, 1.0, 0.99, 0)
    # ties everywhere (all atoms equal) and mask = 0 (all targets equal the reward)
    r = ops.tqc_loss(q, z, s, s + 2.0, s, None, 1.0, 0.99, 10)
    want, gwant, _ = O.tqc_q_loss(np.zeros((4, 125)), np.zeros((4, 125)), np.zeros((4, 1)), np.full((4, 1), 2.0),
                                  np.zeros((4, 1)), None, 1.0, 0.99, 10, use_lower_bound=False)
    np.testing.assert_allclose(r["loss"].cpu().numpy(), want, rtol=1e-5)
    grad_close(r["grad"].cpu().numpy(), gwant, gwant)
    # empty batch
    e = ops.tqc_loss(q[:0], z[:0], s[:0], s[:0], s[:0], s[:0], 1.0, 0.99, 10)
    assert e["loss"].shape == (0, 1)
    with pytest.raises(Exception):  # CPU tensors are refused: no fallback
        ops.tqc_loss(q.cpu(), z.cpu(), s.cpu(), s.cpu(), s.cpu(), s.cpu(), 1.0, 0.99, 10)


def test_full_size_properties(ops):
    """BASELINE.json size (4096 x 125, drop 10): size-independent properties instead of the brute-force oracle."""
    import torch
    gen = torch.Generator(device="cuda").manual_seed(1)
    M, n = 4096, 125
    z = torch.randn(M, n, device="cuda", generator=gen) * 3
    q = torch.randn(M, n, device="cuda", generator=gen) * 3
    lp, rew = torch.randn(M, 1, device="cuda", generator=gen), torch.randn(M, 1, device="cuda", generator=gen)
    mask = (torch.rand(M, 1, device="cuda", generator=gen) > 0.1).float()
    mc = torch.randn(M, 1, device="cuda", generator=gen)
    a = ops.tqc_loss(q, z, lp, rew, mask, mc, 1.0, 0.99, 10, want_target=True)
    # (1) the target does not depend on the order of the pooled atoms (sort)
    perm = torch.randperm(n, device="cuda", generator=gen)
    b = ops.tqc_loss(q, z[:, perm], lp, rew, mask, mc, 1.0, 0.99, 10, want_target=True)
    assert torch.equal(a["td_target"], b["td_target"]) and torch.equal(a["loss"], b["loss"]) and torch.equal(a["grad"], b["grad"])
    # (2) td_target is sorted and equals torch's sort/slice/affine bit for bit
    td = rew + mask * 0.99 * (torch.sort(z, -1)[0][:, :-10] + 1.0 * (-lp))
    assert torch.equal(a["td_target"], td)
    # (3) the fused gradient is the derivative of the fused loss (central differences on a few coordinates, fp64 oracle)
    sub = slice(0, 64)
    loss64, grad64, _ = O.tqc_q_loss(q[sub].cpu().numpy(), z[sub].cpu().numpy(), lp[sub].cpu().numpy(), rew[sub].cpu().numpy(),
                                     mask[sub].cpu().numpy(), mc[sub].cpu().numpy(), 1.0, 0.99, 10)
    np.testing.assert_allclose(a["loss"][sub].cpu().numpy(), loss64, rtol=1e-5, atol=1e-6)
    grad_close(a["grad"][sub].cpu().numpy(), grad64, grad64)
    # (4) the brute-force torch form of the reference on the whole batch
    ref = O.tqc_q_loss_torch(q.cpu(), z.cpu(), lp.cpu(), rew.cpu(), mask.cpu(), mc.cpu(), 1.0, 0.99, 10)
    np.testing.assert_allclose(a["loss"].cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)


def test_autograd_drop_in(ops):
    import torch
    g = load_golden("tqc")
    q = dev(g["qh0_q"]).requires_grad_(True)
    loss = ops.quantile_huber_loss_f(q, dev(g["qh0_s"]))
    assert loss.shape == g["qh0_loss"].shape
    loss.sum().backward()
    grad_close(q.grad.cpu().numpy(), g["qh0_grad"], g["qh0_grad"])
    p = "ql1_"
    qp = dev(g[p + "q_pred"]).requires_grad_(True)
    l = ops.tqc_q_loss_autograd(qp, dev(g[p + "next_z"]), dev(g[p + "log_pi"]), dev(g[p + "reward"]), dev(g[p + "mask"]),
                                dev(g[p + "mc_return"]) if bool(g[p + "lb"]) else None, float(g[p + "alpha"]),
                                float(g[p + "gamma"]), int(g[p + "n_drop"]))
    (l * dev(g[p + "upstream"])).sum().backward()
    grad_close(qp.grad.cpu().numpy(), g[p + "grad"], g[p + "grad"])


@pytest.mark.parametrize("n,k", [(100, 10), (128, 127), (125, 97), (70, 96), (97, 95), (64, 5), (33, 64), (20, 32), (7, 3)])
def test_quantile_huber_mismatched_widths_vs_oracle(ops, n, k):
    """quantile_huber_loss_f with n predicted quantiles against k samples, n != k: the table capacity follows the wider of
    the two, so the staged sample rows are shorter than the table (short rows take the fully bounds-checked row split).
    M = 1027 is not a multiple of the group size either, so the last group runs the plain staging path."""
    rng = np.random.default_rng(1000 * n + k)
    M = 1027
    q = (rng.standard_normal((M, n)) * 2).astype(np.float32)
    s = (rng.standard_normal((M, k)) * 2 + 0.5).astype(np.float32)
    loss, grad = ops.quantile_huber(dev(q), dev(s))
    want = O.quantile_huber(q[:64], s[:64])
    np.testing.assert_allclose(loss.cpu().numpy()[:64], want, rtol=1e-5, atol=1e-7)
    wg = O.quantile_huber_grad(q[:64], s[:64])
    grad_close(grad.cpu().numpy()[:64], wg, wg)
    # the whole batch against the brute-force pairwise form in torch (fp64 on the device)
    import torch
    qd, sd = dev(q).double(), dev(s).double()
    d = sd[:, None, :] - qd[:, :, None]
    tau = ((torch.arange(n, device="cuda", dtype=torch.float32) / n + 1 / 2 / n).double())[None, :, None]
    hub = torch.where(d.abs() > 1, d.abs() - 0.5, 0.5 * d * d)
    ref = ((tau - (d < 0).double()).abs() * hub).mean((-1, -2))
    np.testing.assert_allclose(loss.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("n,n_drop", [(125, 10), (50, 10), (30, 3)])
def test_unaligned_rows_take_the_plain_staging_path(ops, n, n_drop):
    """The group kernel stages rows with one bulk copy per group when the row run is 16-byte aligned; views that start in the middle of an
    allocation (row stride 4n bytes: 500 B for 125 atoms) use plain loads.  Same results bit for bit, stats included."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(n)
    M = 1031
    zf, qf = torch.randn(M + 1, n, device="cuda", generator=g) * 3, torch.randn(M + 1, n, device="cuda", generator=g) * 3
    lpf, rf, mcf = (torch.randn(M + 1, 1, device="cuda", generator=g) for _ in range(3))
    mkf = (torch.rand(M + 1, 1, device="cuda", generator=g) > 0.1).float()
    a = ops.tqc_loss(qf[1:], zf[1:], lpf[1:], rf[1:], mkf[1:], mcf[1:], 0.7, 0.99, n_drop, want_target=True, want_stats=True)
    assert qf[1:].data_ptr() % 16 != 0 or n % 4 == 0
    b = ops.tqc_loss(qf[1:].clone(), zf[1:].clone(), lpf[1:].clone(), rf[1:].clone(), mkf[1:].clone(), mcf[1:].clone(), 0.7, 0.99, n_drop,
                     want_target=True, want_stats=True)
    for k in ("loss", "grad", "td_target"):
        assert torch.equal(a[k], b[k]), k
    torch.testing.assert_close(a["stats"], b["stats"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("n,n_drop", [(125, 10), (50, 10), (20, 4)])
def test_group_kernel_vs_warp_per_transition_kernel(ops, fdql, n, n_drop):
    """Two independent implementations of the same loss (sub-warp sort + warp-wide search vs one warp per transition):
    bit-identical td_target, losses / gradients / summaries equal to rounding."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(n)
    M = 3001
    z, q = torch.randn(M, n, device="cuda", generator=g) * 3, torch.randn(M, n, device="cuda", generator=g) * 3 + 1
    lp, rw, mc = (torch.randn(M, 1, device="cuda", generator=g) for _ in range(3))
    mk = (torch.rand(M, 1, device="cuda", generator=g) > 0.1).float()
    a = ops.tqc_loss(q, z, lp, rw, mk, mc, 0.7, 0.99, n_drop, want_target=True, want_stats=True)
    lib = fdql.lib()
    old = lib.fdql_debug_tqc_warp_kernel(1)
    try:
        b = ops.tqc_loss(q, z, lp, rw, mk, mc, 0.7, 0.99, n_drop, want_target=True, want_stats=True)
    finally:
        lib.fdql_debug_tqc_warp_kernel(old)
    assert torch.equal(a["td_target"], b["td_target"])
    torch.testing.assert_close(a["loss"], b["loss"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(a["grad"], b["grad"], rtol=1e-4, atol=1e-5 * float(b["grad"].abs().max()))
    torch.testing.assert_close(a["stats"], b["stats"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("M", [1, 3, 4, 5, 63, 64, 2368, 9473, 70001])
def test_group_kernel_many_batch_sizes_repeatable(ops, fdql, M):
    """The staging pipeline of the group kernel (bulk copies on per-warp mbarriers whose parity flips every round, groups
    claimed from a per-block counter) over batch sizes from a single partial group to several rounds per warp: two runs are
    bit-identical, and equal to the warp-per-transition kernel (identical td_target, losses / gradients to rounding)."""
    import torch
    n, n_drop = 125, 10
    g = torch.Generator(device="cuda").manual_seed(M)
    z, q = torch.randn(M, n, device="cuda", generator=g) * 3, torch.randn(M, n, device="cuda", generator=g) * 3 + 1
    lp, rw, mc = (torch.randn(M, 1, device="cuda", generator=g) for _ in range(3))
    mk = (torch.rand(M, 1, device="cuda", generator=g) > 0.1).float()
    a = ops.tqc_loss(q, z, lp, rw, mk, mc, 0.7, 0.99, n_drop, want_target=True)
    b = ops.tqc_loss(q, z, lp, rw, mk, mc, 0.7, 0.99, n_drop, want_target=True)
    for k in ("loss", "grad", "td_target"):
        assert torch.equal(a[k], b[k]), k
    lib = fdql.lib()
    old = lib.fdql_debug_tqc_warp_kernel(1)
    try:
        c = ops.tqc_loss(q, z, lp, rw, mk, mc, 0.7, 0.99, n_drop, want_target=True)
    finally:
        lib.fdql_debug_tqc_warp_kernel(old)
    assert torch.equal(a["td_target"], c["td_target"])
    torch.testing.assert_close(a["loss"], c["loss"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(a["grad"], c["grad"], rtol=1e-4, atol=1e-5 * float(c["grad"].abs().max()))
