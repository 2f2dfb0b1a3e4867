"""fdql_hotpath_step_host (the host-buffer C-ABI call, bench.py's e2e leg) vs the device-resident path on the same streams:
bit-exact gathered batch, loss and gradient identical to fdql_sample_gather + fdql_tqc_loss (same kernels, same order)."""
import ctypes as C

import numpy as np
import pytest

from test_gpu_replay import _synthetic, npy

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T", [(700, 2), (20000, 2), (5000, 3)])
def test_host_step_equals_device_path(fdql, B, T):
    import torch
    from fastdeepqlearning_b200 import Replay, ops, _lib as L
    rng = np.random.default_rng(B + T)
    cols, lengths, starts_ep, ends, ep_of = _synthetic(rng, 80, 130, 16)
    N = int(lengths.sum())
    ring = Replay.ReplayMemory(N + 3, 16, T)
    ring.set_reward_op(fdql.RewardOp.bitflip(), 0.98)
    ring.add_rows(cols, episode_lengths=lengths)
    starts = rng.integers(0, N - T, B)
    flags = (rng.random(B) < 0.8).astype(np.uint8)
    goal_rows = np.array([rng.integers(s, ends[ep_of[s]] + 1) for s in starts])
    CQ, n_drop, M = 50, 10, (T - 1) * B
    g = torch.Generator().manual_seed(1)
    z, q, lp = torch.randn(M, CQ, generator=g) * 3, torch.randn(M, CQ, generator=g) * 3, torch.randn(M, generator=g)

    # device path
    dev = ring.temporal_sample(starts=starts, flags=flags, goal_rows=goal_rows, aux=True, length=N)
    want = ops.tqc_loss(q.cuda().view(T - 1, B, CQ), z.cuda().view(T - 1, B, CQ), lp.cuda().view(T - 1, B, 1), dev["reward"][1:],
                        dev["mask"][1:], dev["mc_return"][1:], 0.7, 0.98, n_drop, grad_scale=dev["loss_weight"])
    dev = {k: v.clone() for k, v in dev.items()}

    # host-buffer path
    hs, hg, hf = torch.from_numpy(starts).pin_memory(), torch.from_numpy(goal_rows).pin_memory(), torch.from_numpy(flags).pin_memory()
    hz, hq, hlp = z.pin_memory(), q.pin_memory(), lp.pin_memory()
    hloss, hgrad = torch.empty(M).pin_memory(), torch.empty(M, CQ).pin_memory()
    out = {k: torch.empty((T, B, w), device="cuda") for k, w in zip(ring._keys, ring._widths)}
    outp = L.ptr_array([out[k].data_ptr() for k in ring._keys])
    params, n_params = ring.reward_op.c_params()
    ph = lambda t: C.c_void_p(t.data_ptr())
    stream = torch.cuda.current_stream()
    for _ in range(2):  # the second call reuses the staging and the events
        L.check(fdql.lib().fdql_hotpath_step_host(ring._h, B, T, N, ph(hs), ph(hf), ph(hg), ring.reward_op.op, params, n_params, 0.98, 0,
                                                  outp, CQ, n_drop, ph(hz), ph(hq), ph(hlp), 0.7, ph(hloss), ph(hgrad),
                                                  C.c_void_p(stream.cuda_stream)))
        stream.synchronize()
        for k in ring._keys:
            np.testing.assert_array_equal(npy(out[k]), npy(dev[k]), err_msg=k)
        np.testing.assert_array_equal(hloss.numpy(), npy(want["loss"]).reshape(-1))
        np.testing.assert_array_equal(hgrad.numpy(), npy(want["grad"]).reshape(M, CQ))
        hloss.zero_()
        hgrad.zero_()
