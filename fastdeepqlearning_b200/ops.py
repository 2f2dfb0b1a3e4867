"""Thin torch-facing wrappers of the loss kernels (fdql_tqc_loss, fdql_quantile_huber, fdql_sac_min_target_loss).

Tensors must be CUDA fp32; leading dims are flattened to M transitions.  Nothing here computes on the host: every
function launches the CUDA kernel through the C ABI or raises."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _flat(t, width=None):
    if t is None:
        return None
    if t.device.type != "cuda":
        raise L.FdqlError("libfdql kernels need CUDA tensors (no CPU fallback)")
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.reshape(-1, width).contiguous() if width else t.reshape(-1).contiguous()


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def n_atoms_dropped(top_quantiles_to_drop, n_atoms):
    """distributional_soft_actor_critic.py:51-53: int(p * CQ) atoms are cut from the top."""
    return int(top_quantiles_to_drop * n_atoms)


def action_onehot(action, n_actions):
    """deepQlearning.py:206-210: eye(n)[action.long()] for a gathered [..., 1] action column -> [..., n] (fp32)."""
    lead = tuple(action.shape[:-1])
    a = _flat(action)
    out = torch.empty((a.numel(), int(n_actions)), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        L.check(L.lib().fdql_action_onehot(a.numel(), int(n_actions), _p(a), _p(out), None, _stream(a)))
    return out.reshape(lead + (int(n_actions),))


def tqc_loss(q_pred, next_z, next_log_pi, reward, mask, mc_return, alpha, gamma, n_drop, grad_scale=None,
             want_target=False, want_stats=False, want_grad=True):
    """DistributionalSoftActorCritic.q_loss from the critics' outputs onward (distributional_soft_actor_critic.py:50-82).

    q_pred, next_z: [..., CQ]; next_log_pi (None = no entropy term), reward, mask, mc_return (None = no lower bound),
    grad_scale (None = 1): [..., 1].  Returns dict(loss [..., 1], grad [..., CQ], td_target [..., K], stats)."""
    lead, n = tuple(q_pred.shape[:-1]), int(q_pred.shape[-1])
    q, z = _flat(q_pred, n), _flat(next_z, n)
    M = q.shape[0]
    with torch.cuda.device(q.device):
        loss = torch.empty(M, dtype=torch.float32, device=q.device)
        grad = torch.empty_like(q) if want_grad else None
        K = n - int(n_drop)
        td = torch.empty((M, max(K, 0)), dtype=torch.float32, device=q.device) if want_target else None
        stats = torch.zeros(4, dtype=torch.float64, device=q.device) if want_stats else None
        if torch.is_tensor(alpha):  # temperature kept on the device (no host sync; CUDA-graph safe)
            a_dev = alpha.detach().to(device=q.device, dtype=torch.float32).reshape(-1)
            L.check(L.lib().fdql_tqc_loss_dev_alpha(M, n, int(n_drop), _p(z), _p(q), _p(_flat(next_log_pi)), _p(_flat(reward)),
                                                    _p(_flat(mask)), _p(_flat(mc_return)), _p(_flat(grad_scale)), _p(a_dev),
                                                    float(gamma), _p(loss), _p(grad), _p(td), _p(stats), _stream(q)))
        else:
            L.check(L.lib().fdql_tqc_loss(M, n, int(n_drop), _p(z), _p(q), _p(_flat(next_log_pi)), _p(_flat(reward)),
                                          _p(_flat(mask)), _p(_flat(mc_return)), _p(_flat(grad_scale)), float(alpha), float(gamma),
                                          _p(loss), _p(grad), _p(td), _p(stats), _stream(q)))
    out = {"loss": loss.reshape(lead + (1,))}
    if want_grad:
        out["grad"] = grad.reshape(lead + (n,))
    if want_target:
        out["td_target"] = td.reshape(lead + (K,))
    if want_stats:
        out["stats"] = stats
    return out


def quantile_huber(quantiles, samples, grad_scale=None, want_grad=True):
    """quantile_huber_loss_f (distributional_soft_actor_critic.py:90-103): loss [...] and d loss / d quantiles."""
    lead, n, k = tuple(quantiles.shape[:-1]), int(quantiles.shape[-1]), int(samples.shape[-1])
    q, s = _flat(quantiles, n), _flat(samples, k)
    M = q.shape[0]
    with torch.cuda.device(q.device):
        loss = torch.empty(M, dtype=torch.float32, device=q.device)
        grad = torch.empty_like(q) if want_grad else None
        L.check(L.lib().fdql_quantile_huber(M, n, k, _p(q), _p(s), _p(_flat(grad_scale)), _p(loss), _p(grad), _stream(q)))
    return loss.reshape(lead), (grad.reshape(lead + (n,)) if want_grad else None)


def sac_min_target_loss(q_pred, target_z, next_log_pi, reward, mask, mc_return, alpha, gamma, grad_scale=None,
                        want_stats=False):
    """SoftActorCritic.q_loss from the critics' outputs onward (soft_actor_critic.py:63-99,134)."""
    lead, n = tuple(q_pred.shape[:-1]), int(q_pred.shape[-1])
    q, z = _flat(q_pred, n), _flat(target_z, n)
    M = q.shape[0]
    with torch.cuda.device(q.device):
        loss = torch.empty(M, dtype=torch.float32, device=q.device)
        grad = torch.empty_like(q)
        stats = torch.zeros(4, dtype=torch.float64, device=q.device) if want_stats else None
        if torch.is_tensor(alpha):  # temperature kept on the device (no host sync; CUDA-graph safe)
            a_dev = alpha.detach().to(device=q.device, dtype=torch.float32).reshape(-1)
            L.check(L.lib().fdql_sac_min_target_loss_dev_alpha(M, n, _p(z), _p(q), _p(_flat(next_log_pi)), _p(_flat(reward)),
                                                               _p(_flat(mask)), _p(_flat(mc_return)), _p(_flat(grad_scale)), _p(a_dev),
                                                               float(gamma), _p(loss), _p(grad), _p(stats), _stream(q)))
        else:
            L.check(L.lib().fdql_sac_min_target_loss(M, n, _p(z), _p(q), _p(_flat(next_log_pi)), _p(_flat(reward)), _p(_flat(mask)),
                                                     _p(_flat(mc_return)), _p(_flat(grad_scale)), float(alpha), float(gamma),
                                                     _p(loss), _p(grad), _p(stats), _stream(q)))
    out = {"loss": loss.reshape(lead + (1,)), "grad": grad.reshape(lead + (n,))}
    if want_stats:
        out["stats"] = stats
    return out


class _TqcLossFn(torch.autograd.Function):
    """loss = q_loss(q_pred; targets); backward multiplies the fused d loss/d q_pred by the incoming gradient."""

    @staticmethod
    def forward(ctx, q_pred, next_z, next_log_pi, reward, mask, mc_return, alpha, gamma, n_drop):
        r = tqc_loss(q_pred, next_z, next_log_pi, reward, mask, mc_return, alpha, gamma, n_drop)
        ctx.save_for_backward(r["grad"])
        return r["loss"]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g,) + (None,) * 8


class _QuantileHuberFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, quantiles, samples):
        loss, grad = quantile_huber(quantiles, samples)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g.unsqueeze(-1), None


class _SacLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_pred, target_z, next_log_pi, reward, mask, mc_return, alpha, gamma):
        r = sac_min_target_loss(q_pred, target_z, next_log_pi, reward, mask, mc_return, alpha, gamma)
        ctx.save_for_backward(r["grad"])
        return r["loss"]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g,) + (None,) * 7


def tqc_q_loss_autograd(q_pred, next_z, next_log_pi, reward, mask, mc_return, alpha, gamma, n_drop):
    return _TqcLossFn.apply(q_pred, next_z, next_log_pi, reward, mask, mc_return, alpha, gamma, n_drop)


def quantile_huber_loss_f(quantiles, samples):
    """Drop-in for franQ's quantile_huber_loss_f (differentiable w.r.t. `quantiles`; `samples` is the no-grad target)."""
    return _QuantileHuberFn.apply(quantiles, samples)


def sac_q_loss_autograd(q_pred, target_z, next_log_pi, reward, mask, mc_return, alpha, gamma):
    return _SacLossFn.apply(q_pred, target_z, next_log_pi, reward, mask, mc_return, alpha, gamma)
