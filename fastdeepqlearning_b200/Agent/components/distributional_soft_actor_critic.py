"""Mirror of franQ/Agent/components/distributional_soft_actor_critic.py (TQC, arXiv:2005.04269).

`q_loss(curr_xp, next_xp)` keeps the reference signature and return triple; the part after the MLP forward passes
(pool + sort the ensemble's atoms, drop the top int(p*CQ), soft target, quantile-Huber over the [CQ x K] pairs, n-step
lower bound, summaries, and d loss/d q_pred) is ONE launch of fdql_tqc_loss.  `quantile_huber_loss_f` is the drop-in for
the reference's free function."""
import torch

from .soft_actor_critic import SoftActorCritic, _LossWithStats, _ReducedLossWithStats, _summaries_from_stats
from ... import ops

quantile_huber_loss_f = ops.quantile_huber_loss_f


class DistributionalSoftActorCritic(SoftActorCritic):
    def n_drop(self, n_atoms):
        """int(conf.top_quantiles_to_drop * CQ) (distributional_soft_actor_critic.py:51-53); 0 is the reference's empty
        target (quirk Q8) and is refused by the kernel."""
        return ops.n_atoms_dropped(self.conf.top_quantiles_to_drop, n_atoms)

    def q_loss(self, curr_xp, next_xp, grad_scale=None):
        conf = self.conf
        if conf.use_nStep_lowerbounds and getattr(conf, "use_bootstrap_minibatch_nstep", False):
            raise NotImplementedError("Need to update this from SAC to use the quantile huber loss!")  # reference :84-85
        q_pred, next_z, next_log_pi = self._critic_io(curr_xp, next_xp)
        lb = next_xp["mc_return"] if conf.use_nStep_lowerbounds else None
        lp = next_log_pi if conf.use_max_entropy_q else None
        loss, stats = _LossWithStats.apply(q_pred, lambda q: self._q_loss_fused(q, next_z, lp, next_xp, lb, grad_scale))
        return loss, None, _summaries_from_stats(stats, q_pred.shape[-1], lb is not None)

    def _q_loss_fused(self, q_pred, next_z, lp, next_xp, lb, grad_scale):
        return ops.tqc_loss(q_pred, next_z, lp, next_xp["reward"], next_xp["mask"], lb, self.curr_alpha, self.conf.gamma,
                            self.n_drop(q_pred.shape[-1]), grad_scale=grad_scale, want_stats=True)
