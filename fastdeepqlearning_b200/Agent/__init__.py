"""Mirror of the loss half of franQ.Agent: the actor-critic components whose `q_loss` runs on the fused CUDA kernels, and a
learner that keeps DeepQLearning's `get_losses` / `train_step` contract.  The MLPs are ordinary PyTorch modules."""
from . import components
from .components.soft_actor_critic import SoftActorCritic
from .components.distributional_soft_actor_critic import DistributionalSoftActorCritic, quantile_huber_loss_f
from .learner import Learner, LearnerConf


def make(conf, replays=None):
    """Agent/__init__.py:4-15 builds DeepQLearning(conf); here the learner that owns the hot path."""
    return Learner(conf, replays)
