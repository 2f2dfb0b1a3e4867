from . import models, soft_actor_critic, distributional_soft_actor_critic
