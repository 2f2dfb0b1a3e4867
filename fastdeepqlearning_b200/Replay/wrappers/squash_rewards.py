"""Mirror of franQ/Replay/wrappers/squash_rewards.py:5-18 (Pohlen transform on write; only used without HER,
Replay/__init__.py:28).  The transform runs on the device: the wrapper switches it on for the ring it sits on and every row that
enters the arena afterwards -- one dict at a time through add() (FDQL_APPEND_SQUASH_REWARDS of fdql_arena_append_packed_host) or in
episode batches from an NStepReturn underneath -- has its reward column transformed on the way in, so the returns the device
computes at episode commit are returns of squashed rewards, as in the reference's SquashRewards(NStepReturn(...)) stack.
`_pohlen_transform` keeps the reference's signature (scalars and arrays) for callers that use it directly."""
import math

import numpy as np

from .wrapper_base_class import ReplayMemoryWrapper


def _pohlen_transform(x, epsilon=1e-2, pow=0.5):
    if isinstance(x, (int, float)):
        return math.copysign(1.0, x) * (math.pow(abs(x) + 1, pow) - 1) * (x != 0) + epsilon * x
    x = np.asarray(x)
    return np.sign(x) * (np.power(np.abs(x) + 1, pow) - 1) + epsilon * x


class SquashRewards(ReplayMemoryWrapper):
    def __init__(self, replay_buffer):
        ReplayMemoryWrapper.__init__(self, replay_buffer)
        self.enable_squash_rewards()  # reaches the ring through the wrapper chain
