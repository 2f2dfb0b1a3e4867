"""Mirror of franQ/Replay/wrappers/squash_rewards.py:5-18 (Pohlen transform on write; only used without HER,
Replay/__init__.py:28).  One scalar per env step on the host side of the boundary, before the row is staged."""
import math

from .wrapper_base_class import ReplayMemoryWrapper


def _pohlen_transform(x, epsilon=1e-2, pow=0.5):
    x = float(x)
    return math.copysign(1.0, x) * (math.pow(abs(x) + 1, pow) - 1) * (x != 0) + epsilon * x


class SquashRewards(ReplayMemoryWrapper):
    def add(self, experience_dict):
        experience_dict["reward"] = _pohlen_transform(experience_dict["reward"])
        ReplayMemoryWrapper.add(self, experience_dict)
