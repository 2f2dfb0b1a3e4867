"""Device reward functors R(achieved_goal, goal) -> (reward, done).

franQ's HER wrapper receives an arbitrary Python callable (franQ/Replay/wrappers/her.py:13,62,67).  On the device
that is an enum-dispatched op (include/fdql.h FDQL_REWARD_*); this class is the value the mirror wrappers accept in
the `compute_reward` slot.  A plain Python callable is refused -- there is no host relabelling path."""
from __future__ import annotations

from . import _lib as L


class RewardOp:
    def __init__(self, op: int, params=()):
        self.op = int(op)
        self.params = [float(p) for p in params]

    @classmethod
    def bitflip(cls):
        """franQ/Env/bitflip.py:143-152: all(ag==dg) ? 0 : -1, done = reward==0"""
        return cls(L.REWARD_BITFLIP)

    @classmethod
    def all_geq(cls):
        """franQ/Env/classic_control_goal/classic_goal.py:88-93"""
        return cls(L.REWARD_ALL_GEQ)

    @classmethod
    def first_geq(cls):
        """franQ/Env/classic_control_goal/classic_goal.py:306-311"""
        return cls(L.REWARD_FIRST_GEQ)

    @classmethod
    def weighted_pnorm(cls, weights, success_threshold, p=0.5):
        """franQ/Env/eleurent_parking.py:42-55: -(sum |ag-dg|*w)^p, done = reward > -threshold"""
        return cls(L.REWARD_WEIGHTED_PNORM, [p, success_threshold, *list(weights)])

    @classmethod
    def coerce(cls, obj):
        if isinstance(obj, cls):
            return obj
        if isinstance(obj, str):
            return {"bitflip": cls.bitflip, "all_geq": cls.all_geq, "first_geq": cls.first_geq}[obj]()
        op = getattr(obj, "fdql_reward_op", None)
        if isinstance(op, cls):
            return op
        raise TypeError("compute_reward must be a fastdeepqlearning_b200.RewardOp (device functor); arbitrary Python "
                        "callables cannot run inside the CUDA relabel kernels and there is no host fallback")

    def c_params(self):
        return L.f32_array(self.params)
