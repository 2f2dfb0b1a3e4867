"""Import the unmodified reference (franQ) in THIS container -- TEST INFRASTRUCTURE.

The GPU box has no ``/root/reference``; nothing that runs there may call this module.
It exists so that ``oracle/make_goldens.py`` can execute the reference's own code to
produce ``tests/golden/*.npz`` and so that CPU-side tests can (optionally) re-validate
the restatement against the live reference when it is mounted.

Recipe: the reference imports ``gym``, ``autoslot`` and ``jax`` at
module import time only (``franQ/Env/wrappers/atari_wrappers.py:6``,
``franQ/Replay/wrappers/wrapper_base_class.py:2``, ``franQ/Replay/wrappers/her_vmap.py:2,7``)
and uses the removed ``np.product`` (``franQ/Agent/components/soft_actor_critic.py:42``).
Empty stand-in modules + one alias make the whole package importable; no reference
source is modified or copied.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FDQL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "franQ"))


def _stub(name, **attrs):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def install_stubs():
    import numpy as np

    class _Dummy:  # stands in for gym.spaces.* / gym.Env / gym.Wrapper
        def __init__(self, *a, **k):
            pass

    for missing in ("gym", "autoslot", "jax"):
        try:
            importlib.import_module(missing)
            continue
        except Exception:
            pass
        if missing == "gym":
            spaces = _stub("gym.spaces", Dict=type("Dict", (_Dummy,), {}), Discrete=type("Discrete", (_Dummy,), {}),
                           Box=type("Box", (_Dummy,), {}), MultiBinary=type("MultiBinary", (_Dummy,), {}),
                           Space=type("Space", (_Dummy,), {}))
            wrappers = _stub("gym.wrappers", TimeLimit=type("TimeLimit", (_Dummy,), {}))
            _stub("gym", spaces=spaces, wrappers=wrappers, Env=type("Env", (_Dummy,), {}),
                  Wrapper=type("Wrapper", (_Dummy,), {}), ObservationWrapper=type("ObservationWrapper", (_Dummy,), {}),
                  ActionWrapper=type("ActionWrapper", (_Dummy,), {}), RewardWrapper=type("RewardWrapper", (_Dummy,), {}),
                  GoalEnv=type("GoalEnv", (_Dummy,), {}), make=lambda *a, **k: None)
        elif missing == "autoslot":
            _stub("autoslot", Slots=type("Slots", (), {}))
        elif missing == "jax":
            jnp = _stub("jax.numpy")
            _stub("jax", numpy=jnp)
    if not hasattr(np, "product"):
        np.product = np.prod  # numpy>=2 dropped the alias the reference still calls
    if not hasattr(np, "bool"):
        np.bool = bool


def load_reference():
    """Returns a namespace with the reference classes on the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # /root/reference is read-only
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    from franQ.Replay.replay_memory import ReplayMemory, OversampleError
    from franQ.Replay.wrappers import NStepReturn, HindsightNStepReplay
    from franQ.Replay.wrappers.squash_rewards import SquashRewards, _pohlen_transform
    from franQ.Replay.wrappers import nstep_return as nstep_mod
    from franQ.Agent.components.distributional_soft_actor_critic import (
        DistributionalSoftActorCritic, quantile_huber_loss_f)
    from franQ.Agent.components.soft_actor_critic import SoftActorCritic
    from franQ.Agent.deepQlearning import DeepQLearning
    from franQ.Agent.conf import AgentConf
    from franQ.common_utils import AttrDict
    ns.ReplayMemory, ns.OversampleError = ReplayMemory, OversampleError
    ns.NStepReturn, ns.HindsightNStepReplay = NStepReturn, HindsightNStepReplay
    ns.SquashRewards, ns.pohlen_transform = SquashRewards, _pohlen_transform
    ns.calculate_montecarlo_return = nstep_mod.calculate_montecarlo_return
    ns.DistributionalSoftActorCritic = DistributionalSoftActorCritic
    ns.SoftActorCritic = SoftActorCritic
    ns.quantile_huber_loss_f = quantile_huber_loss_f
    ns.DeepQLearning, ns.AgentConf, ns.AttrDict = DeepQLearning, AgentConf, AttrDict
    return ns
