"""Import the unmodified reference (franQ) in THIS container -- TEST INFRASTRUCTURE.

The GPU box has no ``/root/reference``; nothing that runs there may call this module.
It exists so that ``oracle/make_goldens.py`` can execute the reference's own code to
produce ``tests/golden/*.npz`` and so that CPU-side tests can (optionally) re-validate
the restatement against the live reference when it is mounted.

Recipe: the reference imports ``gym`` and ``autoslot`` at
module import time only; ``jax`` (absent here) is really called by the vmap hindsight wrapper, so its stand-in implements
``jax.vmap`` / ``device_put`` / ``jnp.logical_*`` with numpy (``franQ/Env/wrappers/atari_wrappers.py:6``,
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
            # Functional stand-in for the four jax entry points her_vmap.py uses (jax.vmap, jax.devices, jax.device_put,
            # jnp.logical_*), so that the reference's OWN HindsightVmapWrite code can be executed for goldens.  vmap is a
            # Python loop over axis 0 (in_axes entries None = broadcast); device_put narrows 64-bit inputs to 32 bits like jax
            # with x64 disabled.  Nothing here restates reference logic.
            jnp = _stub("jax.numpy", logical_and=np.logical_and, logical_not=np.logical_not, logical_or=np.logical_or)

            def _vmap(f, in_axes=0):
                def mapped(*args):
                    axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
                    assert all(ax in (0, None) for ax in axes)
                    n = next(np.shape(a)[0] for a, ax in zip(args, axes) if ax is not None)
                    outs = [f(*[(a[i] if ax is not None else a) for a, ax in zip(args, axes)]) for i in range(n)]
                    if isinstance(outs[0], tuple):
                        return tuple(np.stack([np.asarray(o[k]) for o in outs]) for k in range(len(outs[0])))
                    return np.stack([np.asarray(o) for o in outs])
                return mapped

            def _device_put(x, device=None):
                x = np.asarray(x)
                if x.dtype == np.float64:
                    x = x.astype(np.float32)
                elif x.dtype == np.int64:
                    x = x.astype(np.int32)
                return x

            _stub("jax", numpy=jnp, vmap=_vmap, devices=lambda *a: ["cpu"], device_put=_device_put)
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
    from franQ.Replay.wrappers.her_vmap import HindsightVmapWrite, HindsightVmapRead
    from franQ.Replay.wrappers.nstep_return_vmap import NStepReturnVmap
    ns.HindsightVmapWrite, ns.HindsightVmapRead, ns.NStepReturnVmap = HindsightVmapWrite, HindsightVmapRead, NStepReturnVmap
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
