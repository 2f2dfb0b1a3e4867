"""Mirror of franQ/Replay/wrappers/torch_dataloader.py:11-50.

The reference runs a thread that samples numpy batches and copies them key by key to the GPU.  Batches of the
device ring are born on the GPU in the learner dtype, so this wrapper is a pass-through that keeps the interface:
`ready()`, `temporal_sample()`, `sample()` and the `ConfigurationError` for the wrong mode."""
import torch

from .wrapper_base_class import ReplayMemoryWrapper


class ConfigurationError(Exception):
    pass


class TorchDataLoader(ReplayMemoryWrapper):
    def __init__(self, replay_buffer, device="cuda:0", precision=torch.float32, use_temporal=True, vectorized=True):
        ReplayMemoryWrapper.__init__(self, replay_buffer)
        if precision != torch.float32:
            raise ConfigurationError("the device ring stores and serves float32 (conf.dtype of every reference preset)")
        self.device, self.precision, self._use_temporal, self.vectorized = device, precision, use_temporal, vectorized
        if not use_temporal:
            raise NotImplementedError("TODO: Add support for pre-fetching and batching non-temporal samples")

    def ready(self):
        n = len(self.replay_buffer)
        return n >= 2 * self.replay_buffer._temporal_len and n >= self.replay_buffer.batch_size

    def sample(self):
        if self._use_temporal:
            raise ConfigurationError("Incorrect Config! Unset `use_temporal` in init to support this feature")
        return self.replay_buffer.sample()

    def temporal_sample(self, *args, **kwargs):
        if not self._use_temporal:
            raise ConfigurationError("Incorrect config! Set `use_temporal` in init to support this feature")
        return self.replay_buffer.temporal_sample(*args, **kwargs)
