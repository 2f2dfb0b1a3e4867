"""Mirror of franQ/Replay/async_replay_memory.py:9-70.

The reference puts the ring in a child process behind two mp.Queue(3) and pickles every batch across.  The arena
already lives where the learner computes (HBM), so there is no process: this class keeps the constructor, `add`,
`temporal_sample`, `__len__` (saturating at maxlen, async_replay_memory.py:27-28, unlike the inner ring's maxlen-1)
and `batch_size` / `_temporal_len` / `pid` attributes that the wrappers and the Runner read."""
import os
import time

from .replay_memory import ReplayMemory, OversampleError


class AsyncReplayMemory:
    def __init__(self, maxlen, batch_size, temporal_len, **kwargs):
        self.batch_size = batch_size
        self._temporal_len = temporal_len
        self.temporal_len = temporal_len
        self._maxlen = int(maxlen)
        self._len = 0
        kwargs.pop("log_dir", None)
        self.block_on_oversample = kwargs.pop("block_on_oversample", False)
        self.replay = ReplayMemory(maxlen, batch_size, temporal_len, **kwargs)
        self.pid = os.getpid()

    def add(self, experience_dict):
        self._len = min(self._len + 1, self._maxlen)
        self.replay.add(experience_dict)

    def _count(self, n):
        self._len = min(self._len + int(n), self._maxlen)

    def add_rows(self, cols, episode_lengths=None, **kw):
        n = next(iter(cols.values())).shape[0]
        self._count(n)
        return self.replay.add_rows(cols, episode_lengths, **kw)

    def add_hindsight_rows(self, src_begins, lens, goal_rows, **kw):
        self._count(sum(int(x) for x in lens))
        return self.replay.add_hindsight_rows(src_begins, lens, goal_rows, **kw)

    def add_vmap_rows(self, cols, L, picks):
        self._count(L)
        return self.replay.add_vmap_rows(cols, L, picks)

    def reserve_rows(self, n):
        self._count(n)
        return self.replay.reserve_rows(n)

    def temporal_sample(self, *args, **kwargs):  # sample [Time, Batch, Experience]
        while True:
            try:
                return self.replay.temporal_sample(*args, **kwargs)
            except OversampleError:
                # the reference's child thread sleeps and retries, so its consumer blocks (async_replay_memory.py:55-61)
                if not self.block_on_oversample:
                    raise
                time.sleep(1)

    def sample(self, *args, **kwargs):
        return self.replay.sample(*args, **kwargs)

    def ready(self):
        n = len(self.replay)
        return n >= 2 * self._temporal_len and n >= self.batch_size

    def __getattr__(self, item):
        if "replay" in self.__dict__:
            return getattr(self.replay, item)
        raise AttributeError(item)

    def __getitem__(self, item):
        return self.replay[item]

    def __len__(self):
        return self._len
