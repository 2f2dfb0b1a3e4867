"""Mirror of franQ/Replay/wrappers/nstep_return.py:8-72.

Rows of the episode in flight are held on the host (like the reference's deques); when the episode ends they go to the
ring in one batch and the return-to-go recurrence (nstep_return.py:60-72, fp64 step / fp32 store) runs on the device
inside fdql_commit_episodes, together with the episode extents that sample-time relabelling needs."""
import numpy as np

from .wrapper_base_class import ReplayMemoryWrapper


def stack_rows(rows, keys=None):
    keys = list(rows[0]) if keys is None else keys
    return {k: np.stack([np.asarray(r[k], dtype=np.float32).reshape(-1) for r in rows]) for k in keys}


class NStepReturn(ReplayMemoryWrapper):
    def __init__(self, replay_buffer, n_step, discount, reward_name="reward", return_name="mc_return",
                 done_name="episode_done"):
        ReplayMemoryWrapper.__init__(self, replay_buffer)
        if (reward_name, return_name, done_name) != ("reward", "mc_return", "episode_done"):
            raise NotImplementedError("the arena binds the roles by the reference's default key names")
        self.n_step, self.discount = n_step, discount
        self.reward_name, self.return_name, self.done_name = reward_name, return_name, done_name
        self._reset()

    def _reset(self):
        self.rows = []

    def add(self, experience):
        assert self.reward_name in experience
        self.rows.append(dict(experience))
        if experience[self.done_name]:
            rows = self.rows
            self._reset()
            self.add_rows(stack_rows(rows), episode_lengths=[len(rows)])

    # ---- episode-batched protocol used by the hindsight wrapper and by bulk loaders -------------------------
    def add_rows(self, cols, episode_lengths, **kw):
        lens = [int(x) for x in episode_lengths]
        if any(L > self.n_step for L in lens):
            # quirk Q3 (nstep_return.py:33-34,50-57): the oldest row is stored a second time with the n-step-truncated return
            return self._add_rows_q3(cols, lens)
        cols = dict(cols)
        n = next(iter(cols.values())).shape[0]
        if self.return_name not in cols:
            cols[self.return_name] = np.zeros((n, 1), np.float32)
        return self.replay_buffer.add_rows(cols, episode_lengths=lens, with_returns=True, gamma=self.discount)

    def _add_rows_q3(self, cols, lens):
        raise NotImplementedError("episodes longer than nStep_return_steps (quirk Q3 duplicate row) are not supported yet")

    def add_hindsight_rows(self, src_begins, lens, goal_rows, **kw):
        if any(int(L) > self.n_step for L in lens):
            raise NotImplementedError("episodes longer than nStep_return_steps (quirk Q3 duplicate row) are not supported yet")
        return self.replay_buffer.add_hindsight_rows(src_begins, lens, goal_rows, with_returns=True, gamma=self.discount)
