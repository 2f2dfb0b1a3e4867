"""The two places where franQ's Runner / async trainer meet the learner hot path:

  ReplayHandler   mirror of Runner._replay_handler (franQ/Runner/runner.py:177-191): one thread per actor stream drains a queue
                  of experience dicts into that stream's write head (`Replay.make(...)[1][idx].add(xp)`), so the wrappers'
                  episode-batched device flushes run off the environment threads.  `put()` deep-copies like runner.py:161.
  ParamPublisher  mirror of DeepQLearning._push_params / _pull_params (franQ/Agent/deepQlearning.py:136-148): the trainer
                  publishes its weights every `param_update_interval` steps and the inference process loads the newest set.  The
                  reference moves a whole state_dict through `.to("cpu")` (a synchronous copy per tensor) and an mp.Queue(1); here
                  the weights go into one of two pinned host buffers by asynchronous copies on a side stream, and `pull()` returns
                  the newest buffer whose copies have completed -- the training stream never waits for the host.

Everything else of the Runner (environment handlers, evaluator, ranker) is unchanged host code and stays out of scope."""
import copy
import queue
import threading

import torch


class ReplayHandler:
    def __init__(self, write_heads, use_HER=False, maxsize=0):
        self.write_heads = list(write_heads)
        self.use_HER = bool(use_HER)
        self._queues = [queue.Queue(maxsize) for _ in self.write_heads]
        self._errors = []
        self._threads = [threading.Thread(target=self._loop, args=(i,), daemon=True, name=f"fdql-replay-{i}")
                         for i in range(len(self.write_heads))]
        for t in self._threads:
            t.start()

    def put(self, idx, xp):
        """runner.py:159-162: the experience dict is copied before it changes hands ("ensure no mutation between threads")."""
        if self._errors:
            raise self._errors[0]
        self._queues[idx].put(copy.deepcopy(xp))

    def _loop(self, idx):
        q, head = self._queues[idx], self.write_heads[idx]
        while True:
            xp = q.get()
            try:
                if xp is None:
                    return
                if not self.use_HER:
                    xp.pop("info", None)  # runner.py:185-186
                head.add(xp)
            except BaseException as e:  # surfaced by the next put() / join()
                self._errors.append(e)
            finally:
                q.task_done()

    def pending(self):
        """Rows queued or being stored right now."""
        return sum(q.unfinished_tasks for q in self._queues)

    def join(self):
        """Block until every queued row has reached its write head."""
        for q in self._queues:
            q.join()
        if self._errors:
            raise self._errors[0]

    def close(self):
        for q in self._queues:
            q.put(None)
        for t in self._threads:
            t.join(timeout=10)


class ParamPublisher:
    def __init__(self, modules, device=None, interval=1):
        """`modules`: dict name -> nn.Module whose weights the inference side needs (franQ pushes the whole agent; the actor and the
        encoder are what `act()` reads, deepQlearning.py:166-168)."""
        self.modules = dict(modules)
        self.interval = max(int(interval), 1)
        first = next(iter(next(iter(self.modules.values())).parameters()))
        self.device = torch.device(device) if device is not None else first.device
        self._stream = torch.cuda.Stream(self.device)
        self._src = {f"{m}.{k}": v for m, mod in self.modules.items() for k, v in mod.state_dict(keep_vars=True).items()}
        self._bufs = [{k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in self._src.items()} for _ in range(2)]
        self._events = [None, None]
        self._version = [0, 0]
        self._next = 0
        self._pushed = 0
        self._lock = threading.Lock()

    def maybe_push(self, train_step):
        """deepQlearning.py:82-93: every `param_update_interval` steps the training loop signals the push thread."""
        if train_step % self.interval == 0:
            self.push()
            return True
        return False

    def push(self):
        """Snapshot the weights as they are at this point of the CURRENT stream: the copies are ordered after the work already
        enqueued there (the optimizer step) and run on the publisher's stream, so later training work is not held up."""
        with self._lock:
            i = self._next
            cur = torch.cuda.current_stream(self.device)
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream), torch.no_grad():
                if self._events[i] is not None:
                    self._events[i].synchronize()  # a reader may still be copying out of this buffer's previous version
                for k, v in self._src.items():
                    self._bufs[i][k].copy_(v.detach(), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._stream)
            # the weights must not be overwritten by the next optimizer step before the copies have read them
            cur.wait_stream(self._stream)
            self._pushed += 1
            self._events[i], self._version[i] = ev, self._pushed
            self._next = i ^ 1

    def pull(self, wait=False):
        """Newest completed snapshot as {module name: state_dict of CPU tensors}, or None when nothing has been published yet
        (`_pull_params`, deepQlearning.py:144-148, then `load_state_dict`)."""
        with self._lock:
            order = sorted(range(2), key=lambda j: -self._version[j])
            for j in order:
                ev = self._events[j]
                if ev is None:
                    continue
                if wait:
                    ev.synchronize()
                if ev.query():
                    out = {m: {} for m in self.modules}
                    for k, v in self._bufs[j].items():
                        m, name = k.split(".", 1)
                        out[m][name] = v.clone()
                    return out, self._version[j]
        return None
