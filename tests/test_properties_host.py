"""Property tests (hypothesis) of the host-side logic and of the oracle's restatements.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import cpu_restatement as O


class _Cursor:
    """ReplayMemory._advance without an arena: the closed form must equal n single steps of replay_memory.py:45-46 (quirk Q1)."""

    def __init__(self, cap, top, ln):
        self._maxlen, self._top, self._curr_len = cap, top, ln


@settings(max_examples=300, deadline=None)
@given(cap=st.integers(2, 60), top=st.integers(0, 59), grown=st.booleans(), n=st.integers(0, 200))
def test_cursor_closed_form_equals_single_steps(cap, top, grown, n):
    from fastdeepqlearning_b200.Replay.replay_memory import ReplayMemory
    top %= cap
    ln = cap - 1 if grown else top
    a, b = _Cursor(cap, top, ln), _Cursor(cap, top, ln)
    ReplayMemory._advance(a, n)
    for _ in range(n):  # the reference, one row at a time
        b._top = (b._top + 1) % cap
        b._curr_len = max(b._top, b._curr_len)
    assert (a._top, a._curr_len) == (b._top, b._curr_len)


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 40), world=st.integers(1, 9))
def test_shards_partition_the_actor_streams(n, world):
    from fastdeepqlearning_b200.parallel import shards_of_rank
    parts = [shards_of_rank(n, r, world) for r in range(world)]
    assert sorted(sum(parts, [])) == list(range(n))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


@settings(max_examples=100, deadline=None)
@given(data=st.data(), L=st.integers(1, 40), C=st.integers(1, 5), gamma=st.floats(0.5, 1.0))
def test_vmap_returns_sane_mode_is_the_return_up_to_the_next_terminal(data, L, C, gamma):
    r = np.array(data.draw(st.lists(st.floats(-2, 2, width=32), min_size=L * C, max_size=L * C)), np.float32).reshape(L, C)
    d = np.array(data.draw(st.lists(st.booleans(), min_size=L * C, max_size=L * C))).reshape(L, C)
    got = O.vmap_returns(r, d, gamma, reference_done_quirk=False)
    for c in range(C):
        for t in range(L):
            g, w = 0.0, 1.0
            for m in range(t, L):
                g += w * float(r[m, c])
                if d[m, c]:
                    break
                w *= gamma
            assert abs(got[t, c] - g) <= 1e-4 * max(1.0, abs(g))
    # the reference's arithmetic (quirk Q7) differs exactly where a row is not a virtual terminal and has a successor
    quirk = O.vmap_returns(r, d, gamma, reference_done_quirk=True)
    assert np.array_equal(quirk[-1], got[-1])


@settings(max_examples=100, deadline=None)
@given(data=st.data(), n=st.integers(2, 24), k=st.integers(1, 24))
def test_quantile_huber_is_invariant_to_the_order_of_the_samples(data, n, k):
    q = np.array(data.draw(st.lists(st.floats(-5, 5, width=32), min_size=n, max_size=n)), np.float64)[None]
    s = np.array(data.draw(st.lists(st.floats(-5, 5, width=32), min_size=k, max_size=k)), np.float64)[None]
    perm = np.array(data.draw(st.permutations(range(k))))
    a, b = O.quantile_huber(q, s), O.quantile_huber(q, s[:, perm])
    np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)
    ga, gb = O.quantile_huber_grad(q, s), O.quantile_huber_grad(q, s[:, perm])
    np.testing.assert_allclose(ga, gb, rtol=1e-12, atol=1e-12)
    # shifting predictions and samples together changes nothing
    np.testing.assert_allclose(O.quantile_huber(q + 3.25, s + 3.25), a, rtol=1e-9, atol=1e-9)


@settings(max_examples=100, deadline=None)
@given(p=st.floats(0.0, 0.99), cq=st.integers(1, 256))
def test_dropped_atoms_follow_the_reference_truncation(p, cq):
    from fastdeepqlearning_b200.ops import n_atoms_dropped
    assert n_atoms_dropped(p, cq) == int(p * cq) == O.n_atoms_dropped(p, cq)
