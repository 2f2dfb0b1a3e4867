"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy / torch-CPU, fp32 storage with fp64 cross-checks) of the franQ
learner hot path that ``fastdeepqlearning_b200`` accelerates.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import anything from this package.  The product package never does: it fails
loudly when its CUDA library is missing.

Parity status: PINNED.  ``oracle/make_goldens.py`` executes the unmodified reference
modules from ``/root/reference`` (through ``oracle/ref_loader.py``) on seeded inputs with
injected index / goal streams and stores their outputs in ``tests/golden/*.npz``;
``tests/test_oracle_vs_golden.py`` checks every oracle function against those files and
against the reference's own known-answer test (``tests/test_replays.py:16-33``).
The vmap-HER variant (``her_vmap.py``, needs jax) cannot be executed here: the
restatement of that variant alone is "parity unpinned".
"""
