"""CPU oracle for the nerf-experiments hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: it may be imported
only by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.  The product package (``nerf_experiments_b200``) never imports it and has
no CPU fallback.

Every function restates, in plain PyTorch fp32 on the CPU (numpy where the arithmetic order has
to be pinned bit for bit), the algorithm of one reference function and cites the reference
``file:line`` it follows (paths relative to the root of sarphiv/nerf-experiments).

Pinning: the reference ships no golden vectors for this path, so the oracle is
pinned against outputs of the reference itself: ``tests/golden/make_golden.py`` imports the
unmodified reference modules from ``/root/reference`` (with the stub packages under
``oracle/_stubs``), runs them on seeded inputs and stores inputs + outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks every oracle function against those
fixtures, and the closed-form known-answer checks of the reference's notebook
(``barf/bug_hunting_with_Lauge.ipynb`` cells 8, 30, 40-42) are re-encoded in
``tests/test_oracle_golden.py`` (meshgrid / orthogonality / value-range checks) and ``tests/test_host_logic.py``.

Exception — parity unpinned: ``ref_nerfacc`` restates nerfacc's importance sampling, whose
source is not part of the reference tree and which is not installed here (see that module).
"""
