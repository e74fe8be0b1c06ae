"""CPU oracle for the HighRes-net inference + scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The CUDA
path in ``highres-net_b200/`` never falls back to this code.

Parity status: PINNED.  The reference repo holds no tests or golden vectors of
its own, so the oracle is pinned against outputs of the
unmodified reference modules executed in the build container by
``oracle/make_golden.py`` (fixtures under ``tests/golden/``) and re-checked by
``tests/test_oracle_golden.py`` on every CPU test run.
"""
