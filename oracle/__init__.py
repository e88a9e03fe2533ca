"""CPU oracle for the PTV scattered-to-grid hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ptv_interpolation_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker or the timed CPU arm.

Parity status: **pinned against the live reference** -- ``oracle/gen_golden.py`` imports
``/root/reference/interpolator.py`` and ``physics.py`` (with a stub ``tifffile``) in the
build container and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
this restatement against those vectors bit-for-bit (fp64).  The reference itself ships no
golden vectors or known-answer tests for this path.
"""
