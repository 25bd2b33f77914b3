"""CPU oracle for the ManifoldEM per-PD distance / diffusion-map front end.

TEST INFRASTRUCTURE ONLY.  Nothing under ``manifoldem_python_b200/`` imports
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the
checker / the timed CPU arm.  The product path has no CPU fallback.

Parity status: PINNED.  The restatement is checked against outputs of the
reference itself (imported unmodified from /root/reference in the build
container by ``tests/golden/make_golden.py``); the resulting vectors are
committed under ``tests/golden/*.npz`` and re-checked by
``tests/test_oracle_golden.py``.
"""
