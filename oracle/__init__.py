"""CPU oracle for the pyfe3d element-evaluation + assembly hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy restatement of the
reference's algorithm (``/root/reference/pyfe3d/*.pyx``) plus a build recipe
(``build_ref.py``) that compiles the unmodified reference into
``oracle/_ref``.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product package ``pyfe3d_b200`` never does and has no CPU fallback.

Parity status: PINNED.  ``tests/test_oracle_vs_golden.py`` checks every oracle
function against golden COO vectors produced by the compiled reference
(``tests/golden/make_golden.py`` is the generating script), and, when
``oracle/_ref`` is importable, against the reference run live.
"""
