"""CPU oracle: a numpy restatement of the QuantumLiquids/PEPS VMC sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``peps_b200/`` imports this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker / the timed CPU arm.

Every function cites the reference file:line (relative to
``/root/reference/include/qlpeps/``) whose algorithm it restates.  The arithmetic the
reference delegates to TensorToolkit (``Contract``/``QR``/``SVD``; not vendored, see
SURVEY.md section 8c) is restated with numpy ``einsum``/LAPACK ``geqrf``/``gesdd``.

Parity pinning (see tests/test_oracle_kat.py and tests/golden/):
  * K1/K2: exact OBC Ising partition function by transfer matrix (closed form).
  * K4: 2x2 Heisenberg / TFIM fixtures -> analytic / reference-quoted energies.
  * K6: 4x4 D=8 Heisenberg fixture -> ED energy within MC error.
Bit-level agreement with TensorToolkit's SVD truncation tie-breaking is unpinned
(TensorToolkit is absent from this container).
"""
