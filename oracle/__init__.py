"""CPU oracle for the MM-PGO / AMM-PGO* / AMM-PGO# hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy/scipy restatement of the
reference algorithm (MurpheyLab/DPGO, C++/DPGO + C++/Optimization); each
function cites the reference file:line it follows.  It may be imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` -- never by the product path in
``dpgo_b200/``.

PARITY UNPINNED against the upstream binary: the reference cannot be built in
this image (Eigen3, SuiteSparse/CHOLMOD, glog and Boost are absent, see
DESIGN.md) and ships no golden traces for this path.  The oracle is pinned
instead by (i) the analytic STPCG / TNT known-answer tests the reference holds
in C++/Optimization/tests, (ii) algebraic invariants of the data matrices
(tests/test_oracle_*.py) and (iii) an independent direct evaluation of the
global objective.
"""
