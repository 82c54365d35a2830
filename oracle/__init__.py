"""CPU ORACLE -- test infrastructure, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
``f1tenth_planning_b200`` never does (tests/test_no_oracle_in_product.py enforces it).

Parity status: nearest_point / intersect_point / get_actuation /
PurePursuitPlanner are pinned against the imported reference (tests/golden/); the cubic-spiral
generator, raceline deviation and collision stages have no reference code and are
"parity unpinned" -- this package is their definition.
"""
