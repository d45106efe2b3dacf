"""oracle -- float64 CPU restatement of the reference ShipEnv transition.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this package, and only as the checker (or the timed CPU baseline) -- never on the product path.
See shipsim_oracle.c for the pinning status ("reference logic pinned via golden fixtures; Chipmunk layer
parity unpinned").
"""
from .cbind import OracleEnv, build, lib, FLAG_COLLIDING, FLAG_GOAL, FLAG_OOB, FLAG_TIMEOUT, FLAG_ALLGOALS  # noqa: F401
