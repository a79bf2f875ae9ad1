"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/nbody_oracle.c and oracle/ecs_oracle.py).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  Nothing under recs_b200/ imports this package.
"""
from .nbody import NBodyOracle, load as load_nbody  # noqa: F401
