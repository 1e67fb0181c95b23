"""CPU oracle (TEST INFRASTRUCTURE): only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.  See sppm_oracle.h."""
from .oracle import Oracle, build_oracle  # noqa: F401
