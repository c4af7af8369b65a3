"""CPU oracle for integrals.jl_b200 -- TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liboracle.so (oracle.c: a plain-C restatement of the Integrals.jl
hot path; every function there cites the reference file:line it follows).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (integrals.jl_b200) never does.
"""
from .oracle import *  # noqa: F401,F403
