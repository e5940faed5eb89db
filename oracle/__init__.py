"""CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  Nothing under rxmesh_b200/ does.
"""
from .oracle import *  # noqa: F401,F403
