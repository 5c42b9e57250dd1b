"""CPU oracle for the volume ray-march hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  See oracle/oracle.h.
"""
