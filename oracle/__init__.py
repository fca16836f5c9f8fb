"""TEST INFRASTRUCTURE: CPU oracle for the Veritas hot path (see veritas_oracle.c, ref_harness.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
