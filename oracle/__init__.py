"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/s2a_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
