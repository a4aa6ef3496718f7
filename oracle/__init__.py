"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/tron_oracle.c header).

May be imported only by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference).
"""
