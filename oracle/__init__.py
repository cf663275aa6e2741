"""ORACLE -- TEST INFRASTRUCTURE ONLY (see list_oracle.py / ref_port.py headers).

Nothing under oracle/ is imported by the product package `list_b200`; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs use it, and only as the checker or the timed CPU baseline.
"""
