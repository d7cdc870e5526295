"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the STEP hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; step_b200/ never does (tests/test_no_oracle_in_product.py
greps for it).
"""
