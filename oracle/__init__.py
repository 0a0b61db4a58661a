"""Test infrastructure: CPU oracle (plain-C restatement) and the reference-build harness.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker. The product (leela_b200) never does.
"""
