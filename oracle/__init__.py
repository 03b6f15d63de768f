"""ORACLE — test infrastructure, not product code.

CPU restatements of the reference's hot path used only as the checker:
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; nothing under emotiongestures_b200/ does.
"""
