"""TEST INFRASTRUCTURE ONLY — CPU oracle for the PartManip PPO hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (partmanip_b200) never does.
"""
