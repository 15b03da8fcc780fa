"""CPU oracle for the ZUTIS mask-decode + scoring path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; zutis_b200/ never does.
"""
