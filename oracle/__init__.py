"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy) of the reference's tangent-image hot path plus the shim that imports the
live reference where it is mounted.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py may import anything from here; the product package never does.
"""
