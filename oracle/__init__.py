"""CPU oracle for the GIRIH star-stencil path.  TEST INFRASTRUCTURE -- not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  girih_b200/ never does.
"""
