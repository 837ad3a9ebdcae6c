"""CPU oracle of the GLimpSW meshlet raster path — TEST INFRASTRUCTURE ONLY.

Import this only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product (glimpsw_b200/) never touches it. See oracle.cpp.
"""
