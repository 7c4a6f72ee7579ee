"""CPU oracle for the descriptor-matching path -- TEST INFRASTRUCTURE ONLY.

See oracle/pgm_oracle.h.  Import rules: tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs only.
"""
