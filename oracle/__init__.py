"""TEST INFRASTRUCTURE -- CPU oracle for the LTE turbo-decoding hot path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this package, and only as the checker / reported baseline.  The product package
(openair4g_b200/) never imports it.
"""
