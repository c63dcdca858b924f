"""CPU oracle for the PT-MCMC hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu-baseline legs may import this package.
"""
