"""CPU oracle for the hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
anything under oracle/.  The product package (music-generation-emotion-adaptive_b200/) never does:
it has no CPU fallback and fails loudly when the CUDA library is missing.
"""
