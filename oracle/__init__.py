"""TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's FP8 blockwise-GEMM path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import
this package, and only as the checker / the timed CPU baseline -- the product path (``deepgemm_b200``) never does.
"""
