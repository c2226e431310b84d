"""TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference (uoo723/PMGT) algorithms on the PMGT
pre-training hot path, used as the checker for the CUDA path.  Nothing under
``pmgt_b200/`` may import from this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm do.
"""
