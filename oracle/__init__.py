"""ORACLE — test infrastructure only.

CPU restatements of the reference algorithms on the hot path, used as the parity checker.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package. The product (`hpc-ops_b200/`) never does: it fails loudly when its
CUDA library is missing instead of falling back to anything here.

Pinning status:
  * taskmap   — pinned against the real reference CPU scheduler compiled in place from
                /root/reference (oracle/_ref/libref_taskmap.so) and tests/golden/taskmap_*.npz.
  * attention — pinned against golden vectors produced by executing the reference's own test
                reference functions (AST-extracted from /root/reference/tests/*.py) on CPU;
                see tests/golden/make_golden.py.
"""
