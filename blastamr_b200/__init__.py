"""blastamr_b200 -- B200-native AMR decision path (estimate, mark, 2:1 balance, map).

The product is the CUDA C-ABI library `libamr_b200.so` (see include/amr_b200.h); the Python
modules here are the thin host-side binding used by the tests and the benchmark.
"""
__version__ = "0.1.0"
