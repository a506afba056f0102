"""uf3_b200 — B200-native UF3 featurization / evaluator hot path.

Python host layer (reference API mirror) over hand-written sm_100a CUDA reached
through a C ABI (`include/uf3b.h`, `uf3_b200/csrc/`).  See DESIGN.md.
"""
__version__ = "0.1.0"
