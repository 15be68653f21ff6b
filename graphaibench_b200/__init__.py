"""graphaibench_b200 — B200-native (sm_100a) GNN-layer hot path behind GraphAIBench's layer/model API.

Layout: csrc/ (hand-written CUDA kernels + the C ABI of include/gai_b200.h), host/ (C++ mirror of the reference's
LearningGraph / aggregator / layer / Model / Reader classes over that ABI), ops.py / model.py (ctypes + torch-tensor
front end used by tests and bench.py). There is no CPU fallback anywhere in this package."""
from ._abi import GaiError, lib  # noqa: F401
