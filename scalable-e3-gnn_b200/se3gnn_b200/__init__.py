"""se3gnn_b200 — B200-native (sm_100a) hot path of Scalable-E3-GNN.

Octree graph construction + SEGNN (l<=1) steerable message passing as hand-written CUDA
kernels behind a C ABI (``include/se3gnn_b200.h``).  The drop-in module for the reference's
``models.segnn.l1_tensor_prod.L1TensorProduct`` lives at the same import path once the
``scalable-e3-gnn_b200`` directory is on ``sys.path``.
"""
from .irreps import Instruction, Irrep, Irreps, MulIr, as_irreps  # noqa: F401

__version__ = "0.1.0"
