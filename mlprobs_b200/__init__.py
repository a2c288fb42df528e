"""mlprobs_b200: B200-native all-pairs posterior + consistency engine behind the reference's stage interfaces.

Product code lives in csrc/ (CUDA kernels + C ABI, built into libmlprobs_b200.so); this package is the thin
host-side mirror of the reference's operator interface for that path.  Importing never touches oracle/.
"""
from ._capi import (K_NAMES, Engine, PinnedCsrBuffers, PinnedPackedBuffers, qp_guide_tree, qp_guide_tree_ex, qp_finish_alignment_host, cpnp_guide_tree, cpnp_finish_alignment_host, cpnp_np_finish_alignment_host, debug_glibc_rand, debug_glibc_rand_seeded, column_scores, shard_pairs, shard_pairs_within, nccl_unique_id, cpnp_model_adjustment, cpnp_g_features, MlpError, default_tables, load, LIB_PATH, QP, CPNP_P0, CPNP_P1, M_HMM5, M_PART, M_LOCAL)

__all__ = ["K_NAMES", "Engine", "PinnedCsrBuffers", "PinnedPackedBuffers", "qp_guide_tree", "qp_guide_tree_ex", "qp_finish_alignment_host", "cpnp_guide_tree", "cpnp_finish_alignment_host", "cpnp_np_finish_alignment_host", "debug_glibc_rand", "debug_glibc_rand_seeded", "column_scores", "shard_pairs", "shard_pairs_within", "nccl_unique_id", "cpnp_model_adjustment", "cpnp_g_features", "MlpError", "default_tables", "load", "LIB_PATH", "QP", "CPNP_P0", "CPNP_P1", "M_HMM5", "M_PART", "M_LOCAL"]
