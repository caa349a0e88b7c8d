"""ctypes binding of ``libembeddingnet_b200.so`` (the C ABI declared in ``include/embeddingnet_b200.h``).

There is no CPU fallback: if the shared library is missing (or was built for another architecture) every entry
point raises.  Build it with ``python -m embeddingnet_b200.build``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# EMBEDDINGNET_B200_LIB: developer override (A/B timing of a variant build, tools/ab_bh.py); unset in production
LIB_PATH = os.environ.get("EMBEDDINGNET_B200_LIB") or os.path.join(HERE, "libembeddingnet_b200.so")

EN_MODE_SEMIHARD, EN_MODE_HARDEST, EN_MODE_RANDOM_HARD = 0, 1, 2
EN_KNN_SLACK, EN_KNN_MAX_K, EN_KNN_STREAM_MAX_Q, EN_KNN_EXACT_MAX_Q, EN_KNN_SMALLQ_MAX_Q = 3, 29, 8, 64, 64
EN_PREC_TF32X3, EN_PREC_BF16X3 = 0, 1
EN_MINE_MAX_SLOTS = 8
EN_COMM_ID_BYTES = 128

P = c_void_p  # every device pointer / stream crosses the boundary as an opaque address

# name -> (restype, argtypes); mirrors include/embeddingnet_b200.h one to one
SIGNATURES = {
    "en_version": (c_char_p, []),
    "en_last_error": (c_char_p, []),
    "en_launch_count": (c_int64, []),
    "en_launch_count_reset": (None, []),
    "en_prof_enable": (c_int, [c_int]),
    "en_prof_last_ms": (c_int, [P]),
    "en_prof_marks_ms": (c_int, [P, c_int, P]),
    "en_l2_normalize_fwd": (c_int, [P, P, c_int64, c_int, P]),
    "en_l2_normalize_bwd": (c_int, [P, P, P, c_int64, c_int, P]),
    "en_triplet_apn_fwd": (c_int, [P, c_int64, c_int, c_float, P, P]),
    "en_triplet_apn_bwd": (c_int, [P, P, c_int64, c_int, c_float, P, P]),
    "en_contrastive_fwd": (c_int, [P, P, c_int64, P, P]),
    "en_contrastive_bwd": (c_int, [P, P, P, c_int64, P, P]),
    "en_pair_accuracy": (c_int, [P, P, c_int64, P, P]),
    "en_siamese_l2_fwd": (c_int, [P, P, c_int64, c_int, P, P]),
    "en_siamese_l2_bwd": (c_int, [P, P, P, c_int64, c_int, P, P, P]),
    "en_siamese_l1_fwd": (c_int, [P, P, c_int64, P, P]),
    "en_siamese_l1_bwd": (c_int, [P, P, P, c_int64, P, P, P]),
    "en_query_distances": (c_int, [P, P, c_int64, c_int, P, P]),
    "en_scale_inplace": (c_int, [P, c_int64, P, P]),
    "en_ws_bytes_pairwise": (c_size_t, [c_int64, c_int, c_int]),
    "en_pairwise_dist": (c_int, [P, c_int64, c_int, c_int, c_int, P, P, c_size_t, P]),
    "en_mine_batch_scan": (c_int, [P, P, c_int64, P, c_int64, c_float, P, P, P, P]),
    "en_mine_batch_select": (c_int, [P, P, c_int64, P, c_int64, c_float, c_int, P, P, P]),
    "en_loss_scan": (c_int, [P, c_int64, c_float, P, P]),
    "en_loss_select": (c_int, [P, c_int64, c_float, c_int, c_int, P, P]),
    "en_gather_triplet_rows": (c_int, [P, c_int64, c_int64, P, c_int64, P, P, P, P]),
    "en_ws_bytes_batch_hard": (c_size_t, [c_int64, c_int]),
    "en_batch_hard_fwd": (c_int, [P, P, c_int64, c_int, c_float, c_int, c_int, P, P, P, P, P, P, P, c_size_t, P]),
    "en_batch_hard_fwd_bwd": (c_int, [P, P, c_int64, c_int, c_float, c_int, c_int, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    "en_batch_hard_bwd": (c_int, [P, c_int64, c_int, c_int, P, P, P, P, P, P, P, P]),
    "en_bh_host_pipe_device_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "en_bh_host_pipe_create": (c_int, [c_int64, c_int, c_float, c_int, c_int, c_int, P, c_size_t, ctypes.POINTER(c_void_p)]),
    "en_bh_host_pipe_submit": (c_int, [P, P, P, P, P, P, P, ctypes.POINTER(c_int64)]),
    "en_bh_host_pipe_wait": (c_int, [P, c_int64]),
    "en_bh_host_pipe_destroy": (c_int, [P]),
    "en_ws_bytes_batch_all": (c_size_t, [c_int64, c_int, c_int]),
    "en_batch_all_fwd": (c_int, [P, P, c_int64, c_int, c_float, c_int, c_int, P, P, P, c_size_t, P]),
    "en_batch_all_bwd": (c_int, [P, P, c_int64, c_int, c_float, c_int, c_int, P, P, P, P, c_size_t, P]),
    "en_batch_all_fwd_bwd": (c_int, [P, P, c_int64, c_int, c_float, c_int, c_int, P, P, P, P, P, P, c_size_t, P]),
    "en_ws_bytes_contrastive_allpairs": (c_size_t, [c_int64, c_int]),
    "en_contrastive_allpairs_fwd": (c_int, [P, P, c_int64, c_int, P, P, c_size_t, P]),
    "en_contrastive_allpairs_bwd": (c_int, [P, P, c_int64, c_int, P, P, P, c_size_t, P]),
    "en_contrastive_allpairs_fwd_bwd": (c_int, [P, P, c_int64, c_int, P, P, P, P, c_size_t, P]),
    "en_bank_dpad": (c_int, [c_int, c_int]),
    "en_bank_plane_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "en_bank_prepare": (c_int, [P, c_int64, c_int, c_int, P, P, P, P]),
    "en_ws_bytes_knn": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "en_knn_shard_topk": (c_int, [P, c_int64, c_int, P, P, P, P, c_int64, c_int64, c_int, c_int, P, P, P, P, P, P,
                                  c_size_t, P]),
    "en_ws_bytes_knn_smallq": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "en_knn_smallq_topk": (c_int, [P, c_int64, c_int, P, P, P, P, c_int64, c_int64, c_int, P, P, P, P, c_size_t, P]),
    "en_ws_bytes_knn_exact": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "en_knn_exact_topk": (c_int, [P, c_int64, c_int, P, c_int64, c_int64, c_int, P, P, P, P, P, c_size_t, P]),
    "en_knn_exact_redo": (c_int, [P, c_int64, c_int, P, c_int64, c_int64, c_int, P, P, P, P, P, P, c_size_t, P]),
    "en_ws_bytes_knn_stream": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "en_knn_stream_topk": (c_int, [P, c_int64, c_int, P, P, c_int64, c_int64, c_int, P, P, P, P, c_size_t, P]),
    "en_knn_merge": (c_int, [P, P, c_int, c_int64, c_int, P, P, P]),
    "en_knn_merge_packed": (c_int, [P, c_int, c_int64, c_int, P, P, P]),
    "en_knn_finalize_dist": (c_int, [P, c_int64, P, P]),
    "en_knn_vote": (c_int, [P, c_int64, c_int, P, c_int64, P, P]),
    "en_knn_accuracy": (c_int, [P, P, P, c_int64, c_int, P, c_int64, P, P]),
    "en_pair_dist_exact": (c_int, [P, P, c_int64, c_int, P, P]),
    "en_ws_bytes_mine_bank": (c_size_t, [c_int64, c_int]),
    "en_mine_bank_count": (c_int, [P, P, P, c_int64, c_int, c_int, c_float, P, P, P, P, P, c_int64, c_int, P, P, c_size_t, P]),
    "en_mine_bank_select": (c_int, [P, P, P, c_int64, c_int, c_int, c_float, c_int, P, P, P, P, P, P, c_int64, c_int64, c_int,
                                    P, P, c_size_t, P]),
    "en_dense_plane_bytes": (c_size_t, [c_int, c_int]),
    "en_dense_prepare": (c_int, [P, c_int, c_int, P, P, P]),
    "en_ws_bytes_dense": (c_size_t, [c_int64, c_int]),
    "en_dense_relu_fwd": (c_int, [P, c_int64, c_int, P, P, P, c_int, c_int, P, P, P, c_size_t, P]),
    "en_ws_bytes_dense_bwd": (c_size_t, [c_int64, c_int, c_int]),
    "en_dense_relu_bwd": (c_int, [P, c_int64, c_int, P, c_int, c_int, P, P, P, P, P, P, P, c_size_t, P]),
    "en_comm_unique_id": (c_int, [P]),
    "en_comm_init": (c_int, [c_int, c_int, P, ctypes.POINTER(c_void_p)]),
    "en_comm_allgather": (c_int, [P, P, P, c_size_t, P]),
    "en_comm_allreduce_max_i64": (c_int, [P, P, P, c_size_t, P]),
    "en_comm_destroy": (c_int, [P]),
    "en_synth_fill": (c_int, [P, c_int64, c_int, c_int64, c_uint64, c_uint64, c_int64, c_int64, c_float, c_int, P, P]),
}

_lib = None


class EmbeddingNetB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EmbeddingNetB200Error(
            "embeddingnet_b200: %s is missing. Build it with `python -m embeddingnet_b200.build` "
            "(nvcc, sm_100a). There is no CPU / PyTorch fallback." % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    """Translate a C-ABI status into the reference's error convention (Python exceptions)."""
    if rc == 0:
        return
    msg = load().en_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError("%s: %s" % (what or "embeddingnet_b200", msg))
    raise EmbeddingNetB200Error("%s: CUDA error %d: %s" % (what or "embeddingnet_b200", rc, msg))


def call(name: str, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)
