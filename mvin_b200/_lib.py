"""ctypes binding of libmvin_b200.so (include/mvin_b200.h).  No CPU fallback: if the library is missing and cannot
be built, or a call fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

MVIN_OK = 0
FLAGS_ALL = 0x1F
ABI_VERSION = 4

EXPORTS = ["mvin_abi_version", "mvin_last_error", "mvin_create", "mvin_destroy", "mvin_bind_params",
           "mvin_bind_grads", "mvin_bind_adjacency", "mvin_bind_entity_shards", "mvin_set_batch_scale", "mvin_ipc_export", "mvin_ipc_open", "mvin_pack_adjacency", "mvin_workspace_bytes",
           "mvin_get_neighbors", "mvin_forward", "mvin_importance", "mvin_backward", "mvin_adam_step",
           "mvin_feed_bytes", "mvin_train_step_host", "mvin_launch_count", "mvin_profile_enable",
           "mvin_profile_read", "mvin_test_umma_gemm", "mvin_test_umma_dw", "mvin_bind_user_triplets", "mvin_gather_feed",
           "mvin_train_step_users_host", "mvin_ctr_metrics", "mvin_sample_adjacency", "mvin_build_ripple_sets", "mvin_feed_prefetch",
           "mvin_train_step_prefetched", "mvin_topk_metrics", "mvin_test_umma_bf16", "mvin_xchg_bind", "mvin_xchg_expand",
           "mvin_xchg_owner_forward", "mvin_xchg_owner_backward", "mvin_xchg_finish_backward"]


class Config(C.Structure):
    _fields_ = [("dim", C.c_int32), ("neighbor_sample_size", C.c_int32), ("h_hop", C.c_int32),
                ("n_mix_hop", C.c_int32), ("p_hop", C.c_int32), ("n_memory", C.c_int32),
                ("n_user", C.c_int32), ("n_entity", C.c_int32), ("n_relation", C.c_int32),
                ("max_batch", C.c_int32), ("l2_weight", C.c_float), ("l2_agg_weight", C.c_float),
                ("flags", C.c_int32)]


PARAM_FIELDS = ["user_emb", "entity_emb", "relation_emb", "relation_kge", "mix_w", "mix_b", "user_mlp_w",
                "user_mlp_b", "transfer_w", "transfer_b", "h_item_w", "h_item_b", "agg_w", "agg_b", "agg_urh_w",
                "agg_urh_b"]


class Params(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in PARAM_FIELDS]


class MvinError(RuntimeError):
    pass


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path):
        path = _build.build()                       # raises if nvcc is unavailable
    lib = C.CDLL(path)
    vp, i32, i64p = C.c_void_p, C.c_int32, C.c_void_p
    lib.mvin_abi_version.restype = C.c_int
    lib.mvin_last_error.restype = C.c_char_p
    lib.mvin_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.mvin_destroy.argtypes = [vp]
    lib.mvin_bind_params.argtypes = [vp, C.POINTER(Params)]
    lib.mvin_bind_grads.argtypes = [vp, C.POINTER(Params)]
    lib.mvin_bind_adjacency.argtypes = [vp, vp]
    lib.mvin_bind_entity_shards.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp)]
    lib.mvin_set_batch_scale.argtypes = [vp, i32, C.c_float]
    lib.mvin_ipc_export.argtypes = [vp, vp, C.POINTER(C.c_int64)]
    lib.mvin_ipc_open.argtypes = [vp, C.c_int64, C.POINTER(vp)]
    lib.mvin_pack_adjacency.argtypes = [vp, vp, i32, i32, vp, vp]
    lib.mvin_workspace_bytes.argtypes = [vp, i32]
    lib.mvin_workspace_bytes.restype = C.c_size_t
    lib.mvin_get_neighbors.argtypes = [vp, i64p, i32, i32, C.POINTER(vp), C.POINTER(vp), vp]
    lib.mvin_forward.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp]
    lib.mvin_importance.argtypes = [vp, vp, vp, vp, vp]
    lib.mvin_backward.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.mvin_adam_step.argtypes = [vp, C.POINTER(Params), C.POINTER(Params), C.c_float, C.c_float, C.c_float,
                                   C.c_float, i32, vp]
    lib.mvin_feed_bytes.argtypes = [vp, i32]
    lib.mvin_feed_bytes.restype = C.c_size_t
    lib.mvin_train_step_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, C.POINTER(Params),
                                         C.POINTER(Params), C.c_float, i32, vp, vp]
    lib.mvin_test_umma_gemm.argtypes = [vp, vp, vp, C.c_int64, i32, vp]
    lib.mvin_test_umma_dw.argtypes = [vp, vp, vp, C.c_int64, i32, i32, vp]
    lib.mvin_test_umma_bf16.argtypes = [vp, vp, vp, vp, vp, C.c_int64, i32, i32, vp]
    lib.mvin_bind_user_triplets.argtypes = [vp, vp]
    lib.mvin_gather_feed.argtypes = [vp, vp, i32, vp, vp, vp, vp]
    lib.mvin_train_step_users_host.argtypes = [vp, vp, vp, vp, i32, vp, vp, C.POINTER(Params), C.POINTER(Params),
                                               C.c_float, i32, vp, vp]
    lib.mvin_ctr_metrics.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    lib.mvin_sample_adjacency.argtypes = [vp, vp, vp, i32, i32, C.c_uint64, vp, vp, vp, vp, vp]
    lib.mvin_build_ripple_sets.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, C.c_uint64, vp, vp, vp]
    lib.mvin_feed_prefetch.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, vp, i32, vp]
    lib.mvin_train_step_prefetched.argtypes = [vp, i32, vp, i32, vp, C.POINTER(Params), C.POINTER(Params), C.c_float, i32,
                                               vp, vp]
    lib.mvin_topk_metrics.argtypes = [vp, vp, vp, vp, i32, i32, vp, i32, vp, vp, vp, vp]
    lib.mvin_xchg_bind.argtypes = [vp, i32, i32, C.c_int64, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.mvin_xchg_expand.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.mvin_xchg_owner_forward.argtypes = [vp, i32, vp, vp]
    lib.mvin_xchg_owner_backward.argtypes = [vp, i32, vp, vp]
    lib.mvin_xchg_finish_backward.argtypes = [vp, vp, vp]
    lib.mvin_launch_count.argtypes = [vp]
    lib.mvin_launch_count.restype = C.c_int64
    lib.mvin_profile_enable.argtypes = [vp, i32]
    lib.mvin_profile_read.argtypes = [vp, C.c_char_p, C.c_size_t]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is None or (name not in ("mvin_last_error", "mvin_workspace_bytes", "mvin_feed_bytes",
                                               "mvin_launch_count")):
            fn.restype = C.c_int
    lib.mvin_last_error.restype = C.c_char_p
    if lib.mvin_abi_version() != ABI_VERSION:
        raise MvinError(f"libmvin_b200.so ABI {lib.mvin_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != MVIN_OK:
        msg = load().mvin_last_error()
        raise MvinError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
