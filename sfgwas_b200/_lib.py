"""ctypes loader for libsfgwas_b200.so (the C ABI of include/sfgwas_b200.h).

The library is the product: there is no Python / CPU fallback.  Loading fails loudly when the shared object is
missing, and every compute entry point fails when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SFG_B200_LIB: load another build of the library (A/B runs of kernel variants under profiles/); the default is the in-tree build
LIB_PATH = os.environ.get("SFG_B200_LIB") or os.path.join(_HERE, "lib", "libsfgwas_b200.so")

# every symbol declared in include/sfgwas_b200.h (checked by tests/test_abi.py)
SYMBOLS = [
    "sfg_ctx_create", "sfg_ctx_destroy", "sfg_last_error", "sfg_version", "sfg_ctx_set_cache_budget",
    "sfg_ctx_launch_count", "sfg_ctx_encoder_stats", "sfg_ctx_psi", "sfg_ctx_set_rotation_key",
    "sfg_ctx_set_rotation_key_ptrs", "sfg_ctx_has_rotation_key", "sfg_ntt", "sfg_mul_coeffs_and_add128",
    "sfg_reduce_and_add_uint128", "sfg_mform_lvl", "sfg_rotate_right", "sfg_geno_create", "sfg_geno_push_rows",
    "sfg_geno_destroy", "sfg_encode_diag", "sfg_matmult4_stream_preprocess", "sfg_cache_destroy", "sfg_cache_info",
    "sfg_cache_get_diag", "sfg_matmult4_stream_compute", "sfg_matmult4_stream_compute_ptrs", "sfg_matmult4_stream",
    "sfg_cv_elems", "sfg_matmult4_partial", "sfg_cv_mod_reduce", "sfg_matmult4_finish", "sfg_ct_add", "sfg_ctx_sync",
    "sfg_matmult4_stream_compute_dev", "sfg_ctx_last_timings", "sfg_ctx_stream", "sfg_ctx_set_relin_key",
    "sfg_ctx_set_relin_key_ptrs", "sfg_ct_mul_relin", "sfg_ct_mul_plain", "sfg_ct_rescale", "sfg_ct_sub", "sfg_ct_add2",
    "sfg_inner_sum_all", "sfg_encode_slots_i8", "sfg_cache_write_files", "sfg_cache_load_files",
    "sfg_geno_count_sketch", "sfg_ntt_dev", "sfg_rotate_right_dev",
    "sfg_matmult4_stream_preprocess_rows", "sfg_matmult4_stream_preprocess_giants", "sfg_ct_mod_reduce",
    "sfg_matmult4_finish_dev", "sfg_cipher_matrix_save", "sfg_cipher_matrix_info", "sfg_cipher_matrix_load",
    "sfg_refresh_gen_shares", "sfg_refresh_finish", "sfg_matmult4_baby_chunk_bytes", "sfg_matmult4_baby_dev",
    "sfg_matmult4_stream_compute_r_dev", "sfg_cts_upload", "sfg_cts_download", "sfg_cts_shape", "sfg_cts_destroy", "sfg_cts_slice",
    "sfg_cts_matmult4_stream_compute", "sfg_cts_mul_relin", "sfg_cts_mul_plain", "sfg_cts_addsub", "sfg_cts_inner_sum_all",
]

_lib = None


class SfgError(RuntimeError):
    pass


def load():
    """Load the shared library (no compute). Raises if it has not been built: python -m sfgwas_b200.build"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SfgError(
            f"{LIB_PATH} not found: build the CUDA library first (python -m sfgwas_b200.build or __graft_entry__.build()); "
            "there is no CPU fallback"
        )
    L = C.CDLL(LIB_PATH)
    vp, u64p, i32, sz = C.c_void_p, C.POINTER(C.c_uint64), C.c_int, C.c_size_t
    L.sfg_version.restype = i32
    L.sfg_ctx_create.restype = i32
    L.sfg_ctx_create.argtypes = [i32, i32, u64p, i32, u64p, i32, C.c_double, u64p, C.POINTER(vp)]
    L.sfg_ctx_destroy.argtypes = [vp]
    L.sfg_last_error.restype = C.c_char_p
    L.sfg_last_error.argtypes = [vp]
    L.sfg_ctx_set_cache_budget.argtypes = [vp, sz]
    L.sfg_ctx_launch_count.restype = C.c_ulonglong
    L.sfg_ctx_launch_count.argtypes = [vp]
    L.sfg_ctx_encoder_stats.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    L.sfg_ctx_psi.argtypes = [vp, u64p]
    L.sfg_ctx_set_rotation_key.argtypes = [vp, i32, vp]
    L.sfg_ctx_set_rotation_key_ptrs.argtypes = [vp, i32, vp]
    L.sfg_ctx_has_rotation_key.argtypes = [vp, i32]
    L.sfg_ntt.argtypes = [vp, vp, i32, C.POINTER(i32), i32, i32]
    L.sfg_mul_coeffs_and_add128.argtypes = [vp, vp, vp, vp, sz]
    L.sfg_reduce_and_add_uint128.argtypes = [vp, vp, vp, i32, sz]
    L.sfg_mform_lvl.argtypes = [vp, i32, vp]
    L.sfg_rotate_right.argtypes = [vp, i32, vp, i32, i32, vp]
    L.sfg_geno_create.argtypes = [vp, sz, sz, C.POINTER(vp)]
    L.sfg_geno_push_rows.argtypes = [vp, vp, sz]
    L.sfg_geno_destroy.argtypes = [vp]
    L.sfg_encode_diag.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    L.sfg_matmult4_stream_preprocess.argtypes = [vp, vp, i32, C.POINTER(vp)]
    L.sfg_cache_destroy.argtypes = [vp]
    L.sfg_cache_info.argtypes = [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.sfg_cache_get_diag.argtypes = [vp, vp, i32, i32, i32, vp, C.POINTER(i32)]
    L.sfg_matmult4_stream_compute.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.sfg_matmult4_stream_compute_ptrs.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.sfg_matmult4_stream_compute_dev.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.sfg_matmult4_stream.argtypes = [vp, vp, i32, i32, vp, i32, i32, i32, vp, vp, vp]
    L.sfg_cv_elems.restype = sz
    L.sfg_cv_elems.argtypes = [vp, vp, i32, i32]
    L.sfg_matmult4_partial.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, i32, vp]
    L.sfg_cv_mod_reduce.argtypes = [vp, vp, i32, i32, vp, sz, sz]
    L.sfg_matmult4_finish.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp]
    L.sfg_ct_add.argtypes = [vp, vp, vp, i32, i32, vp]
    L.sfg_ctx_set_relin_key.argtypes = [vp, vp]
    L.sfg_ctx_set_relin_key_ptrs.argtypes = [vp, vp]
    L.sfg_ct_mul_relin.argtypes = [vp, i32, vp, i32, i32, vp, i32, i32, i32, vp]
    L.sfg_ct_mul_plain.argtypes = [vp, i32, vp, i32, i32, vp, i32, i32, i32, vp]
    L.sfg_ct_rescale.argtypes = [vp, i32, vp, i32, i32, vp]
    L.sfg_ct_sub.argtypes = [vp, i32, vp, i32, i32, vp, i32, i32, vp]
    L.sfg_ct_add2.argtypes = [vp, i32, vp, i32, i32, vp, i32, i32, vp]
    L.sfg_inner_sum_all.argtypes = [vp, i32, vp, i32, i32, vp]
    L.sfg_encode_slots_i8.argtypes = [vp, vp, i32, i32, vp]
    L.sfg_cache_write_files.argtypes = [vp, vp, C.c_char_p]
    L.sfg_cache_load_files.argtypes = [vp, C.c_char_p, sz, sz, i32, C.POINTER(vp)]
    L.sfg_geno_count_sketch.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, C.POINTER(C.c_float)]
    L.sfg_ntt_dev.argtypes = [vp, vp, i32, C.POINTER(i32), i32, i32]
    L.sfg_rotate_right_dev.argtypes = [vp, i32, vp, i32, i32, vp]
    L.sfg_matmult4_stream_preprocess_rows.argtypes = [vp, vp, i32, i32, i32, C.POINTER(vp)]
    L.sfg_matmult4_stream_preprocess_giants.argtypes = [vp, vp, i32, i32, i32, C.POINTER(vp)]
    L.sfg_ct_mod_reduce.argtypes = [vp, vp, sz, i32]
    L.sfg_matmult4_finish_dev.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp]
    L.sfg_cipher_matrix_save.argtypes = [C.c_char_p, i32, vp, vp, i32, i32, i32]
    L.sfg_cipher_matrix_info.argtypes = [C.c_char_p, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.sfg_cipher_matrix_load.argtypes = [C.c_char_p, i32, vp, vp, i32, i32, i32]
    L.sfg_refresh_gen_shares.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, i32, C.c_double, C.c_double, vp, vp, vp, vp]
    L.sfg_refresh_finish.argtypes = [vp, i32, i32, vp, i32, C.c_double, C.c_double, vp, vp, vp, vp]
    L.sfg_matmult4_baby_chunk_bytes.restype = sz
    L.sfg_matmult4_baby_chunk_bytes.argtypes = [vp, vp, i32, i32]
    L.sfg_matmult4_baby_dev.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, i32, vp]
    L.sfg_matmult4_stream_compute_r_dev.argtypes = [vp, vp, i32, i32, vp, vp]
    L.sfg_cts_upload.argtypes = [vp, vp, i32, i32, C.POINTER(vp)]
    L.sfg_cts_download.argtypes = [vp, vp, vp]
    L.sfg_cts_shape.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.sfg_cts_destroy.argtypes = [vp]
    L.sfg_cts_slice.argtypes = [vp, vp, i32, i32, C.POINTER(vp)]
    L.sfg_cts_matmult4_stream_compute.argtypes = [vp, vp, i32, i32, i32, vp, C.POINTER(vp)]
    L.sfg_cts_mul_relin.argtypes = [vp, i32, vp, vp, i32, C.POINTER(vp)]
    L.sfg_cts_mul_plain.argtypes = [vp, i32, vp, i32, i32, vp, i32, C.POINTER(vp)]
    L.sfg_cts_addsub.argtypes = [vp, i32, vp, vp, i32, C.POINTER(vp)]
    L.sfg_cts_inner_sum_all.argtypes = [vp, i32, vp, i32, i32, C.POINTER(vp)]
    L.sfg_ctx_sync.argtypes = [vp]
    L.sfg_ctx_last_timings.argtypes = [vp, C.POINTER(C.c_float)]
    L.sfg_ctx_stream.restype = vp
    L.sfg_ctx_stream.argtypes = [vp]
    _lib = L
    return L
