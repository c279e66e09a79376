"""ctypes binding of include/rj_b200.h -- the C-ABI of the CUDA engine.

This is the Python twin of the cgo/JNI-style stub shown in INTEGRATION.md: struct layouts and
prototypes only, no logic.  `load_library()` fails loudly when librj_b200.so has not been built;
there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

PAGE_SIZE = 8192  # reference include/plan.h:54

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librj_b200.so")


class rj_column_t(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("reserved", C.c_uint32),
        ("n_pages", C.c_uint64),
        ("pages", C.POINTER(C.c_void_p)),
        ("contiguous", C.c_void_p),
    ]


class rj_table_t(C.Structure):
    _fields_ = [
        ("num_rows", C.c_uint64),
        ("n_columns", C.c_uint32),
        ("reserved", C.c_uint32),
        ("columns", C.POINTER(rj_column_t)),
    ]


class rj_dense_column_t(C.Structure):
    _fields_ = [("type", C.c_int32), ("reserved", C.c_uint32), ("d_values", C.c_void_p), ("d_valid", C.c_void_p)]


class rj_dense_table_t(C.Structure):
    _fields_ = [("num_rows", C.c_uint64), ("n_columns", C.c_uint32), ("reserved", C.c_uint32),
                ("columns", C.POINTER(rj_dense_column_t))]


class rj_scatter_multi_t(C.Structure):
    _fields_ = [("keys_out", C.c_void_p * 8), ("rows_out", C.c_void_p * 8), ("n_payload", C.c_uint32),
                ("reserved", C.c_uint32), ("pay_src", C.c_void_p * 6), ("pay_width", C.c_int32 * 6),
                ("pay_dst", (C.c_void_p * 8) * 6)]


class rj_attr_t(C.Structure):
    _fields_ = [("index", C.c_uint64), ("type", C.c_int32), ("reserved", C.c_uint32)]


class rj_node_t(C.Structure):
    _fields_ = [
        ("is_join", C.c_int32),
        ("build_left", C.c_int32),
        ("base_table_id", C.c_uint64),
        ("left", C.c_uint64),
        ("right", C.c_uint64),
        ("left_attr", C.c_uint64),
        ("right_attr", C.c_uint64),
        ("n_output_attrs", C.c_uint32),
        ("reserved", C.c_uint32),
        ("output_attrs", C.POINTER(rj_attr_t)),
    ]


class rj_plan_t(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_uint32),
        ("n_inputs", C.c_uint32),
        ("nodes", C.POINTER(rj_node_t)),
        ("inputs", C.POINTER(rj_table_t)),
        ("root", C.c_uint64),
    ]


class rj_carry_scatter_t(C.Structure):
    _fields_ = [("d_keys", C.c_void_p), ("d_valid", C.c_void_p), ("n", C.c_uint64),
                ("d_region_start", C.c_void_p), ("d_tile_start", C.c_void_p), ("n_regions", C.c_uint32),
                ("shift", C.c_int32), ("bits", C.c_int32), ("d_cursor", C.c_void_p), ("d_keys_out", C.c_void_p),
                ("n_val", C.c_uint32), ("val_src", C.c_void_p * 2), ("val_dst", C.c_void_p * 2), ("val_width", C.c_int32 * 2),
                ("n_flag", C.c_uint32), ("flag_src", C.c_void_p * 2), ("flag_dst", C.c_void_p * 2),
                ("n_owners", C.c_uint32), ("owner_shift", C.c_int32), ("keys_dst_multi", C.c_void_p * 8),
                ("val_dst_multi", (C.c_void_p * 8) * 2), ("flag_dst_multi", (C.c_void_p * 8) * 2),
                ("d_src_table", C.c_void_p), ("d_region_group", C.c_void_p)]


class rj_part_side_t(C.Structure):
    _fields_ = [("d_keys", C.c_void_p), ("n", C.c_uint64), ("n_cols", C.c_uint32), ("d_vals", C.c_void_p * 2),
                ("types", C.c_int32 * 2), ("d_valid_bytes", C.c_void_p * 2),
                ("n_sub", C.c_uint32), ("d_src_table", C.c_void_p), ("d_sub_start", C.c_void_p), ("d_sub_tile", C.c_void_p),
                ("d_sub_group", C.c_void_p)]


class rj_part_out_t(C.Structure):
    _fields_ = [("side", C.c_int32), ("col", C.c_int32)]


class rj_pred_t(C.Structure):
    _fields_ = [("kind", C.c_int32), ("op", C.c_int32), ("column", C.c_uint32), ("lit_type", C.c_int32),
                ("rhs_i", C.c_int64), ("rhs_d", C.c_double), ("rhs_s", C.c_char_p), ("rhs_s_len", C.c_uint64)]


class rj_stage_stat_t(C.Structure):
    _fields_ = [("ms", C.c_double), ("launches", C.c_uint64), ("bytes", C.c_uint64)]


RJ_ST_COUNT = 10
STAGE_NAMES = ["h2d", "row_offsets", "decode", "histogram", "scatter", "join", "gather", "encode", "d2h", "join_emit"]

_vp = C.c_void_p
_u64 = C.c_uint64
_u32 = C.c_uint32
_i32 = C.c_int32
_pvp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/rj_b200.h declares
# void* sink(void* user, uint32_t column, int32_t type, uint64_t n_pages)
rj_page_sink_t = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64)

class rj_page_alloc_t(C.Structure):
    _fields_ = [("user", C.c_void_p),
                ("new_pages", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p))),
                ("append", C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_int32, C.POINTER(C.c_void_p), C.c_uint64)),
                ("free_pages", C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)))]


PROTOTYPES = {
    "rj_ctx_create": (C.c_int, [C.c_int, _pvp]),
    "rj_ctx_create_multi": (C.c_int, [C.POINTER(C.c_int), _u32, _pvp]),
    "rj_ctx_group_size": (C.c_int, [_vp]),
    "rj_ctx_destroy": (None, [_vp]),
    "rj_last_error": (C.c_char_p, [_vp]),
    "rj_ctx_device": (C.c_int, [_vp]),
    "rj_ctx_sm_count": (C.c_int, [_vp]),
    "rj_ctx_stream": (_vp, [_vp]),
    "rj_kernel_launch_count": (_u64, []),
    "rj_ctx_set_host_threads": (C.c_int, [_vp, C.c_int]),
    "rj_execute": (C.c_int, [_vp, C.POINTER(rj_plan_t), _pvp]),
    "rj_execute_streamed": (C.c_int, [_vp, C.POINTER(rj_plan_t), _u64, rj_page_sink_t, _vp, C.POINTER(_u64)]),
    "rj_execute_pages": (C.c_int, [_vp, C.POINTER(rj_plan_t), _u64, C.POINTER(rj_page_alloc_t), C.POINTER(_u64)]),
    "rj_inputs_upload": (C.c_int, [_vp, C.POINTER(rj_table_t), _u32, _pvp]),
    "rj_inputs_adopt_device": (C.c_int, [_vp, C.POINTER(rj_table_t), _u32, _pvp]),
    "rj_inputs_adopt_dense": (C.c_int, [_vp, C.POINTER(rj_dense_table_t), _u32, _pvp]),
    "rj_inputs_free": (None, [_vp, _vp]),
    "rj_execute_resident": (C.c_int, [_vp, C.POINTER(rj_plan_t), _vp, _pvp]),
    "rj_result_num_rows": (_u64, [_vp]),
    "rj_result_num_columns": (_u32, [_vp]),
    "rj_result_column_type": (_i32, [_vp, _u32]),
    "rj_result_column_pages": (_u64, [_vp, _u32]),
    "rj_result_column_device_ptr": (_u64, [_vp, _u32]),
    "rj_result_fetch": (C.c_int, [_vp, _vp, _u32, _pvp, _vp]),
    "rj_result_free": (None, [_vp, _vp]),
    "rj_page_row_offsets": (C.c_int, [_vp, _vp, _u64, _i32, _vp, _vp, _vp]),
    "rj_decode_fixed": (C.c_int, [_vp, _vp, _u64, _i32, _vp, _vp, _vp, _vp]),
    "rj_decode_varchar": (C.c_int, [_vp, _vp, _u64, _vp, _vp, _vp, _vp]),
    "rj_radix_histogram": (C.c_int, [_vp, _vp, _vp, _u64, _i32, _i32, _i32, _vp, _vp]),
    "rj_radix_scatter": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "rj_radix_scatter_multi": (C.c_int, [_vp, _vp, _vp, _u64, _i32, _i32, _i32, _vp, C.POINTER(rj_scatter_multi_t), _vp]),
    "rj_join_keys": (C.c_int, [_vp, _vp, _vp, _u64, _vp, _vp, _u64, _i32, _u64, _vp, _vp,
                               C.POINTER(_u64), _vp]),
    "rj_gather": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _i32, _vp, _vp, _vp]),
    "rj_bitmap_to_bytes": (C.c_int, [_vp, _vp, _u64, _vp, _vp]),
    "rj_bytes_to_bitmap": (C.c_int, [_vp, _vp, _u64, _vp, _vp]),
    "rj_fixed_rows_per_page": (_u32, [_i32]),
    "rj_encode_fixed": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _i32, _vp, _vp]),
    "rj_encode_varchar_plan": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _u64, _pvp, C.POINTER(_u64), _vp]),
    "rj_encode_varchar_write": (C.c_int, [_vp, _vp, _vp, _vp]),
    "rj_encode_varchar_free": (None, [_vp, _vp]),
    "rj_gen_fixed_pages": (C.c_int, [_vp, _vp, _vp, _u64, _i32, _vp, C.POINTER(_u64), _vp]),
    "rj_dist_layout": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rj_scatter_carry": (C.c_int, [_vp, C.POINTER(rj_carry_scatter_t), _vp]),
    "rj_join_partitioned": (C.c_int, [_vp, C.POINTER(rj_part_side_t), C.POINTER(rj_part_side_t), _vp, _vp, _i32, _i32, _i32,
                                      C.POINTER(rj_part_out_t), _u32, _pvp]),
    "rj_filter_compare": (C.c_int, [_vp, _vp, _vp, _u64, _i32, _i32, C.c_int64, C.c_double, _vp, _vp]),
    "rj_filter_varchar": (C.c_int, [_vp, _vp, _vp, _vp, _u64, _i32, C.c_char_p, _u64, _vp, _vp]),
    "rj_filter_null": (C.c_int, [_vp, _vp, _u64, _i32, _vp, _vp]),
    "rj_bitmap_logic": (C.c_int, [_vp, _vp, _vp, _u64, _i32, _vp, _vp]),
    "rj_bitmap_select": (C.c_int, [_vp, _vp, _u64, _vp, C.POINTER(_u64), _vp]),
    "rj_filter_table": (C.c_int, [_vp, C.POINTER(rj_table_t), C.POINTER(rj_pred_t), _u32, _pvp]),
    "rj_tables_equal": (C.c_int, [_vp, C.POINTER(rj_table_t), C.POINTER(rj_table_t), C.POINTER(_i32), C.POINTER(_u64)]),
    "rj_varchar_descriptors": (C.c_int, [_vp, _vp, _u64, _vp, _vp]),
    "rj_profile_enable": (C.c_int, [_vp, C.c_int]),
    "rj_profile_reset": (C.c_int, [_vp]),
    "rj_profile_read": (C.c_int, [_vp, C.POINTER(rj_stage_stat_t)]),
    "rj_stage_name": (C.c_char_p, [C.c_int]),
    "rj_version": (C.c_char_p, []),
}

_lib = None


class EngineMissing(RuntimeError):
    """librj_b200.so is not built: the product has no other code path."""


def load_library(path=None):
    """dlopen librj_b200.so and bind every prototype.  Raises EngineMissing if it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise EngineMissing(
            f"{p} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback."
        )
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    if path is None:
        _lib = lib
    return lib
