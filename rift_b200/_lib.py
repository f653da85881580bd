"""ctypes binding of librift_b200.so (C ABI in include/rift_b200.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, the product
path raises.  (The CPU oracle under oracle/ is test infrastructure and is never imported here.)
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librift_b200.so")

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)
c_u8_p = C.POINTER(C.c_uint8)
c_i8_p = C.POINTER(C.c_int8)
c_int_p = C.POINTER(C.c_int)
c_ll_p = C.POINTER(C.c_longlong)


class ModelConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "dim", "num_heads", "encoder_depth", "decoder_depth", "num_modes", "history_steps", "future_steps",
        "state_channel", "ref_points", "value_hidden0", "value_hidden1")]


class ParamEntry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("offset", C.c_longlong), ("numel", C.c_longlong), ("trainable", C.c_int)]


class Batch(C.Structure):
    _fields_ = [
        ("bs", C.c_int), ("A", C.c_int), ("Mp", C.c_int), ("P", C.c_int), ("R", C.c_int), ("Pr", C.c_int),
        ("agent_T", C.c_int),
        ("agent_position", C.c_void_p), ("agent_heading", C.c_void_p), ("agent_velocity", C.c_void_p),
        ("agent_shape", C.c_void_p), ("agent_category", C.c_void_p), ("agent_valid_mask", C.c_void_p),
        ("map_point_position", C.c_void_p), ("map_point_vector", C.c_void_p), ("map_point_orientation", C.c_void_p),
        ("map_polygon_center", C.c_void_p), ("map_polygon_type", C.c_void_p), ("map_polygon_on_route", C.c_void_p),
        ("map_polygon_tl_status", C.c_void_p), ("map_polygon_has_speed_limit", C.c_void_p),
        ("map_polygon_speed_limit", C.c_void_p), ("map_valid_mask", C.c_void_p),
        ("ref_position", C.c_void_p), ("ref_vector", C.c_void_p), ("ref_orientation", C.c_void_p),
        ("ref_valid_mask", C.c_void_p),
        ("current_state", C.c_void_p), ("cs_stride", C.c_int),
        ("ref_valid_mask_global", C.c_void_p), ("bs_global", C.c_int), ("b_offset", C.c_int),
    ]


class GatherField(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("item_stride_bytes", C.c_longlong), ("copy_bytes", C.c_longlong),
                ("dst_stride_bytes", C.c_longlong)]


class Outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "probability", "trajectory", "prediction", "hidden", "ref_free_trajectory", "candidate_trajectories",
        "r_padding_mask")]


FWD_SAVE_FOR_BACKWARD = 1
GEMM_SIMT = 2

_lib = None

# name -> (restype, argtypes); every symbol of include/rift_b200.h
_V = C.c_void_p
_SIGNATURES = {
    "rift_b200_last_error": (C.c_char_p, []),
    "rift_b200_version": (C.c_int, []),
    "rift_b200_launch_count": (C.c_longlong, []),
    "rift_b200_debug_gemm_trace": (None, [_V]),
    "rift_b200_debug_fused_trace": (None, [_V]),
    "rift_b200_create": (C.c_int, [C.POINTER(ModelConfig), C.POINTER(ParamEntry), C.c_int, C.POINTER(_V)]),
    "rift_b200_destroy": (None, [_V]),
    "rift_b200_bind_arena": (C.c_int, [_V, _V, _V, C.c_longlong]),
    "rift_b200_workspace_bytes": (C.c_size_t, [_V, C.POINTER(Batch)]),
    "rift_b200_forward": (C.c_int, [_V, C.POINTER(Batch), C.POINTER(Outputs), _V, C.c_size_t, C.c_int, _V]),
    "rift_b200_backward": (C.c_int, [_V, C.POINTER(Batch), _V, _V, C.c_size_t, C.c_int, _V]),
    "rift_b200_objective_scratch_bytes": (C.c_size_t, [C.c_int]),
    "rift_b200_group_objective": (C.c_int, [C.c_int, _V, _V, _V, _V, _V, _V, C.c_int, C.c_int, C.c_int, C.c_float,
                                            C.c_float, C.c_float, C.c_float, _V, _V, _V, C.c_int, _V]),
    "rift_b200_action_objective": (C.c_int, [C.c_int, _V, _V, _V, _V, _V, C.c_int, C.c_int, C.c_int, C.c_float,
                                             C.c_float, C.c_float, _V, _V, _V, _V, _V, _V]),
    "rift_b200_teacher_objective": (C.c_int, [_V, _V, _V, _V, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _V,
                                              _V, _V, C.c_int, _V, _V]),
    "rift_b200_smooth_l1": (C.c_int, [_V, _V, C.c_int, C.c_float, _V, _V, _V]),
    "rift_b200_group_advantage": (C.c_int, [_V, _V, C.c_longlong, C.c_int, _V, _V]),
    "rift_b200_gae": (C.c_int, [_V, _V, _V, _V, _V, C.c_int, C.c_float, C.c_float, _V, _V, _V, _V]),
    "rift_b200_discounted_return": (C.c_int, [_V, _V, C.c_int, C.c_float, _V, _V]),
    "rift_b200_optim_scratch_bytes": (C.c_size_t, []),
    "rift_b200_clip_adamw": (C.c_int, [_V, _V, _V, _V, C.c_longlong, C.c_longlong, _V, C.c_float, C.c_float, C.c_float,
                                       C.c_float, C.c_float, C.c_float, C.c_int, _V, _V, _V]),
    "rift_b200_clip_adamw_dev": (C.c_int, [_V, _V, _V, _V, C.c_longlong, C.c_longlong, _V, C.c_float, _V, C.c_float,
                                           C.c_float, C.c_float, C.c_float, _V, _V, _V]),
    "rift_b200_refresh_weights": (C.c_int, [_V, _V]),
    "rift_b200_eval_ref_line_info": (C.c_int, [_V, C.c_int, C.c_int, C.c_int, _V, _V, _V, _V, _V, _V]),
    "rift_b200_eval_center_rollout": (C.c_int, [_V, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_float, _V, _V, _V, C.c_int,
                                                _V, _V, _V, _V, _V, _V, _V, _V]),
    "rift_b200_eval_other_rollout": (C.c_int, [_V, _V, _V, _V, _V, C.c_int, C.c_int, C.c_int, C.c_double, _V, _V]),
    "rift_b200_eval_returns": (C.c_int, [_V, _V, _V, _V, _V, _V, _V, _V, _V, C.c_int, C.c_int, _V, C.c_int, C.c_int,
                                         C.POINTER(C.c_double), C.c_int, C.c_double, _V, _V, _V, _V]),
    "rift_b200_gather_fields": (C.c_int, [C.POINTER(GatherField), C.c_int, _V, C.c_int, _V]),
    "rift_b200_op_add_inplace": (C.c_int, [_V, _V, C.c_longlong, _V]),
    "rift_b200_op_linear": (C.c_int, [_V, C.c_int, C.c_int, _V, _V, C.c_int, C.c_int, _V, _V, C.c_int, _V]),
    "rift_b200_weight_cache_bytes": (C.c_size_t, [_V]),
    "rift_b200_bind_weight_cache": (C.c_int, [_V, _V, C.c_size_t]),
    "rift_b200_params_updated": (C.c_int, [_V, C.c_int]),
    "rift_b200_op_fused_mlp_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "rift_b200_op_fused_mlp": (C.c_int, [_V, C.c_int, C.c_int, C.c_int, C.c_int, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V,
                                         _V, _V, _V, C.c_size_t, C.c_int, _V]),
    "rift_b200_op_wgrad_group": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                           C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _V, C.c_size_t, _V]),
    "rift_b200_op_wgrad_tc_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "rift_b200_op_wgrad_tc": (C.c_int, [_V, _V, C.c_int, C.c_int, C.c_int, _V, _V, C.c_int, _V, C.c_size_t, _V]),
    "rift_b200_op_linear_tc_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "rift_b200_op_linear_tc": (C.c_int, [_V, C.c_int, C.c_int, _V, _V, C.c_int, C.c_int, _V, _V, _V, C.c_size_t,
                                         C.c_int, _V]),
    "rift_b200_op_linear_tc_full": (C.c_int, [_V, C.c_int, C.c_int, _V, _V, C.c_int, C.c_int, _V, _V, C.c_float, _V, _V, _V, _V,
                                              C.c_size_t, _V]),
    "rift_b200_op_gemm": (C.c_int, [_V, C.c_longlong, C.c_longlong, _V, C.c_longlong, C.c_longlong, _V, C.c_longlong,
                                    C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _V, C.c_int, _V]),
    "rift_b200_op_layernorm": (C.c_int, [_V, C.c_int, C.c_int, _V, _V, C.c_int, _V, _V, _V, _V]),
    "rift_b200_op_layernorm_bwd": (C.c_int, [_V, _V, C.c_int, C.c_int, _V, _V, _V, _V, _V, _V, _V, _V, _V]),
    "rift_b200_op_attention": (C.c_int, [_V, C.c_int, C.c_int, C.c_int, C.c_int, _V, _V, _V]),
    "rift_b200_op_nat_attention": (C.c_int, [_V, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _V, _V, _V]),
    "rift_b200_op_attention_bwd": (C.c_int, [_V, _V, C.c_int, C.c_int, C.c_int, C.c_int, _V, _V, _V, _V, _V]),
    "rift_b200_op_nat_attention_bwd": (C.c_int, [_V, _V, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _V, _V, _V, _V]),
    "rift_b200_op_act_bwd": (C.c_int, [_V, _V, C.c_longlong, C.c_int, _V]),
    "rift_b200_op_colsum": (C.c_int, [_V, C.c_int, C.c_int, _V, C.c_int, _V, _V]),
    "rift_b200_op_masked_maxpool": (C.c_int, [_V, _V, C.c_int, C.c_int, C.c_int, _V, _V, _V]),
}


def declared_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m rift_b200.build` "
                "(rift_b200 has no CPU or PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().rift_b200_last_error()
        raise RuntimeError(f"rift_b200 {what} failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
