"""ctypes binding of libmrblip_b200.so (include/mrblip_b200.h).  There is no fallback: if the
library is missing or a call fails, the product path raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmrblip_b200%s.so" % os.environ.get("MRB_LIB_VARIANT", ""))   # variants: A/B builds (build.py)

_p, _ll, _i, _f, _u = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float, ctypes.c_uint

# name -> argument ctypes, in header order (must match include/mrblip_b200.h)
SIGNATURES = {
    "mrb_gemm": [_p, _ll, _p, _ll, _i, _i, _i, _i, _p, _i, _p, _ll, _p, _i, _ll, _i, _i, _p],
    "mrb_gemm_splitk_plan": [_i, _i, _i, _i, _i, _i, _p, _p, _p],
    "mrb_gemm_sm_limit": [_i],
    "mrb_gemm_splitk": [_p, _ll, _p, _ll, _i, _i, _i, _i, _p, _i, _p, _ll, _p, _i, _ll, _i, _i, _p, _ll, _i, _p],
    "mrb_attention_fwd": [_p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _i, _i, _i, _i, _i, _i, _f, _p, _i, _i,
                          _p, _i, _i, _i, _p, _p],
    "mrb_attention_row": [_p, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _i, _i, _i, _i, _i, _f, _p],
    "mrb_attention_vit": [_p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _i, _i, _i, _i, _i, _f, _p],
    "mrb_attention_fwd_tc": [_p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _i, _i, _i, _i, _i, _i, _f, _p, _i, _i,
                             _p, _i, _i, _i, _p, _p],
    "mrb_attention_bwd": [_p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _p, _p,
                          _i, _i, _i, _i, _i, _i, _f, _p, _i, _i, _p, _i, _i, _p, _p, _p],
    "mrb_attention_bwd_tc": [_p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _p, _p,
                             _i, _i, _i, _i, _i, _i, _f, _p, _i, _i, _p, _i, _i, _p, _p, _p],
    "mrb_norm": [_p, _p, _p, _p, _f, _i, _i, _i, _p, _p, _i, _ll, _p, _p],
    "mrb_rmsnorm_bwd": [_p, _p, _p, _i, _ll, _p, _i, _f, _i, _i, _p, _p],
    "mrb_lora_up_add": [_p, _ll, _p, _i, _i, _i, _i, _p, _p],
    "mrb_patchify": [_p, _p, _i, _i, _i, _i, _i, _p],
    "mrb_patchify_u8": [_p, _p, _i, _i, _i, _i, _i, _f, _f, _f, _f, _f, _f, _p],
    "mrb_cls_pos": [_p, _p, _p, _i, _i, _i, _p],
    "mrb_gated_gelu_fwd": [_p, _p, _i, _i, _ll, _i, _p],
    "mrb_gated_gelu_bwd": [_p, _p, _ll, _p, _ll, _i, _i, _i, _p],
    "mrb_gather_rows": [_p, _p, _p, _p, _i, _i, _p],
    "mrb_scatter_frames": [_p, _p, _p, _i, _i, _p],
    "mrb_group_mean": [_p, _p, _i, _i, _i, _p],
    "mrb_group_mean_bwd": [_p, _p, _i, _i, _i, _p],
    "mrb_cross_entropy": [_p, _p, _i, _i, _p, _p, _i, _ll, _f, _p, _p],
    "mrb_lora_down": [_p, _ll, _p, _i, _i, _i, _i, _p],
    "mrb_skinny_wgrad": [_p, _ll, _p, _ll, _i, _i, _p, _i, _i, _p],
    "mrb_skinny_wgrad_tc": [_p, _ll, _p, _ll, _i, _i, _p, _i, _i, _p],
    "mrb_skinny_wgrad_tc2": [_p, _ll, _p, _ll, _i, _i, _p, _p, _i, _i, _p],
    "mrb_lora_pack": [_p, _i, _i, _i, _p],
    "mrb_small_down": [_p, _ll, _p, _ll, _i, _i, _p, _ll, _i, _p],
    "mrb_cast_f32_to_h": [_p, _p, _ll, _i, _p],
    "mrb_cast2d_f32_to_h": [_p, _ll, _p, _ll, _i, _i, _i, _p],
    "mrb_transpose16": [_p, _ll, _p, _ll, _i, _i, _p],
    "mrb_colsum": [_p, _i, _i, _p, _p],
    "mrb_axpby": [_p, _p, _ll, _f, _f, _p],
    "mrb_attention_fwd_drop": [_p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _i, _i, _i, _i, _i, _i, _f, _p, _i, _i,
                               _p, _i, _i, _i, _p, _p, _u, _f, _p],
    "mrb_attention_fwd_tc_drop": [_p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _i, _i, _i, _i, _i, _i, _f, _p, _i, _i,
                               _p, _i, _i, _i, _p, _p, _u, _f, _p],
    "mrb_attention_bwd_drop": [_p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _p, _p,
                               _i, _i, _i, _i, _i, _i, _f, _p, _i, _i, _p, _i, _i, _p, _p, _p, _u, _f, _p],
    "mrb_attention_bwd_tc_drop": [_p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _ll, _ll, _p, _p, _p,
                               _i, _i, _i, _i, _i, _i, _f, _p, _i, _i, _p, _i, _i, _p, _p, _p, _u, _f, _p],
    "mrb_dropout": [_p, _ll, _p, _ll, _i, _i, _i, _i, _p, _u, _f, _p],
    "mrb_dropout_add": [_p, _p, _p, _i, _i, _p, _u, _f, _p],
    "mrb_dropout_add_norm": [_p, _p, _p, _f, _i, _i, _p, _i, _ll, _p, _p, _u, _f, _p],
    "mrb_rmsnorm_bwd_drop": [_p, _p, _p, _i, _ll, _f, _i, _i, _p, _p, _i, _ll, _p, _u, _f, _p],
    "mrb_gated_gelu_fwd_drop": [_p, _p, _i, _i, _ll, _i, _p, _u, _f, _p],
    "mrb_gated_gelu_bwd_drop": [_p, _p, _ll, _p, _ll, _i, _i, _i, _p, _u, _f, _p],
    "mrb_lora_down_drop": [_p, _ll, _p, _ll, _i, _i, _i, _p, _ll, _i, _p, _u, _f, _p],
    "mrb_lora_wgrad_drop": [_p, _ll, _p, _ll, _i, _i, _p, _i, _p, _u, _f, _p],
    "mrb_lora_dx_drop": [_p, _ll, _p, _ll, _i, _p, _ll, _i, _i, _i, _i, _p, _u, _f, _p],
}

_lib = None
launch_count = 0     # kernels launched through the C ABI so far (bench.py reports the delta as gpu_launches)
_KERNELS_PER_CALL = {"mrb_attention_bwd": 3, "mrb_attention_bwd_tc": 3, "mrb_attention_bwd_drop": 3, "mrb_attention_bwd_tc_drop": 3}


class MrbError(RuntimeError):
    pass


def load():
    """Load the shared library (after torch, so both share one libcudart.so.12)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MrbError("libmrblip_b200.so is not built (%s): run `python -m mr_blip_b200.build`; "
                       "there is no CPU / PyTorch fallback for the hot path" % LIB_PATH)
    import torch  # noqa: F401  (loads libcudart.so.12 first)
    lib = ctypes.CDLL(LIB_PATH)
    lib.mrb_last_error.restype = ctypes.c_char_p
    lib.mrb_abi_version.restype = _i
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _i
    _lib = lib
    return lib


PROFILE = None       # set to a list: every call appends (name, start_event, end_event) (tools/step_breakdown.py)


def call(name, *args):
    global launch_count
    lib = _lib or load()
    if PROFILE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        PROFILE.append((name, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    launch_count += _KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        kind = {-1: "bad argument", -2: "CUDA error", -3: "unsupported shape"}.get(rc, "error")
        raise MrbError("%s failed: %s (%d) %s" % (name, kind, rc, lib.mrb_last_error().decode()))
