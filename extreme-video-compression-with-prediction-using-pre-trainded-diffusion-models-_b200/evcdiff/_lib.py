"""ctypes binding of libevcdiff.so (the C ABI declared in include/evcdiff.h).

The library is the only compute path of this package: if it is missing or a call fails the caller gets an
exception -- there is no CPU or PyTorch fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# EVC_LIB: another build of the same library (tools/build_prof.sh: the wait-time probe build, experiments only)
LIB_PATH = os.environ.get("EVC_LIB") or os.path.join(_HERE, "lib", "libevcdiff.so")

EVC_OUT_BF16_ROWS, EVC_OUT_F32_ROWS, EVC_OUT_BF16_T, EVC_OUT_F32_T = 0, 1, 2, 3


class EvcError(RuntimeError):
    pass


class Tensor4(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int32), ("W", C.c_int32), ("H", C.c_int32), ("B", C.c_int32),
                ("stride_w", C.c_int64), ("stride_h", C.c_int64), ("stride_b", C.c_int64)]


class GemmDesc(C.Structure):
    _fields_ = [("n_seg", C.c_int32), ("a", Tensor4 * 3), ("taps", C.c_int32 * 3), ("w", C.c_void_p),
                ("w_rows", C.c_int32), ("w_k", C.c_int32), ("w_batches", C.c_int32),
                ("w_row_stride", C.c_int64), ("w_batch_stride", C.c_int64),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("bn", C.c_int32),
                ("out", C.c_void_p), ("out_mode", C.c_int32), ("out_ld", C.c_int64), ("out_bs", C.c_int64),
                ("bias", C.c_void_p), ("resid", C.c_void_p), ("resid_ld", C.c_int64),
                ("alpha", C.c_float), ("max_ctas", C.c_int32), ("stats", C.c_void_p), ("stride", C.c_int32), ("cta_group", C.c_int32),
                ("a_lo", Tensor4 * 3), ("w_lo", C.c_void_p), ("out_lo", C.c_void_p), ("resid_lo", C.c_void_p),
                ("gn_ss", C.c_void_p), ("gn_ticket", C.c_void_p), ("gn_eps", C.c_float), ("gn_groups", C.c_int32),
                ("gn_adagn", C.c_int32), ("split_k", C.c_int32), ("sk_ws", C.c_void_p), ("sk_ws_bytes", C.c_int64),
                ("sk_ticket", C.c_void_p)]


class AttnDesc(C.Structure):
    _fields_ = [("qk", C.c_void_p), ("qk_ld", C.c_int64), ("vT", C.c_void_p), ("vT_ld", C.c_int64),
                ("out", C.c_void_p), ("out_ld", C.c_int64), ("B", C.c_int32), ("N", C.c_int32), ("C", C.c_int32),
                ("heads", C.c_int32), ("scale", C.c_float), ("v", C.c_void_p), ("v_ld", C.c_int64)]


class StepCoef(C.Structure):
    _fields_ = [("mode", C.c_int32), ("clip", C.c_int32), ("k0", C.c_float), ("k1", C.c_float),
                ("c_x0", C.c_float), ("c_x", C.c_float), ("c_eps", C.c_float), ("c_noise", C.c_float)]


class PndmCoef(C.Structure):
    _fields_ = [("n_e", C.c_int32), ("clip", C.c_int32), ("w", C.c_float * 4), ("w_scale", C.c_float),
                ("d", C.c_float), ("p", C.c_float), ("q", C.c_float)]


ABI_STRUCTS = (Tensor4, GemmDesc, AttnDesc, StepCoef, PndmCoef)  # order = evc_struct_size ids

# name -> (restype, argtypes); must list every symbol include/evcdiff.h declares (tests check this).
_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "evc_version": (C.c_int, []),
    "evc_last_error": (C.c_char_p, []),
    "evc_launch_count": (C.c_int64, []),
    "evc_struct_size": (C.c_int64, [C.c_int]),
    "evc_set_pdl": (None, [C.c_int]),
    "evc_gemm_plan_create": (C.c_int, [C.POINTER(GemmDesc), C.POINTER(C.c_void_p)]),
    "evc_gemm_plan_launch": (C.c_int, [_vp, _vp, _vp]),
    "evc_gemm_plan_launch_gn": (C.c_int, [_vp, _vp, _vp, _vp]),
    "evc_gemm_plan_launch_ex": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "evc_gemm_fault_count": (C.c_int64, []),
    "evc_gemm_plan_destroy": (None, [_vp]),
    "evc_gemm_plan_flops": (C.c_double, [_vp]),
    "evc_gemm_plan_cta_group": (C.c_int, [_vp]),
    "evc_attn_plan_create": (C.c_int, [C.POINTER(AttnDesc), C.POINTER(C.c_void_p)]),
    "evc_attn_plan_launch": (C.c_int, [_vp, _vp]),
    "evc_attn_plan_destroy": (None, [_vp]),
    "evc_attn_plan_flops": (C.c_double, [_vp]),
    "evc_gn_stats_workspace": (C.c_int, [_i32, _i32, _i32, C.POINTER(C.c_int64)]),
    "evc_gn_stats": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _i64, _vp]),
    "evc_gn_apply": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _f32, _vp, _i32, _i32, _vp, _vp]),
    "evc_gn_stats_split": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _i64, _vp]),
    "evc_gn_apply_split": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _f32, _vp, _i32, _i32, _vp,
                                     _vp, _vp]),
    "evc_fir_resample_split": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evc_softmax_rows_split": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "evc_pack_nchw_split": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _f32, _f32, _vp, _vp, _i32, _i32, _vp]),
    "evc_fir_resample": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evc_gn_fir": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _f32, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "evc_nearest_up2": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "evc_softmax_rows": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "evc_timestep_embedding": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "evc_linear_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "evc_pack_nchw": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _f32, _f32, _vp, _i32, _i32, _vp]),
    "evc_fill_zero": (C.c_int, [_vp, _i64, _vp]),
    "evc_sampler_update": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, C.POINTER(StepCoef), _vp]),
    "evc_pndm_update": (C.c_int, [_vp, C.POINTER(C.c_void_p), _vp, _vp, _vp, _i32, _i32, _i32, _i32,
                                  C.POINTER(PndmCoef), _vp]),
    "evc_inverse_transform": (C.c_int, [_vp, _vp, _i64, _vp]),
    "evc_frames_to_uint8": (C.c_int, [_vp, _vp, _i64, _vp]),
    "evc_frame_psnr": (C.c_int, [_vp, _vp, _i32, _i64, C.c_double, _vp, _vp]),
    "evc_accept_prefix": (C.c_int, [_vp, _i32, _i32, C.c_double, _i32, _vp, _vp]),
}

_lib = None


def load():
    """Load libevcdiff.so (once). Raises EvcError when the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EvcError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` (nvcc, sm_100a). "
            "evcdiff has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for which, struct in enumerate(ABI_STRUCTS):
        if lib.evc_struct_size(which) != C.sizeof(struct):
            raise EvcError(f"ABI mismatch: {struct.__name__} is {C.sizeof(struct)} bytes here, "
                           f"{lib.evc_struct_size(which)} in {LIB_PATH} (rebuild the library)")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().evc_last_error()
        raise EvcError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
