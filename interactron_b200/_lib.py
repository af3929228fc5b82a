"""ctypes binding of libinteractron_b200.so (the C ABI in include/interactron_b200.h).

The library is the only arithmetic backend of the package: there is no CPU or
eager-PyTorch fallback.  `load()` raises if the shared object is missing, and every
wrapper raises `ItnError` with `itn_last_error()` on a non-zero return code.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ITN_LIB: load another build of the same ABI (the trace / experiment builds of csrc/Makefile)
LIB_PATH = os.environ.get("ITN_LIB") or os.path.join(_HERE, "libinteractron_b200.so")


class ItnError(RuntimeError):
    pass


class Operand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("major", C.c_int), ("ld", C.c_longlong),
                ("sb0", C.c_longlong), ("sb1", C.c_longlong)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("nb0", C.c_int), ("nb1", C.c_int),
        ("A", Operand), ("B", Operand),
        ("C", C.c_void_p), ("ldc", C.c_longlong), ("c_sb0", C.c_longlong), ("c_sb1", C.c_longlong),
        ("bias", C.c_void_p), ("bias_sb0", C.c_longlong), ("bias_sb1", C.c_longlong),
        ("residual", C.c_void_p), ("ldr", C.c_longlong), ("r_sb0", C.c_longlong), ("r_sb1", C.c_longlong),
        ("aux", C.c_void_p), ("ldaux", C.c_longlong), ("aux_sb0", C.c_longlong), ("aux_sb1", C.c_longlong),
        ("C2", C.c_void_p), ("ldc2", C.c_longlong), ("c2_sb0", C.c_longlong), ("c2_sb1", C.c_longlong),
        ("alpha", C.c_float), ("act", C.c_int), ("epi", C.c_int), ("accumulate", C.c_int),
        ("round_out", C.c_int), ("precision", C.c_int), ("act_pos", C.c_int), ("c_pad", C.c_int),
        ("conv_kh", C.c_int), ("conv_kw", C.c_int), ("conv_stride", C.c_int), ("conv_pad", C.c_int), ("conv_dil", C.c_int),
        ("conv_n", C.c_int), ("conv_h", C.c_int), ("conv_w", C.c_int), ("conv_c", C.c_int), ("conv_ho", C.c_int),
        ("conv_wo", C.c_int),
        ("B_lo", C.c_void_p),
    ]


class AttentionDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("nh", C.c_int), ("hd", C.c_int), ("Lq", C.c_int), ("Lk", C.c_int),
        ("scale", C.c_float),
        ("q", C.c_void_p), ("q_ld", C.c_longlong), ("q_sb", C.c_longlong),
        ("k", C.c_void_p), ("k_ld", C.c_longlong), ("k_sb", C.c_longlong),
        ("v", C.c_void_p), ("v_ld", C.c_longlong), ("v_sb", C.c_longlong),
        ("key_mask", C.c_void_p),
        ("o", C.c_void_p), ("o_ld", C.c_longlong), ("o_sb", C.c_longlong),
        ("lse", C.c_void_p),
        ("d_o", C.c_void_p), ("do_ld", C.c_longlong), ("do_sb", C.c_longlong),
        ("dq", C.c_void_p), ("dq_ld", C.c_longlong), ("dq_sb", C.c_longlong),
        ("dk", C.c_void_p), ("dk_ld", C.c_longlong), ("dk_sb", C.c_longlong),
        ("dv", C.c_void_p), ("dv_ld", C.c_longlong), ("dv_sb", C.c_longlong),
        ("delta", C.c_void_p),
        ("drop_p", C.c_float), ("drop_seed", C.c_void_p), ("drop_site", C.c_uint),
    ]


# name -> (restype, argtypes); mirrors include/interactron_b200.h one to one.
_P, _LL, _I, _F = C.c_void_p, C.c_longlong, C.c_int, C.c_float
SIGNATURES = {
    "itn_last_error": (C.c_char_p, []),
    "itn_version": (C.c_char_p, []),
    "itn_launch_count": (_LL, []),
    "itn_tf32_residual": (_I, [_P, _P, _LL, _P]),
    "itn_gemm_tf32": (_I, [C.POINTER(GemmDesc), _P]),
    "itn_gemm_tf32_supported": (_I, [C.POINTER(GemmDesc)]),
    "itn_gemm_simt": (_I, [C.POINTER(GemmDesc), _P]),
    "itn_attention_supported": (_I, [C.POINTER(AttentionDesc)]),
    "itn_attention_fwd": (_I, [C.POINTER(AttentionDesc), _P]),
    "itn_attention_bwd": (_I, [C.POINTER(AttentionDesc), _P]),
    "itn_layernorm_fwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _LL, _I, _I, _LL, _F, _P]),
    "itn_layernorm_fwd_plus": (_I, [_P, _P, _P, _P, _P, _P, _LL, _I, _I, _LL, _F, _P, _P, _LL, _LL, _LL, _P]),
    "itn_layernorm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _LL, _I, _I, _LL, _LL, _P]),
    "itn_layernorm_bwd_fused_workspace": (_LL, [_LL, _I, _I]),
    "itn_layernorm_bwd_fused": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _LL, _I, _I, _LL, _LL, _LL, _P, _LL, _P]),
    "itn_softmax_fwd": (_I, [_P, _LL, _I, _LL, _F, _P, _LL, _I, _P]),
    "itn_softmax_bwd": (_I, [_P, _P, _LL, _I, _LL, _F, _I, _P]),
    "itn_colsum": (_I, [_P, _P, _I, _LL, _I, _LL, _LL, _P]),
    "itn_add": (_I, [_P, _P, _P, _LL, _LL, _LL, _LL, _I, _P]),
    "itn_copy2d": (_I, [_P, _LL, _P, _LL, _LL, _I, _I, _P]),
    "itn_transpose": (_I, [_P, _P, _I, _I, _I, _LL, _LL, _P]),
    "itn_round_tf32": (_I, [_P, _P, _LL, _P]),
    "itn_sigmoid_fwd": (_I, [_P, _P, _LL, _P]),
    "itn_sigmoid_bwd": (_I, [_P, _P, _P, _LL, _P]),
    "itn_l2norm_fwd_bwd": (_I, [_P, _P, _P, _I, _I, _P]),
    "itn_sgd_clip_update": (_I, [_P, _LL, _P, _P, _P, _P, _I, _LL, _F, _F, _P]),
    "itn_pos_embed_sine": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "itn_im2col_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _LL, _P]),
    "itn_maxpool3x3s2_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "itn_matcher_cost": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _F, _P]),
    "itn_criterion_scratch_bytes": (_LL, [_I, _I, _I]),
    "itn_layernorm_fwd_jvp": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _LL, _I, _I, _LL, _I, _LL, _P]),
    "itn_layernorm_bwd_jvp": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _LL, _I, _I, _LL, _I, _LL, _P]),
    "itn_softmax_bwd_jvp": (_I, [_P, _P, _P, _P, _LL, _I, _LL, _F, _P]),
    "itn_mask_mul": (_I, [_P, _P, _LL, _P]),
    "itn_mul_mask_u8": (_I, [_P, _P, _F, _P, _LL, _P]),
    "itn_gelu_grad_dual": (_I, [_P, _P, _P, _P, _P, _P, _LL, _P]),
    "itn_sigmoid_bwd_jvp": (_I, [_P, _P, _P, _P, _P, _LL, _P]),
    "itn_l2norm_jvp": (_I, [_P, _P, _P, _P, _P, _I, _I, _P]),
    "itn_sumsq_partials": (_I, [_P, _LL, _P, _I, C.POINTER(C.c_int), _P]),
    "itn_clip_adam_step": (_I, [_P, _P, _P, _P, _LL, _P, _I, _F, C.c_double, C.c_double, C.c_double, C.c_double, _I, _I, _P, _P]),
    "itn_dropout": (_I, [_P, _LL, _P, _LL, _P, _LL, _LL, _I, _LL, _F, _P, C.c_uint, _P]),
    "itn_ckpt_accumulate": (_I, [_P, _P, _LL, _F, _I, _P]),
    "itn_detect_postprocess": (_I, [_P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P]),
    "itn_criterion": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _F, _F, _P, _P, _P, _P, _P]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ItnError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
            " (or `make -C interactron_b200/csrc`). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise ItnError(f"itn error {rc}: {load().itn_last_error().decode()}")
