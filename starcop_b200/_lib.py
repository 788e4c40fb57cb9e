"""ctypes binding of libstarcop_b200.so (the C ABI declared in include/starcop_b200.h).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os
from ctypes import c_double, c_float, c_int, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# STARCOP_LIB: load another build of the same library (A/B experiments of compile-time variants)
LIB_PATH = os.environ.get("STARCOP_LIB") or os.path.join(HERE, "lib", "libstarcop_b200.so")

SC_F32, SC_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_RELU6 = 0, 1, 2

P, I, L, F, D = c_void_p, c_int, c_int64, c_float, c_double

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "sc_abi_version": [],
    "sc_last_cuda_error": [],
    "sc_normalize_pack": [P, P, P, P, P, I, I, I, I, I, P, I, I, P, P],
    "sc_conv_fprop": [P, I, P, P, P, I, I, I, I, I, I, I, I, I, I, I, I, P],
    "sc_conv_wgrad_workspace_bytes": [I, I, I, I, I, I, I, I, I],
    "sc_conv_wgrad": [P, I, P, I, P, P, I, I, I, I, I, I, I, I, I, I, P],
    "sc_pack_weights": [P, P, I, I, I, I, I, P],
    "sc_dwconv_fprop": [P, I, P, P, I, P, P, I, P, P, I, I, I, I, I, I, P],
    "sc_dwconv_dgrad": [P, I, P, P, I, I, I, I, I, I, I, P],
    "sc_dwconv_wgrad_workspace_bytes": [I],
    "sc_dwconv_wgrad": [P, I, P, P, I, P, I, P, P, I, I, I, I, I, I, P],
    "sc_bn_partials_bytes": [I],
    "sc_bn_stats": [P, I, P, P, L, I, I, P],
    "sc_bn_finalize": [P, I, L, I, P, P, P, P, F, F, I, P, P, P, P, P],
    "sc_bn_act": [P, I, P, P, I, P, I, P, I, I, I, I, I, I, I, P],
    "sc_bn_bwd_reduce": [P, I, I, P, I, P, P, P, P, I, P, P, I, I, I, I, I, P],
    "sc_bn_bwd_apply": [P, I, I, P, I, P, P, P, P, P, I, P, I, P, I, P, P, I, I, I, I, I, P],
    "sc_add_into": [P, I, I, P, I, I, I, I, I, I, I, P],
    "sc_head_fprop": [P, I, P, P, P, I, I, I, I, I, P],
    "sc_head_bwd": [P, I, P, P, P, I, I, I, I, I, I, P],
    "sc_head_wgrad_workspace_bytes": [I],
    "sc_head_wgrad_tiled": [P, I, P, P, P, P, I, I, I, I, I, P],
    "sc_bce_loss_words": [I, L],
    "sc_bce_fused": [P, P, P, F, I, L, F, P, P, P, P, P, P, P, P, P, P, P, P],
    "sc_adam_step": [P, P, P, P, L, F, F, F, F, I, F, P],
    "sc_adam_step_dev": [P, P, P, P, L, P, F, F, F, P, F, P],
    "sc_mag1c_smem_bytes": [I, I, I],
    "sc_mag1c_filter": [P, L, P, P, I, P, P, P, I, I, I, D, I, I, P, P],
    "sc_debug_mag1c_clocks": [P],
    "sc_debug_fastdiv": [ctypes.c_uint, ctypes.c_uint],
    "sc_ratio_workspace_bytes": [I, L],
    "sc_ratio_product": [P, P, P, I, L, F, F, P, P],
    "sc_weight_mag1c": [P, P, L, P],
    "sc_mlr_workspace_bytes": [I],
    "sc_mlr_reconstruct": [P, P, P, I, I, L, P, P],
    "sc_zero_override": [P, P, L, F, P],
    "sc_emit_rescale": [P, P, P, I, I, P],
    "sc_threshold_opening": [P, F, P, P, I, I, I, P],
    "sc_threshold_sweep": [P, P, P, I, L, P, P],
    "sc_affine_warp": [P, P, P, I, I, I, I, I, P],
    "sc_srf_aggregate": [P, L, L, I, P, P, I, F, P, P],
    "sc_chain_pack": [P, P, I, I, I, I, P, P, P, P, I, L, P, I, I, P, P, P],
    "sc_tc_supported": [],
    "sc_tc_pack_weights": [P, P, I, I, I, I, I, I, I, P],
    "sc_tc_pack_weights_batch": [P, I, L, P],
    "sc_tc_cin_pad": [I],
    "sc_tc_conv_fprop": [P, I, P, P, I, P, P, I, I, I, I, I, I, I, I, I, P],
    "sc_tc_halo_cin_pad": [I],
    "sc_stem_s2d": [P, I, I, P, I, I, I, P],
    "sc_stem_s2d_pack_weights": [P, P, I, I, P],
    "sc_stem_s2d_unpack_grad": [P, P, I, I, P],
    "sc_tc_halo_supported": [I, I],
    "sc_tc_conv3x3_halo": [P, I, P, P, I, P, P, I, I, I, I, I, I, P],
    "sc_tc_conv_wgrad_workspace_bytes": [I, I, I, I, I, I, I, I],
    "sc_tc_wgrad_halo_supported": [I, I, I, I, I],
    "sc_tc_wgrad_halo_workspace_bytes": [I, I, I, I, I],
    "sc_tc_wgrad_halo": [P, I, P, I, P, P, I, I, I, I, I, P],
    "sc_tc_conv_wgrad": [P, I, P, I, P, P, I, I, I, I, I, I, I, I, P],
}
_RESTYPES = {"sc_last_cuda_error": ctypes.c_char_p, "sc_mag1c_smem_bytes": c_int64, "sc_bn_partials_bytes": c_int64, "sc_mlr_workspace_bytes": c_int64, "sc_dwconv_wgrad_workspace_bytes": c_int64, "sc_head_wgrad_workspace_bytes": c_int64,
             "sc_ratio_workspace_bytes": c_int64, "sc_conv_wgrad_workspace_bytes": c_int64,
             "sc_tc_conv_wgrad_workspace_bytes": c_int64, "sc_tc_wgrad_halo_workspace_bytes": c_int64, "sc_bce_loss_words": c_int64,
             "sc_debug_fastdiv": ctypes.c_uint}

_lib = None


class StarcopB200Error(RuntimeError):
    pass


def load():
    """dlopen the in-tree library (never a site-packages copy) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StarcopB200Error(
            f"{LIB_PATH} is missing: build it with `python -m starcop_b200.build` "
            "(there is no CPU or PyTorch fallback for the starcop_b200 hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


_ERR = {-1: "SC_ERR_BAD_ARG", -2: "SC_ERR_CUDA", -3: "SC_ERR_UNSUPPORTED", -4: "SC_ERR_NO_DEVICE"}


def check(status, what):
    if status != 0:
        detail = ""
        if status == -2:
            detail = ": " + load().sc_last_cuda_error().decode()
        raise StarcopB200Error(f"{what} failed with {_ERR.get(status, status)}{detail}")


_launches = 0


def launch_count():
    """Number of C-ABI kernel-launching calls made so far (each launches >= 1 of our kernels)."""
    return _launches


def call(name, *args):
    global _launches
    _launches += 1
    check(getattr(load(), name)(*args), name)
