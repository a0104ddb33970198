"""ctypes binding of the C-ABI library ``libvae_b200.so`` (include/vae_b200.h).

There is no CPU fallback: if the library is missing, or an op is called
without a CUDA device, a RuntimeError is raised (the reference's own error
convention for bad shapes is torch's RuntimeError).
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VAE_B200_LIB: another build of the same library (A/B timing of compile-time variants on one GPU box)
LIB_PATH = os.environ.get("VAE_B200_LIB") or os.path.join(_HERE, "libvae_b200.so")

_vp, _i, _ll, _f, _ull = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_ulonglong

# name -> argtypes (restype is int unless listed in _RESTYPES)
PROTOTYPES = {
    "vae_abi_version": [],
    "vae_set_reserved_sms": [_i],
    "vae_set_zero_arena": [_vp, _ll],
    "vae_last_error": [],
    "vae_has_umma": [],
    "vae_pack_conv_weight": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "vae_conv_fprop": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vae_conv_dgrad": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vae_conv_wgrad": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vae_set_umma": [_i],
    "vae_umma_set_trace": [_vp],
    "vae_umma_packed_elems": [_i, _i, _i, _i],
    "vae_umma_pack_weight": [_vp, _vp, _i, _i, _i, _i, _vp],
    "vae_conv_umma": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vae_set_edge_tc": [_i],
    "vae_set_umma_tmat": [_i],
    "vae_set_umma_tma": [_i],
    "vae_umma_tma_eligible": [_i, _i, _i, _i, _i, _i, _i],
    "vae_umma_tma_packed_elems": [_i, _i, _i, _i],
    "vae_umma_tma_pack_weight": [_vp, _vp, _i, _i, _i, _i, _vp],
    "vae_conv_umma_tma": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vae_set_umma_rw": [_i],
    "vae_umma_rw_eligible": [_i, _i, _i, _i, _i, _i, _i],
    "vae_umma_rw_packed_elems": [_i, _i],
    "vae_umma_rw_pack_weight": [_vp, _vp, _i, _i, _i, _vp],
    "vae_conv_umma_rw": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vae_conv_umma_rw_act": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vae_conv_umma_tma_act": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vae_conv_umma_rw_bnbwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vae_conv_umma_tma_bnbwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vae_set_wgrad_tma": [_i],
    "vae_wgrad_tma_eligible": [_i, _i, _i, _i, _i, _i],
    "vae_conv_wgrad_tma": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "vae_set_umma_wgrad": [_i],
    "vae_umma_wgrad_supported": [_i, _i],
    "vae_conv_wgrad_umma": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "vae_bn_finalize": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _ll, _f, _f, _i, _vp],
    "vae_bn_finalize_apply_lrelu": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _ll, _f, _f, _i, _f, _i, _vp],
    "vae_bn_apply_lrelu": [_vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _i, _vp],
    "vae_bn_lrelu_bwd_reduce": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _i, _vp],
    "vae_bn_lrelu_bwd_apply": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _i, _i, _vp],
    "vae_bn_lrelu_bwd_fused": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _i, _i, _vp],
    "vae_final_block_supported": [_i, _i, _i, _i],
    "vae_final_conv_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp],
    "vae_final_bwd_reduce": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp],
    "vae_final_bwd_apply": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f,
                            _i, _i, _vp],
    "vae_linear_tc_supported": [_i, _i, _i],
    "vae_linear_reduce_tc": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "vae_linear_expand_tc": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "vae_linear_wgrad_tc": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "vae_set_final_tc": [_i],
    "vae_final_tc_supported": [_i, _i, _i, _i, _i],
    "vae_final_conv_fwd_tc": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp],
    "vae_final_bwd_reduce_tc": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp],
    "vae_final_bwd_apply_tc": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f,
                               _i, _vp],
    "vae_adamw_step": [_vp, _i, C.c_double, C.c_double, C.c_double, _vp],
    "vae_adamw_step_dev": [_vp, _vp, _i, _vp, _i, C.c_double, _vp],
    "vae_onecycle_step": [_vp, _vp, _vp, _i, _ll, C.c_double, _i, _i, _vp],
    "vae_heads_supported": [_i, _i],
    "vae_skinny_reduce": [_vp, _vp, _vp, _vp, _i, _i, _i, _ll, _ll, _i, _vp],
    "vae_skinny_reduce_pair": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _ll, _ll, _i, _vp],
    "vae_skinny_wgrad": [_vp, _vp, _vp, _vp, _i, _i, _i, _ll, _ll, _i, _vp],
    "vae_skinny_expand": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _ll, _ll, _i, _vp],
    "vae_nhwc_to_nchw": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "vae_nchw_to_nhwc": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "vae_gemm": [_vp, _vp, _vp, _vp, _i, _i, _i, _ll, _ll, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _vp],
    "vae_colsum": [_vp, _vp, _i, _i, _i, _vp],
    "vae_philox_randn": [_vp, _ll, _ull, _ull, _vp],
    "vae_reparam_fwd": [_vp, _vp, _vp, _vp, _vp, _ll, _ull, _ull, _vp, _vp],
    "vae_reparam_bwd": [_vp, _vp, _vp, _vp, _vp, _ll, _i, _vp],
    "vae_elbo_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _f, _i, _i, _vp],
    "vae_elbo_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _f, _i, _i, _vp],
    "vae_elbo_fused": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _f, _i, _i, _vp],
    "vae_elbo_scale_grads": [_vp, _vp, _vp, _vp, _ll, _ll, _i, _vp],
    "vae_sigmoid": [_vp, _vp, _ll, _i, _vp],
    "vae_sigmoid_bwd": [_vp, _vp, _vp, _ll, _i, _i, _vp],
    "vae_roll_transform": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _f, _ull, _ull, _vp],
}
_RESTYPES = {"vae_last_error": C.c_char_p, "vae_umma_packed_elems": C.c_longlong, "vae_umma_tma_packed_elems": C.c_longlong,
             "vae_umma_rw_packed_elems": C.c_longlong}



class AdamWTensor(C.Structure):
    """include/vae_b200.h: vae_adamw_tensor"""
    _fields_ = [("p", _vp), ("g", _vp), ("m", _vp), ("v", _vp), ("step", _vp), ("n", _ll), ("lr", _f), ("wd", _f)]


_lib = None


def load():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or torch-vae_b200/csrc/build.sh). "
            "torch_vae_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.vae_abi_version() != 1:
        raise RuntimeError("libvae_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().vae_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
