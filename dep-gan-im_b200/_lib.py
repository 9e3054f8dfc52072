"""ctypes binding of libdepgan_b200.so (the C ABI declared in include/depgan_b200.h).

This is the only place Python touches the native library.  There is no CPU fallback: if the shared library
cannot be built/loaded, or a call returns an error code, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

HERE = Path(__file__).resolve().parent
HEADER = HERE.parent / "include" / "depgan_b200.h"

MODEL_GEN, MODEL_CRITIC = 0, 1
PREC_FP32, PREC_BF16, PREC_F16, PREC_F16X3 = 0, 1, 2, 3


class Cfg(C.Structure):
    _fields_ = [("H", C.c_int), ("W", C.c_int), ("nicg", C.c_int), ("nc_out", C.c_int), ("noise_len", C.c_int),
                ("max_batch", C.c_int), ("precision", C.c_int), ("training", C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [("in0", C.c_void_p), ("in1", C.c_void_p), ("C0", C.c_int), ("C1", C.c_int),
                ("w_f32", C.c_void_p), ("w_bf16", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("out", C.c_void_p), ("out_pre", C.c_void_p),
                ("film_g", C.c_void_p), ("film_b", C.c_void_p), ("film_stride", C.c_int), ("res", C.c_void_p),
                ("add_src", C.c_void_p), ("mask_src", C.c_void_p), ("relu", C.c_int), ("deconv", C.c_int),
                ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("head_out", C.c_void_p), ("head_nc", C.c_int),
                ("head_act", C.c_int),
                ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cout", C.c_int), ("ks", C.c_int),
                ("in_bf16", C.c_int), ("out_bf16", C.c_int), ("use_tc", C.c_int), ("pool_out", C.c_void_p)]


_P, _I, _LL, _F, _D = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double
_CFGP = C.POINTER(Cfg)

# name -> (restype, argtypes); must list every function include/depgan_b200.h declares (tests check this)
SIGNATURES = {
    "depgan_last_error": (C.c_char_p, []),
    "depgan_abi_version": (_I, []),
    "depgan_manifest_count": (_I, [_I, _CFGP]),
    "depgan_manifest_entry": (_I, [_I, _CFGP, _I, C.c_char_p, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_LL),
                                   C.POINTER(_I)]),
    "depgan_manifest_floats": (_LL, [_I, _CFGP]),
    "depgan_workspace_bytes": (_LL, [_I, _CFGP]),
    "depgan_net_create": (_P, [_I, _CFGP, _P, _P, _P, _LL]),
    "depgan_net_destroy": (None, [_P]),
    "depgan_net_prepare": (_I, [_P, _P]),
    "depgan_gen_forward": (_I, [_P, _P, _P, _P, _I, _P]),
    "depgan_critic_forward": (_I, [_P, _P, _P, _I, _P]),
    "depgan_critic_grads": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _P]),
    "depgan_critic_grads_dem": (_I, [_P, _I, _I, _P, _P, _P, _P, _P, _I, _I, _P]),
    "depgan_gen_eval": (_I, [_P, _P, _P, _P, _P, _P, _F, _P, _P, _I, _I, _P]),
    "depgan_gen_grads": (_I, [_P, _P, _P, _P, _P, _P, _F, _P, _P, _I, _I, _P]),
    "depgan_gen_loss_finalize": (_I, [_P, _P, _P]),
    "depgan_gen_eval_multi": (_I, [_P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _I, _I, _I, _P]),
    "depgan_gen_loss_finalize_multi": (_I, [_P, _P, _I, _I, _LL, _P]),
    "depgan_uresnet_grads": (_I, [_P, _P, _P, _P, _P, _P, _I, _P]),
    "depgan_cce_loss": (_I, [_P, _P, _P, _P, _LL, _I, _F, _P]),
    "depgan_adam_step": (_I, [_P, _P, _P, _P, _LL, _I, _F, _F, _F, _F, _F, _P]),
    "depgan_dem_accumulate": (_I, [_P, _P, _P, _LL, _I, _P]),
    "depgan_dem_postproc": (_I, [_P, _I, _P, _D, _P, _D, _P, _P, _P, _P, _LL, _P]),
    "depgan_uresnet_labels": (_I, [_P, _D, _I, _P, _P, _P, _LL, _P]),
    "depgan_label_confusion": (_I, [_P, _P, _LL, _P, _P]),
    "depgan_set_sync_hook": (_I, [_P, _P, _I]),
    "depgan_nccl_load": (_I, [C.c_char_p]),
    "depgan_nccl_unique_id": (_I, [_P]),
    "depgan_nccl_init": (_I, [C.POINTER(_P), _I, _I, _P]),
    "depgan_nccl_destroy": (_I, [_P]),
    "depgan_allreduce_attach": (_I, [_P, _P, _I]),
    "depgan_peer_create": (_P, [_LL, _I, _I]),
    "depgan_peer_handle": (_I, [_P, _P]),
    "depgan_peer_connect": (_I, [_P, _P]),
    "depgan_peer_destroy": (None, [_P]),
    "depgan_peer_attach": (_I, [_P, _P]),
    "depgan_dp_update": (_I, [_P, _P, _P, _I, _F, _F, _F, _F, _P, _I, _P]),
    "depgan_dp_allreduce_grads": (_I, [_P, _P, _I, _P]),
    "depgan_dp_allreduce_f64": (_I, [_P, _P, _I, _P]),
    "depgan_launch_count": (_LL, []),
    "depgan_profile_begin": (_I, []),
    "depgan_profile_end": (_I, [_P, _P, _P, _P, _I]),
    "depgan_debug_activation": (_I, [_P, C.c_char_p, _P, _LL, C.POINTER(_LL), _I, _P]),
    "depgan_op_conv2d": (_I, [C.POINTER(ConvDesc), _P]),
    "depgan_op_conv_plan": (_I, [C.POINTER(ConvDesc), _P]),
    "depgan_op_wgrad": (_I, [_P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "depgan_op_wgrad_csum": (_I, [_P, _P, _I, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "depgan_op_pack_weights": (_I, [_P, _P, _I, _I, _I, _P]),
    "depgan_op_f32_to_bf16": (_I, [_P, _P, _LL, _P]),
    "depgan_op_f32_to_f16": (_I, [_P, _P, _LL, _P]),
    "depgan_op_bf16_to_f32": (_I, [_P, _P, _LL, _P]),
}


def header_functions():
    """Names of all functions declared in include/depgan_b200.h."""
    txt = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(depgan_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def lib():
    """Loads (building first if needed) the native library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    import os
    # A/B benchmarking hook: an explicitly named build of this same library (never a fallback implementation)
    path = os.environ.get("DEPGAN_B200_LIB") or _build.build()
    try:
        L = C.CDLL(str(path))
    except OSError as e:  # pragma: no cover
        raise RuntimeError("depgan_b200: cannot load %s (%s); there is no CPU fallback" % (path, e))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("depgan_b200 %s failed (%d): %s" % (what, rc, lib().depgan_last_error().decode()))


def last_error():
    return lib().depgan_last_error().decode()
