"""ctypes binding of libzsv_b200.so (the C ABI declared in include/zsv_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError carrying
``zsv_last_error()`` is raised.  PyTorch only provides device memory and streams here.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libzsv_b200.so"

ABI_VERSION = 8
X_NDHWC = 0
X_WFOLD = 1


class ConvDesc(C.Structure):
    """Mirror of ``zsv_conv_desc`` (include/zsv_b200.h)."""

    _fields_ = [(n, C.c_int32) for n in (
        "N", "T", "H", "W", "Cin", "Cout", "kt", "kh", "kw", "st", "sh", "sw", "pt", "ph", "pw", "x_layout")]

    def key(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


class BnBwdFuse(C.Structure):
    """Mirror of ``zsv_bn_bwd_fuse`` (include/zsv_b200.h)."""

    _fields_ = [("y", C.c_void_p), ("table", C.c_void_p), ("relu", C.c_int32), ("partial", C.c_void_p),
                ("partial_rows", C.c_int32), ("rows_written", C.c_int32)]


class AdamHyper(C.Structure):
    """Mirror of ``zsv_adam_hyper`` (include/zsv_b200.h)."""

    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("grad_scale", C.c_float), ("lr_dev", C.c_void_p)]


class BnFold(C.Structure):
    """Mirror of ``zsv_bn_fold`` (include/zsv_b200.h)."""

    _fields_ = [("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p), ("running_var", C.c_void_p),
                ("bias_out", C.c_void_p), ("eps", C.c_float)]


_P = C.c_void_p
_I = C.c_int
_LL = C.c_longlong
_F = C.c_float
_SZ = C.c_size_t
_DP = C.POINTER(ConvDesc)

# name -> (restype, argtypes); every symbol of include/zsv_b200.h is listed (tests check the export table)
SIGNATURES = {
    "zsv_last_error": (C.c_char_p, []),
    "zsv_abi_version": (_I, []),
    "zsv_cpad": (_I, [_I]),
    "zsv_launch_count": (C.c_ulonglong, []),
    "zsv_sm_count": (_I, []),
    "zsv_conv3d_out_shape": (_I, [_DP, C.POINTER(C.c_int32)]),
    "zsv_conv3d_packed_weight_bytes": (_SZ, [_DP, _I]),
    "zsv_conv3d_pack_weight": (_I, [_DP, _P, _P, _P, _P]),
    "zsv_conv3d_pack_weights": (_I, [_I, _P, _P, _P, _P, _P]),
    "zsv_conv3d_stat_rows": (_I, [_DP]),
    "zsv_conv3d_fprop": (_I, [_DP, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "zsv_conv3d_pack_weights_folded": (_I, [_I, _P, _P, _P, _P, _P]),
    "zsv_conv3d_dgrad": (_I, [_DP, _P, _P, _P, _P, C.POINTER(BnBwdFuse), _P]),
    "zsv_conv3d_wgrad_workspace": (_SZ, [_DP]),
    "zsv_conv3d_wgrad": (_I, [_DP, _P, _P, _P, _P, _SZ, _P]),
    "zsv_bias_grad_workspace": (_SZ, [_I]),
    "zsv_bias_grad": (_I, [_P, _P, _LL, _I, _P, _SZ, _P]),
    "zsv_relu_bwd": (_I, [_P, _P, _P, _LL, _I, _P, _P, _SZ, _P]),
    "zsv_linear_workspace": (_SZ, [_I, _I, _I]),
    "zsv_linear_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _SZ, _P]),
    "zsv_linear_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _SZ, _P]),
    "zsv_l2norm_fwd": (_I, [_P, _P, _P, _I, _I, _F, _P]),
    "zsv_l2norm_bwd": (_I, [_P, _P, _P, _P, _I, _I, _F, _P]),
    "zsv_repack_input": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "zsv_clip_transform": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "zsv_ndhwc_to_ncdhw": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "zsv_ncdhw_to_ndhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "zsv_bn_finalize_workspace": (_SZ, [_I]),
    "zsv_bn_finalize": (_I, [_P, _P, _I, _I, _LL, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "zsv_bn_bwd_finish": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _LL, _I, _P, _SZ, _P]),
    "zsv_bn_eval_scale_shift": (_I, [_I, _P, _P, _P, _P, _F, _P, _P, _P]),
    "zsv_bn_apply": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _LL, _I, _I, _P]),
    "zsv_bn_bwd_workspace": (_SZ, [_I]),
    "zsv_bn_bwd": (_I, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _LL, _I, _P, _SZ,
                        _P]),
    "zsv_head_fwd": (_I, [_P, _I, _I, _I, _P, _P, _I, _P, _P, _I, _F, _P, _P, _P, _P, _P]),
    "zsv_head_bwd_scratch": (_SZ, [_I, _I, _I, _I]),
    "zsv_head_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _I, _P, _I, _F, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "zsv_mse_fwd_bwd": (_I, [_P, _P, _I, _I, _F, _P, _P, _P]),
    "zsv_adam_step": (_I, [_I, _P, _P, _P, _P, _P, _P, C.POINTER(AdamHyper), _P]),
    "zsv_adam_pack_step": (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(AdamHyper), _P]),
    "zsv_nearest_class": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "zsv_maxpool3d_fwd": (_I, [_P, _P, _P] + [_I] * 11 + [_P]),
    "zsv_maxpool3d_bwd": (_I, [_P, _P, _P, _P] + [_I] * 11 + [_P, _P, _SZ, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ["ZSV_LIB_PATH"]) if os.environ.get("ZSV_LIB_PATH") else LIB_PATH   # A/B of two builds
    if not path.exists():
        raise RuntimeError(
            f"{path} is missing: build it with `python -m zeroshotvideoclassification_b200.build` "
            "(or __graft_entry__.build()). There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(os.fspath(path))
    lib.zsv_abi_version.restype = C.c_int
    if lib.zsv_abi_version() != ABI_VERSION:       # first: a stale build may lack newer symbols altogether
        raise RuntimeError(f"libzsv_b200.so ABI {lib.zsv_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().zsv_last_error().decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    if status != 0:
        raise RuntimeError(f"{what} failed with zsv_status {status}: {last_error()}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None passes NULL)."""
    return None if t is None else t.data_ptr()


def launch_count() -> int:
    """Kernels launched by the library in this process so far."""
    return int(load().zsv_launch_count())


def cpad(c: int) -> int:
    return (c + 7) & ~7
