"""ctypes binding of ``libseeme_b200.so`` (the C ABI declared in ``include/seeme_b200.h``).

There is no fallback: if the library is missing or a call fails, a ``RuntimeError`` carrying
``seeme_last_error()`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SEEME_B200_LIB selects another build of the SAME library (e.g. the -DPF_TRACE instrumented one of tools/pf_trace.py)
LIB_PATH = os.environ.get("SEEME_B200_LIB") or os.path.join(_HERE, "lib", "libseeme_b200.so")

_lib = None

c_float_p = C.c_void_p      # device pointers travel as integers
c_handle = C.c_void_p

# name -> (restype, argtypes); must list every symbol of include/seeme_b200.h
SIGNATURES = {
    "seeme_abi_version": (C.c_int, []),
    "seeme_last_error": (C.c_char_p, []),
    "seeme_launch_count": (C.c_ulonglong, []),
    "seeme_prof_enable": (C.c_int, [C.c_int]),
    "seeme_prof_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "seeme_test_umma_linear": (C.c_int, [c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_void_p, C.c_int, c_float_p, c_float_p, C.c_void_p]),
    "seeme_pointnet_create": (C.c_int, [C.POINTER(c_handle), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]),
    "seeme_pointnet_create_ex": (C.c_int, [C.POINTER(c_handle), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]),
    "seeme_pointnet_forward": (C.c_int, [c_handle, c_float_p, C.c_int, C.c_int, c_float_p, c_float_p, C.c_void_p]),
    "seeme_pointnet_destroy": (C.c_int, [c_handle]),
    "seeme_resnet50_create": (C.c_int, [C.POINTER(c_handle), C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "seeme_resnet50_forward": (C.c_int, [c_handle, c_float_p, C.c_int, c_float_p, c_float_p, C.c_void_p]),
    "seeme_resnet50_destroy": (C.c_int, [c_handle]),
    "seeme_vae_create": (C.c_int, [C.POINTER(c_handle), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]),
    "seeme_vae_encode": (C.c_int, [c_handle, c_float_p, C.c_void_p, c_float_p, C.c_int, C.c_int, c_float_p, c_float_p,
                                   c_float_p, C.c_void_p]),
    "seeme_vae_decode": (C.c_int, [c_handle, c_float_p, C.c_void_p, C.c_int, C.c_int, c_float_p, C.c_void_p]),
    "seeme_vae_destroy": (C.c_int, [c_handle]),
    "seeme_denoiser_create": (C.c_int, [C.POINTER(c_handle), C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "seeme_denoiser_set_time_table": (C.c_int, [c_handle, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "seeme_denoiser_forward": (C.c_int, [c_handle, c_float_p, C.c_int, c_float_p, C.c_int, C.c_int, c_float_p, C.c_void_p]),
    "seeme_sampler_run": (C.c_int, [c_handle, c_float_p, c_float_p, C.c_int, C.c_int, C.c_float, C.c_int,
                                    C.POINTER(C.c_int32), C.POINTER(C.c_float), c_float_p, C.c_void_p]),
    "seeme_denoiser_set_backend": (C.c_int, [c_handle, C.c_int]),
    "seeme_denoiser_destroy": (C.c_int, [c_handle]),
    "seeme_ddim_step": (C.c_int, [c_float_p, c_float_p, c_float_p, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float,
                                  C.c_void_p]),
    "seeme_smpl_create": (C.c_int, [C.POINTER(c_handle), c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                    C.POINTER(C.c_int32), C.c_int]),
    "seeme_smpl_forward": (C.c_int, [c_handle, c_float_p, c_float_p, c_float_p, c_float_p, C.c_int, c_float_p, c_float_p,
                                     c_float_p, C.c_void_p]),
    "seeme_smpl_forward_feats": (C.c_int, [c_handle, c_float_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, c_float_p, C.c_int,
                                           C.c_void_p, c_float_p, c_float_p, c_float_p, C.c_void_p]),
    "seeme_smpl_destroy": (C.c_int, [c_handle]),
}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first (python -m seeme_b200.build). "
                "seeme_b200 has no CPU or PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError here = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        if l.seeme_abi_version() != 1:
            raise RuntimeError(f"libseeme_b200 ABI version {l.seeme_abi_version()} != 1")
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().seeme_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with code {rc}: {msg}")


def prof_enable(on: bool) -> None:
    check(lib().seeme_prof_enable(1 if on else 0), "seeme_prof_enable")


def prof_read(prof_id: int):
    ms, n = C.c_double(), C.c_longlong()
    check(lib().seeme_prof_read(prof_id, C.byref(ms), C.byref(n)), "seeme_prof_read")
    return ms.value, n.value


def launch_count() -> int:
    return int(lib().seeme_launch_count())
