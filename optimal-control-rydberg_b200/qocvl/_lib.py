"""ctypes binding of libqocvl.so (include/qocvl.h).  There is NO fallback: if the CUDA
library is missing the import fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QOCVL_LIB", os.path.join(os.path.dirname(_HERE), "libqocvl.so"))


class Desc(C.Structure):
    """qocvl_desc (include/qocvl.h)."""
    _fields_ = [("model", C.c_int), ("n", C.c_int), ("n_gen", C.c_int), ("n_slices", C.c_int),
                ("n_basis_pts", C.c_int), ("pad", C.c_int), ("n_gauss", C.c_int), ("n_chan", C.c_int),
                ("n_vec", C.c_int), ("adiab_index", C.c_int), ("dt", C.c_double), ("duration", C.c_double),
                ("w_infid", C.c_double), ("w_adiab", C.c_double), ("fid_norm", C.c_double),
                ("eps", C.c_double), ("scale", C.c_double * 4)]


MODEL_STATE_TRANSFER, MODEL_TWO_QUBIT, MODEL_CHAIN = 0, 1, 2
OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_NORM, ERR_UNSUPPORTED = range(6)

# name -> (restype, argtypes); every symbol include/qocvl.h declares
_P, _D, _V = C.c_void_p, C.c_void_p, C.c_void_p   # plan*, device/host double*, stream
SIGNATURES = {
    "qocvl_plan_create": (C.c_int, [C.POINTER(_P), C.POINTER(Desc), C.c_int]),
    "qocvl_plan_destroy": (C.c_int, [_P]),
    "qocvl_set_generators": (C.c_int, [_P, _D, _D, _D]),
    "qocvl_set_filter": (C.c_int, [_P, _D]),
    "qocvl_set_cost": (C.c_int, [_P, _D, _D]),
    "qocvl_set_noise": (C.c_int, [_P, C.c_int, _D, _D]),
    "qocvl_set_symmetry": (C.c_int, [_P, _D]),
    "qocvl_workspace_bytes": (C.c_size_t, [_P]),
    "qocvl_physical_amplitudes": (C.c_int, [_P, _D, _D, _V]),
    "qocvl_physical_amplitudes_vjp": (C.c_int, [_P, _D, _D, _D, _V]),
    "qocvl_auxiliary_amplitudes": (C.c_int, [_P, _D, _D, _V]),
    "qocvl_exponentials": (C.c_int, [_P, _D, _D, _V]),
    "qocvl_propagate": (C.c_int, [_P, _D, _D, _V]),
    "qocvl_cost": (C.c_int, [_P, _D, _D, _V]),
    "qocvl_cost_grad": (C.c_int, [_P, _D, _D, _V]),
    "qocvl_cost_grad_from_wave": (C.c_int, [_P, _D, _D, _D, _V]),
    "qocvl_set_shard": (C.c_int, [_P, C.c_int]),
    "qocvl_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "qocvl_set_precision": (C.c_int, [_P, C.c_int]),
    "qocvl_precision": (C.c_int, [_P]),
    "qocvl_check": (C.c_int, [_P]),
    "qocvl_last_taylor_degree": (C.c_int, [_P]),
    "qocvl_vectors_propagated": (C.c_int, [_P]),
    "qocvl_reduced_rows": (C.c_int, [_P]),
    "qocvl_hbm_bytes_per_call": (C.c_double, [_P, C.c_int]),
    "qocvl_chain_info": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "qocvl_launch_count": (C.c_longlong, [_P]),
    "qocvl_flops_per_call": (C.c_double, [_P, C.c_int]),
    "qocvl_set_timing": (C.c_int, [_P, C.c_int]),
    "qocvl_last_chain_ms": (C.c_double, [_P]),
    "qocvl_last_stage_ms": (C.c_double, [_P, C.c_int]),
    "qocvl_slice_chunks": (C.c_int, [_P]),
    "qocvl_role_degrees": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "qocvl_measure_fp64_peak": (C.c_double, [C.c_int, C.c_int]),
    "qocvl_measure_fp64_peak_reg": (C.c_double, [C.c_int, C.c_int]),
    "qocvl_lindblad_propagate": (C.c_int, [C.c_int] * 7 + [_D] * 6 + [C.c_double, _D, _D]),
    "qocvl_strerror": (C.c_char_p, [C.c_int]),
    "qocvl_last_cuda_error": (C.c_char_p, []),
    "qocvl_version": (C.c_int, []),
}

_lib = None


def load():
    """dlopen libqocvl.so and bind every declared symbol.  Raises ImportError when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"qocvl: CUDA library not found at {LIB_PATH}. Build it with "
            "optimal-control-rydberg_b200/build.sh (or __graft_entry__.build()); there is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class QocvlError(RuntimeError):
    def __init__(self, status, where):
        lib = load()
        msg = lib.qocvl_strerror(status).decode()
        if status == ERR_CUDA:
            msg += " -- " + lib.qocvl_last_cuda_error().decode()
        super().__init__(f"{where}: {msg} (status {status})")
        self.status = status


def check(status, where):
    if status != OK:
        raise QocvlError(status, where)
