"""Loads libposidonius_b200.so (the hand-written CUDA library) and declares the C ABI for ctypes.

There is no fallback: if the library is missing or a symbol of include/posidonius_b200.h is not
exported, importing fails loudly.
"""
import ctypes as C
import os

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
# PB200_LIB selects an experimental build variant of the same library (tuning runs only)
LIB_PATH = os.environ.get("PB200_LIB") or os.path.join(HERE, "libposidonius_b200.so")

_SIGNATURES = {
    "pb200_version": (C.c_char_p, []),
    "pb200_last_error": (C.c_char_p, []),
    "pb200_device_count": (C.c_int, []),
    "pb200_case_validate": (C.c_int, [C.POINTER(abi.Case), C.POINTER(abi.Table), C.c_size_t]),
    "pb200_case_load": (C.c_int, [C.c_char_p, C.POINTER(abi.Case), C.POINTER(C.c_void_p)]),
    "pb200_case_save": (C.c_int, [C.c_char_p, C.POINTER(abi.Case), C.POINTER(abi.Table), C.c_size_t]),
    "pb200_table_store_tables": (C.POINTER(abi.Table), [C.c_void_p]),
    "pb200_table_store_count": (C.c_size_t, [C.c_void_p]),
    "pb200_table_store_free": (None, [C.c_void_p]),
    "pb200_ensemble_create": (C.c_int, [C.POINTER(abi.Case), C.c_size_t, C.c_size_t, C.POINTER(abi.Table), C.c_size_t,
                                        C.c_int, C.POINTER(C.c_void_p)]),
    "pb200_ensemble_create_perturbed": (C.c_int, [C.POINTER(abi.Case), C.c_size_t, C.c_uint64, C.c_double, C.POINTER(abi.Table),
                                                  C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "pb200_ensemble_create_perturbed_range": (C.c_int, [C.POINTER(abi.Case), C.c_uint64, C.c_size_t, C.c_uint64, C.c_double,
                                                        C.POINTER(abi.Table), C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "pb200_ensemble_destroy": (None, [C.c_void_p]),
    "pb200_ensemble_n_particles": (C.c_int, [C.c_void_p]),
    "pb200_ensemble_n_systems": (C.c_size_t, [C.c_void_p]),
    "pb200_ensemble_set_time_limit": (C.c_int, [C.c_void_p, C.c_double]),
    "pb200_ensemble_set_snapshot_periods": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "pb200_ensemble_set_arithmetic": (C.c_int, [C.c_void_p, C.c_int]),
    "pb200_ensemble_initialize_physical_values": (C.c_int, [C.c_void_p]),
    "pb200_ensemble_step": (C.c_int, [C.c_void_p, C.c_uint64]),
    "pb200_ensemble_synchronize": (C.c_int, [C.c_void_p]),
    "pb200_ensemble_last_step_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "pb200_ensemble_launch_count": (C.c_uint64, [C.c_void_p]),
    "pb200_ensemble_last_pieces": (C.c_uint, [C.c_void_p]),
    "pb200_ensemble_last_kernel": (C.c_char_p, [C.c_void_p]),
    "pb200_case_step_kernel": (C.c_char_p, [C.POINTER(abi.Case), C.c_size_t, C.c_int, C.c_int]),
    "pb200_ensemble_history_capacity": (C.c_size_t, [C.c_void_p]),
    "pb200_ensemble_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "pb200_ensemble_download": (C.c_int, [C.c_void_p, C.POINTER(abi.StateView)]),
    "pb200_ensemble_upload": (C.c_int, [C.c_void_p, C.POINTER(abi.StateView)]),
    "pb200_ensemble_get_case": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(abi.Case)]),
    "pb200_ensemble_history_pending": (C.c_size_t, [C.c_void_p]),
    "pb200_ensemble_history_drain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "pb200_ensemble_summary": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "pb200_ensemble_run_host": (C.c_int, [C.c_void_p, C.POINTER(abi.StateView), C.c_uint64]),
    "pb200_measure_fp64_peak": (C.c_int, [C.c_int, C.c_double, C.POINTER(C.c_double)]),
}

_LIB = None


class LibraryError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryError(
                "%s is missing: build it with `python -m posidonius_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(L, name)
            except AttributeError:
                raise LibraryError("libposidonius_b200.so does not export %s" % name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def last_error():
    return lib().pb200_last_error().decode("utf-8", "replace")
