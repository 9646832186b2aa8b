"""ctypes binding of liboracle (CPU restatement of the reference). TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from posidonius_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libpb200_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "oracle_core.hpp")] + [
        os.path.join(_HERE, "..", "include", "posidonius_b200.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        L.pb200_oracle_create.restype = C.c_void_p
        L.pb200_oracle_create.argtypes = [C.POINTER(abi.Case), C.POINTER(abi.Table), C.c_size_t]
        L.pb200_oracle_destroy.argtypes = [C.c_void_p]
        L.pb200_oracle_initialize_physical_values.argtypes = [C.c_void_p]
        L.pb200_oracle_iterate.restype = C.c_uint64
        L.pb200_oracle_iterate.argtypes = [C.c_void_p, C.c_uint64]
        L.pb200_oracle_status.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.pb200_oracle_store.argtypes = [C.c_void_p, C.POINTER(abi.Case)]
        L.pb200_oracle_history_bytes.restype = C.c_size_t
        L.pb200_oracle_history_bytes.argtypes = [C.c_void_p]
        L.pb200_oracle_history_drain.restype = C.c_size_t
        L.pb200_oracle_history_drain.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.pb200_oracle_summary.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.pb200_oracle_last_midpoint_iterations.argtypes = [C.c_void_p]
        L.pb200_oracle_kepler_branches.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.pb200_oracle_additional_effects.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.pb200_oracle_run_ensemble.argtypes = [C.POINTER(abi.Case), C.c_size_t, C.c_size_t, C.POINTER(abi.Table),
                                                C.c_size_t, C.c_uint64, C.c_int, C.c_int, C.POINTER(abi.Case),
                                                C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.pb200_oracle_count_flops.argtypes = [C.POINTER(abi.Case), C.POINTER(abi.Table), C.c_size_t, C.c_uint64,
                                               C.c_int, C.POINTER(C.c_uint64)]
        L.pb200_oracle_version.restype = C.c_char_p
        _LIB = L
    return _LIB


class OracleSystem:
    """One system integrated on the CPU by the restatement of WHFast::iterate."""

    def __init__(self, case, tables):
        self._tables = tables  # keep arrays alive
        self._case = case
        self._h = lib().pb200_oracle_create(C.byref(case), tables.as_ctypes(), len(tables))

    def close(self):
        if self._h:
            lib().pb200_oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def initialize_physical_values(self):
        return lib().pb200_oracle_initialize_physical_values(self._h)

    def iterate(self, n_steps=1):
        return lib().pb200_oracle_iterate(self._h, n_steps)

    def status(self):
        w = C.c_uint32()
        it = C.c_uint64()
        st = lib().pb200_oracle_status(self._h, C.byref(w), C.byref(it))
        return st, w.value, it.value

    def case(self):
        from posidonius_b200.case import copy_case
        out = copy_case(self._case)
        lib().pb200_oracle_store(self._h, C.byref(out))
        return out

    def history(self):
        nb = lib().pb200_oracle_history_bytes(self._h)
        buf = C.create_string_buffer(nb)
        got = lib().pb200_oracle_history_drain(self._h, buf, nb)
        return buf.raw[:got]

    def summary(self):
        e = C.c_double()
        l = C.c_double()
        lib().pb200_oracle_summary(self._h, C.byref(e), C.byref(l))
        return e.value, l.value

    def last_midpoint_iterations(self):
        return lib().pb200_oracle_last_midpoint_iterations(self._h)

    def kepler_branches(self):
        """Calls of kepler_individual_step so far by branch: (newton_converged, quartic, bisection, hyperbolic)."""
        out = (C.c_uint64 * 4)()
        lib().pb200_oracle_kepler_branches(self._h, out)
        return tuple(int(x) for x in out)

    def additional_effects(self):
        n = self._case.n_particles
        acc = np.zeros((n, 3))
        dl = np.zeros((n, 3))
        lib().pb200_oracle_additional_effects(self._h, acc.ctypes.data_as(C.POINTER(C.c_double)),
                                              dl.ctypes.data_as(C.POINTER(C.c_double)))
        return acc, dl


def run_ensemble(cases, n_systems, tables, n_steps, init_physical=True, n_threads=1):
    """cases: abi.Case (replicated) or ctypes array of n_systems cases. Returns (out_cases, status, seconds)."""
    if isinstance(cases, abi.Case):
        arr = (abi.Case * 1)(cases)
        n_cases = 1
    else:
        arr = cases
        n_cases = len(cases)
    out = (abi.Case * n_systems)()
    status = (C.c_int32 * n_systems)()
    secs = C.c_double()
    lib().pb200_oracle_run_ensemble(arr, n_cases, n_systems, tables.as_ctypes(), len(tables), n_steps,
                                    int(init_physical), n_threads, out, status, C.byref(secs))
    return out, np.ctypeslib.as_array(status).copy(), secs.value


def count_flops(case, tables, n_steps=100, init_physical=True):
    counts = (C.c_uint64 * 6)()
    lib().pb200_oracle_count_flops(C.byref(case), tables.as_ctypes(), len(tables), n_steps, int(init_physical), counts)
    add, mul, div, sqrt, _, stumpff = [int(x) for x in counts]
    return {"add": add, "mul": mul, "div": div, "sqrt": sqrt, "total": add + mul + div + sqrt, "steps": n_steps,
            "flops_per_step": (add + mul + div + sqrt) / n_steps, "stumpff_per_step": stumpff / n_steps}
