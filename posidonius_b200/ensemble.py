"""`Ensemble`: the host-side mirror of the reference's `trait Integrator`
(src/integrator/mod.rs:16-26) for a device-resident ensemble of WHFast systems.

Method names and meaning follow the trait: get_n_particles, get_current_time, get_n_historic_snapshots,
set_time_limit, set_snapshot_periods, initialize_physical_values, iterate. All arithmetic happens in the
CUDA library behind the C ABI (include/posidonius_b200.h); nothing here integrates on the CPU.
"""
import ctypes as C

import numpy as np

from . import abi
from ._lib import lib, last_error
from .case import InvalidCaseError, UnsupportedCaseError


class EnsembleError(RuntimeError):
    pass


def _check(rc):
    if rc == abi.OK:
        return
    msg = last_error()
    if rc == abi.E_UNSUPPORTED:
        raise UnsupportedCaseError(msg)
    if rc == abi.E_INVALID:
        raise InvalidCaseError(msg)
    raise EnsembleError("pb200 error %d: %s" % (rc, msg))


def validate_case(case, tables):
    _check(lib().pb200_case_validate(C.byref(case), tables.as_ctypes(), len(tables)))


class Ensemble:
    def __init__(self, cases, tables, n_systems=None, device=0, arithmetic=None):
        """cases: one abi.Case (replicated n_systems times) or a ctypes array / list of n_systems cases.
        arithmetic: None = the library default (abi.ARITH_HYBRID: fast iterates, exact committed evaluation), abi.ARITH_STRICT
        (every evaluation exact, bit-reproducible) or abi.ARITH_FAST (see the header)."""
        if isinstance(cases, abi.Case):
            arr = (abi.Case * 1)(cases)
            n_cases = 1
            if n_systems is None:
                n_systems = 1
        else:
            n_cases = len(cases)
            arr = cases if isinstance(cases, C.Array) else (abi.Case * n_cases)(*cases)
            if n_systems is None:
                n_systems = n_cases
        self._tables = tables
        self._h = C.c_void_p()
        _check(lib().pb200_ensemble_create(arr, n_cases, n_systems, tables.as_ctypes(), len(tables), device, C.byref(self._h)))
        self.n_systems = n_systems
        self.n_particles = lib().pb200_ensemble_n_particles(self._h)
        self.device = device
        if arithmetic is not None:
            self.set_arithmetic(arithmetic)

    @classmethod
    def perturbed(cls, base, tables, n_systems, seed, amplitude=1e-3, device=0, arithmetic=None, first_member=0):
        """pb200_ensemble_create_perturbed(_range): the synthetic ensemble of SURVEY §8d built on the device (member 0 = base,
        member k = SplitMix64-perturbed copy); the host never holds per-member case images. first_member: this ensemble is
        the shard [first_member, first_member + n_systems) of the global ensemble (one process per GPU)."""
        self = cls.__new__(cls)
        self._tables = tables
        self._h = C.c_void_p()
        _check(lib().pb200_ensemble_create_perturbed_range(C.byref(base), first_member, n_systems, seed, amplitude, tables.as_ctypes(),
                                                           len(tables), device, C.byref(self._h)))
        self.n_systems = n_systems
        self.n_particles = lib().pb200_ensemble_n_particles(self._h)
        self.device = device
        if arithmetic is not None:
            self.set_arithmetic(arithmetic)
        return self

    def set_arithmetic(self, mode):
        _check(lib().pb200_ensemble_set_arithmetic(self._h, mode))

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().pb200_ensemble_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- trait Integrator
    def get_n_particles(self):
        return self.n_particles

    def get_current_time(self):
        return self.download(("current_time",))["current_time"]

    def get_n_historic_snapshots(self, system=0):
        return self.get_case(system).n_historic_snapshots

    def set_time_limit(self, time_limit):
        _check(lib().pb200_ensemble_set_time_limit(self._h, time_limit))

    def set_snapshot_periods(self, historic_snapshot_period, recovery_snapshot_period):
        _check(lib().pb200_ensemble_set_snapshot_periods(self._h, historic_snapshot_period, recovery_snapshot_period))

    def initialize_physical_values(self):
        _check(lib().pb200_ensemble_initialize_physical_values(self._h))

    def iterate(self, n_steps=1, synchronize=True):
        """n_steps calls of Integrator::iterate on every live system. Returns nothing; see status()."""
        _check(lib().pb200_ensemble_step(self._h, n_steps))
        if synchronize:
            self.synchronize()

    # -- ensemble additions
    def synchronize(self):
        _check(lib().pb200_ensemble_synchronize(self._h))

    def last_step_ms(self):
        ms = C.c_float()
        _check(lib().pb200_ensemble_last_step_ms(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return lib().pb200_ensemble_launch_count(self._h)

    def last_pieces(self):
        return lib().pb200_ensemble_last_pieces(self._h)

    def last_kernel(self):
        """Name of the step-kernel build that ran the last iterate() (diagnostics; include/posidonius_b200.h)."""
        return lib().pb200_ensemble_last_kernel(self._h).decode()

    def history_capacity(self):
        return lib().pb200_ensemble_history_capacity(self._h)

    def status(self):
        st = np.zeros(self.n_systems, dtype=np.int32)
        w = np.zeros(self.n_systems, dtype=np.uint32)
        it = np.zeros(self.n_systems, dtype=np.uint64)
        _check(lib().pb200_ensemble_status(self._h, st.ctypes.data_as(C.POINTER(C.c_int32)),
                                           w.ctypes.data_as(C.POINTER(C.c_uint32)), it.ctypes.data_as(C.POINTER(C.c_uint64))))
        return st, w, it

    def _shape(self, field):
        if field in abi.STATE_FIELDS_VEC:
            return (3, self.n_particles, self.n_systems)
        if field in abi.STATE_FIELDS_BODY:
            return (self.n_particles, self.n_systems)
        return (self.n_systems,)

    def make_state_buffers(self, fields=None, pinned=False):
        fields = fields or (abi.STATE_FIELDS_VEC + abi.STATE_FIELDS_BODY + abi.STATE_FIELDS_SYS)
        out = {}
        for f in fields:
            if pinned:
                import torch
                out[f] = torch.zeros(self._shape(f), dtype=torch.float64).pin_memory().numpy()
            else:
                out[f] = np.zeros(self._shape(f), dtype=np.float64)
        return out

    def _view(self, arrays):
        v = abi.StateView()
        for f, a in arrays.items():
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.shape == self._shape(f), f
            setattr(v, f, a.ctypes.data_as(C.POINTER(C.c_double)))
        return v

    def download(self, fields=None, out=None):
        arrays = out if out is not None else self.make_state_buffers(fields)
        v = self._view(arrays)
        _check(lib().pb200_ensemble_download(self._h, C.byref(v)))
        return arrays

    def upload(self, arrays):
        v = self._view(arrays)
        _check(lib().pb200_ensemble_upload(self._h, C.byref(v)))

    def run_host(self, arrays, n_steps):
        """Upload `arrays` (host SoA), advance n_steps, download back into `arrays` — the boundary call timed as e2e."""
        v = self._view(arrays)
        _check(lib().pb200_ensemble_run_host(self._h, C.byref(v), n_steps))

    def get_case(self, system=0):
        out = abi.Case()
        _check(lib().pb200_ensemble_get_case(self._h, system, C.byref(out)))
        return out

    def history_pending(self):
        return lib().pb200_ensemble_history_pending(self._h)

    def history_drain(self, out=None):
        """Returns a uint8 array [n_systems, n_snapshots, n_particles, 156] of reference-layout records.
        out: a (pinned) uint8 buffer of at least that many bytes to reuse; the result is a view of it."""
        n_snap = self.history_pending()
        shape = (self.n_systems, n_snap, self.n_particles, abi.HISTORIC_RECORD_BYTES)
        if out is None:
            buf = np.zeros(shape, dtype=np.uint8)
        else:
            buf = out.reshape(-1)[:int(np.prod(shape))].reshape(shape)
        if n_snap:
            _check(lib().pb200_ensemble_history_drain(self._h, buf.ctypes.data_as(C.c_void_p), buf.nbytes))
        return buf

    def summary(self):
        e = np.zeros(self.n_systems)
        l = np.zeros(self.n_systems)
        _check(lib().pb200_ensemble_summary(self._h, e.ctypes.data_as(C.POINTER(C.c_double)), l.ctypes.data_as(C.POINTER(C.c_double))))
        return e, l


def step_kernel_for(case, n_systems, sm_count=148, arithmetic=abi.ARITH_HYBRID):
    """Name of the step-kernel build that would integrate an ensemble of `n_systems` members of `case` on a GPU with
    `sm_count` SMs (pb200_case_step_kernel; needs no device): "generic", "n8" / "n8w", "s2", "s2t", "s3", "s3e", "s3j", "s3p",
    "s2any", "s3any", "s3jany" — DESIGN.md §3, geometry builds."""
    return lib().pb200_case_step_kernel(C.byref(case), int(n_systems), int(sm_count), int(arithmetic)).decode()


def measure_fp64_peak(device=0, ms_target=50.0):
    v = C.c_double()
    _check(lib().pb200_measure_fp64_peak(device, ms_target, C.byref(v)))
    return v.value
