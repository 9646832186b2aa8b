"""Shared helpers of the GPU parity tests: run the CUDA ensemble and the CPU oracle on the same cases."""
import ctypes as C

import numpy as np

from posidonius_b200 import abi


def body_vectors(case, field):
    n = case.n_particles
    return np.array([getattr(case.bodies[b], field)[:] for b in range(n)])


def rel_err(got, want):
    """max over bodies of |got - want| / |want| with vector norms (the 1e-10 criterion of BASELINE.json)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    num = np.linalg.norm(got - want, axis=-1)
    den = np.linalg.norm(want, axis=-1)
    den = np.where(den > 0, den, 1.0)
    return float(np.max(num / den))


def gpu_state_of(ens):
    """Download the SoA state and return arrays shaped [system, body, 3] / [system, body] / [system]."""
    st = ens.download()
    out = {}
    for k, v in st.items():
        if v.ndim == 3:
            out[k] = np.transpose(v, (2, 1, 0)).copy()
        elif v.ndim == 2:
            out[k] = v.T.copy()
        else:
            out[k] = v.copy()
    return out


def oracle_state_of(cases_out):
    n_sys = len(cases_out)
    n = cases_out[0].n_particles
    out = {k: np.zeros((n_sys, n, 3)) for k in ("position", "velocity", "acceleration", "angular_momentum", "spin",
                                                   "velocity_errors", "angular_momentum_errors")}
    out.update({k: np.zeros((n_sys, n)) for k in ("radius", "radius_of_gyration_2", "moment_of_inertia")})
    out["current_time"] = np.zeros(n_sys)
    for s in range(n_sys):
        c = cases_out[s]
        for b in range(n):
            B = c.bodies[b]
            out["position"][s, b] = B.inertial_position[:]
            out["velocity"][s, b] = B.inertial_velocity[:]
            out["acceleration"][s, b] = B.inertial_acceleration[:]
            out["angular_momentum"][s, b] = B.angular_momentum[:]
            out["spin"][s, b] = B.spin[:]
            out["velocity_errors"][s, b] = c.inertial_velocity_errors[b][:]
            out["angular_momentum_errors"][s, b] = c.particle_angular_momentum_errors[b][:]
            out["radius"][s, b] = B.radius
            out["radius_of_gyration_2"][s, b] = B.radius_of_gyration_2
            out["moment_of_inertia"][s, b] = B.moment_of_inertia
        out["current_time"][s] = c.current_time
    return out
