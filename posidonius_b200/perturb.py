"""Synthetic perturbed ensembles: member k = base case with every non-host body's
heliocentric position and velocity multiplied component-wise by (1 + delta), delta ~ U(-amp, amp);
masses, radii, spins and parameters unchanged; barycentric coordinates recomputed as
Universe::new does (reference src/particles/universe.rs:95-105, 663-697). Member 0 is the base case.
"""
import ctypes as C

import numpy as np

from . import abi

CASE_DTYPE = np.dtype(abi.Case)


def cases_as_numpy(cases):
    """Structured-array view of a ctypes array of abi.Case (no copy)."""
    return np.frombuffer(cases, dtype=CASE_DTYPE)


def make_ensemble_cases(base, n_systems, seed, amplitude=1e-3):
    cases = (abi.Case * n_systems)()
    C.memmove(cases, (abi.Case * 1)(base), 0)  # no-op, keeps ctypes happy about types
    arr = cases_as_numpy(cases)
    arr[:] = np.frombuffer((abi.Case * 1)(base), dtype=CASE_DTYPE)[0]
    n = base.n_particles
    host = base.host_most_massive
    rng = np.random.default_rng(seed)
    delta = rng.uniform(-amplitude, amplitude, size=(n_systems, n, 6))
    delta[0] = 0.0
    delta[:, host, :] = 0.0
    bodies = arr["bodies"]
    hp = np.array([base.bodies[b].heliocentric_position[:] for b in range(n)])
    hv = np.array([base.bodies[b].heliocentric_velocity[:] for b in range(n)])
    mass = np.array([base.bodies[b].mass for b in range(n)])
    pos = hp[None] * (1.0 + delta[:, :, 0:3])
    vel = hv[None] * (1.0 + delta[:, :, 3:6])
    # calculate_center_of_mass (universe.rs:663-697): running pairwise centre of mass in body order
    cp = np.zeros((n_systems, 3))
    cv = np.zeros((n_systems, 3))
    cm = 0.0
    for b in range(n):
        cp = cp * cm + pos[:, b] * mass[b]
        cv = cv * cm + vel[:, b] * mass[b]
        new = cm + mass[b]
        if new > 0.0:
            cp = cp / new
            cv = cv / new
        cm = new
    bodies["heliocentric_position"][:, :n, :] = pos
    bodies["heliocentric_velocity"][:, :n, :] = vel
    bodies["inertial_position"][:, :n, :] = pos - cp[:, None, :]
    bodies["inertial_velocity"][:, :n, :] = vel - cv[:, None, :]
    return cases


def _splitmix64(state):
    """One step of SplitMix64 on an array of uint64 states (in place); returns the outputs."""
    with np.errstate(over="ignore"):
        state += np.uint64(0x9e3779b97f4a7c15)
        z = state.copy()
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xbf58476d1ce4e5b9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94d049bb133111eb)
        return z ^ (z >> np.uint64(31))


def splitmix_cases(base, n_systems, seed, amplitude=1e-3):
    """Host-side statement of the recipe of pb200_ensemble_create_perturbed / `posidonius-b200 ensemble` (header of
    include/posidonius_b200.h): used by the tests to check the device-built ensemble bit for bit."""
    cases = (abi.Case * n_systems)()
    arr = cases_as_numpy(cases)
    arr[:] = np.frombuffer((abi.Case * 1)(base), dtype=CASE_DTYPE)[0]
    n = base.n_particles
    host = base.host_most_massive
    with np.errstate(over="ignore"):
        state = np.uint64(seed) * np.uint64(0x100000001b3) + np.arange(n_systems, dtype=np.uint64)
    hp = np.tile(np.array([base.bodies[b].heliocentric_position[:] for b in range(n)])[None], (n_systems, 1, 1))
    hv = np.tile(np.array([base.bodies[b].heliocentric_velocity[:] for b in range(n)])[None], (n_systems, 1, 1))
    mass = np.array([base.bodies[b].mass for b in range(n)])
    for b in range(n):
        if b == host:
            continue
        for target in (hp, hv):
            for c in range(3):
                u = (_splitmix64(state) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
                target[:, b, c] = target[:, b, c] * (1.0 + (2.0 * u - 1.0) * amplitude)
    hp[0] = [base.bodies[b].heliocentric_position[:] for b in range(n)]   # member 0 is the base case
    hv[0] = [base.bodies[b].heliocentric_velocity[:] for b in range(n)]
    cp = np.zeros((n_systems, 3))
    cv = np.zeros((n_systems, 3))
    cm = 0.0
    for b in range(n):
        cp = cp * cm + hp[:, b] * mass[b]
        cv = cv * cm + hv[:, b] * mass[b]
        new = cm + mass[b]
        if new > 0.0:
            cp = cp / new
            cv = cv / new
        cm = new
    bodies = arr["bodies"]
    bodies["heliocentric_position"][:, :n, :] = hp
    bodies["heliocentric_velocity"][:, :n, :] = hv
    bodies["inertial_position"][:, :n, :] = hp - cp[:, None, :]
    bodies["inertial_velocity"][:, :n, :] = hv - cv[:, None, :]
    # member 0 keeps the base case's own barycentric coordinates
    arr[0] = np.frombuffer((abi.Case * 1)(base), dtype=CASE_DTYPE)[0]
    return cases
