"""Solver branches of the reference that the configuration ensembles never reach, on the GPU against the oracle:
the Kepler bisection fallback (whfast.rs:776-808) after Newton and after the quartic solver, hyperbolic orbits with the
STATE compared, a time step longer than the orbital period (whfast.rs:702-707), an implicit midpoint that exhausts its 10
iterations (whfast.rs:389-391), and a collision that is not a Roche destruction (universe.rs:229-233).

The oracle's branch counters (test instrumentation, oracle.binding.OracleSystem.kepler_branches) prove that each case
really takes the branch; PB200_ARITH_STRICT must then agree bit for bit, the default hybrid arithmetic to 1e-12."""
import numpy as np
import pytest

from conftest import config_case
from parity_util import gpu_state_of, rel_err

pytestmark = pytest.mark.gpu


def _orbit_variant(name, vscale, dt, rscale):
    """Config `name` with body 1's heliocentric velocity scaled by vscale, its distance by rscale, and time step dt."""
    from posidonius_b200.case import case_from_dict, copy_case
    base, tables = case_from_dict(config_case(name))
    c = copy_case(base)
    c.time_step = dt
    c.half_time_step = dt / 2
    h = c.host_most_massive
    for k in range(3):
        rel = c.bodies[1].inertial_position[k] - c.bodies[h].inertial_position[k]
        c.bodies[1].inertial_position[k] = c.bodies[h].inertial_position[k] + rscale * rel
        relv = c.bodies[1].inertial_velocity[k] - c.bodies[h].inertial_velocity[k]
        c.bodies[1].inertial_velocity[k] = c.bodies[h].inertial_velocity[k] + vscale * relv
        c.bodies[1].heliocentric_position[k] = c.bodies[1].inertial_position[k] - c.bodies[h].inertial_position[k]
        c.bodies[1].heliocentric_velocity[k] = c.bodies[1].inertial_velocity[k] - c.bodies[h].inertial_velocity[k]
    return c, tables


def _oracle(case, tables, steps):
    from oracle.binding import OracleSystem
    o = OracleSystem(case, tables)
    o.initialize_physical_values()
    n = o.iterate(steps)
    st, w, it = o.status()
    return o, n, st, w, it


def _compare(E, case, tables, steps, want):
    from posidonius_b200 import abi
    for arithmetic in (abi.ARITH_STRICT, abi.ARITH_HYBRID):
        with E.Ensemble(case, tables, n_systems=5, arithmetic=arithmetic) as ens:
            ens.initialize_physical_values()
            ens.iterate(steps)
            st, w, it = ens.status()
            got = ens.get_case(4)
        for i in range(case.n_particles):
            for key in ("inertial_position", "inertial_velocity", "angular_momentum", "spin"):
                a, b = np.array(getattr(got.bodies[i], key)[:]), np.array(getattr(want.bodies[i], key)[:])
                if arithmetic == abi.ARITH_STRICT:
                    assert np.array_equal(a, b), (i, key, a, b)
                else:
                    assert np.all(np.abs(a - b) <= 1e-12 * max(np.linalg.norm(b), 1e-300)), (i, key, a, b)
        yield st, w, it


@pytest.fixture(scope="module")
def E():
    from posidonius_b200 import ensemble
    return ensemble


# (velocity scale, time step [d], distance scale) of config 2's planet; found by a random search with the oracle's counters
@pytest.mark.parametrize("vscale,dt,rscale,expect", [
    (0.8653681121103338, 7.518985387539499, 1.6452784869868802, "quartic+bisection"),
    (0.2358268472159889, 2.4198183830707607, 6.415270451208836, "newton+quartic+bisection"),
    (0.23183788074137218, 1.5263344898695164, 23.912937538131736, "newton+bisection"),
    (1.2748729284237375, 5.1736931439259575, 9.7843745828414, "hyperbolic+bisection"),
])
def test_kepler_bisection_and_hyperbolic_branches(E, vscale, dt, rscale, expect):
    case, tables = _orbit_variant("c2_case3", vscale, dt, rscale)
    o, n, st, w, it = _oracle(case, tables, 300)
    newton, quartic, bisection, hyperbolic = o.kepler_branches()
    assert n == 300 and st == 0
    assert bisection > 0, "the case must reach the bisection fallback"
    if "quartic" in expect:
        assert quartic > 0
    if "newton" in expect:
        assert newton > 0
    if "hyperbolic" in expect:
        assert hyperbolic == 600   # every drift of the run is on the hyperbola
    else:
        assert hyperbolic == 0
    for gst, gw, git in _compare(E, case, tables, 300, o.case()):
        assert np.all(gst == 0) and np.all(gw == w)


def test_time_step_longer_than_the_period_warns_once_and_matches(E):
    """whfast.rs:702-707: |dt| x invperiod > 1 sets the warning (once); the quartic solver then carries every drift."""
    from posidonius_b200 import abi
    case, tables = _orbit_variant("c2_case3", 1.0, 8.0, 1.0)
    o, n, st, w, it = _oracle(case, tables, 50)
    assert w & abi.WARN_TIMESTEP_GT_PERIOD and o.kepler_branches()[1] > 0
    for gst, gw, git in _compare(E, case, tables, 50, o.case()):
        assert np.all(gst == 0) and np.all(gw == w)
    with E.Ensemble(case, tables, n_systems=2) as ens:
        ens.initialize_physical_values()
        ens.iterate(50)
        assert ens.get_case(0).timestep_warning == 1


def test_midpoint_that_exhausts_its_iterations_warns_and_matches(E):
    """whfast.rs:389-391: a dissipation 1e5 times stronger keeps the implicit midpoint from converging within 10 iterations
    during the first steps; the warning bit is set and the (unconverged) state equals the oracle's."""
    from posidonius_b200 import abi
    from posidonius_b200.case import case_from_dict, copy_case
    base, tables = case_from_dict(config_case("c2_case3"))
    case = copy_case(base)
    for b in range(case.n_particles):
        case.bodies[b].tides_scaled_dissipation_factor *= 1.0e5
    o, n, st, w, it = _oracle(case, tables, 60)
    assert n == 60 and st == 0 and (w & abi.WARN_MIDPOINT_NOT_CONVERGED)
    for gst, gw, git in _compare(E, case, tables, 60, o.case()):
        assert np.all(gst == 0) and np.all(gw & abi.WARN_MIDPOINT_NOT_CONVERGED)


def test_collision_that_is_not_a_roche_destruction(E):
    """universe.rs:224-233 tests the Roche radius first, then the sum of the radii: a host inflated to ten times its radius
    reaches past the planet's orbit while the pair's Roche radius (set by the planet's radius) does not."""
    from posidonius_b200 import abi
    from posidonius_b200.case import case_from_dict, copy_case
    base, tables = case_from_dict(config_case("c2_case3"))
    case = copy_case(base)
    case.bodies[0].radius *= 10.0
    o, n, st, w, it = _oracle(case, tables, 20)
    assert st == abi.STATUS_COLLISION and it == 0
    for arithmetic in (abi.ARITH_STRICT, abi.ARITH_HYBRID, abi.ARITH_FAST):
        with E.Ensemble(case, tables, n_systems=3, arithmetic=arithmetic) as ens:
            ens.initialize_physical_values()
            ens.iterate(20)
            gst, _, git = ens.status()
        assert np.all(gst == abi.STATUS_COLLISION) and np.all(git == 0)


def test_wide_and_narrow_builds_of_the_eight_body_kernel_agree_bit_for_bit(E, monkeypatch):
    """The 8-body kernel exists in 64-thread and 384-thread CTAs (one per SM; chosen by ensemble size): same arithmetic,
    same state — every arithmetic mode, time slicing on (more CTAs than resident slots) and snapshots inside the run."""
    from posidonius_b200 import abi
    from posidonius_b200.case import case_from_dict
    d = config_case("c4_trappist1")
    d["historic_snapshot_period"] = 4.0
    case, tables = case_from_dict(d)
    n_sys = 384 * 148 // 8 + 48 * 200 + 5   # past the switch-over, more CTAs than SMs, the last CTA partly filled
    for arithmetic in (abi.ARITH_HYBRID, abi.ARITH_STRICT, abi.ARITH_FAST):
        out = []
        for flag in ("0", "1"):
            monkeypatch.setenv("PB200_NARROW_BLOCKS", flag)
            with E.Ensemble.perturbed(case, tables, n_sys, 5, 1e-3, arithmetic=arithmetic) as ens:
                ens.initialize_physical_values()
                ens.iterate(120)
                out.append((ens.download(), ens.status(), ens.history_drain()))
        (a, sa, ha), (b, sb, hb) = out
        for k in a:
            assert np.array_equal(a[k], b[k]), (arithmetic, k)
        assert all(np.array_equal(x, y) for x, y in zip(sa, sb)) and np.array_equal(ha, hb) and ha.shape[1] == 3
