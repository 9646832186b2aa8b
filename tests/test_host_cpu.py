"""CPU-side tests: the C ABI library loads and exports every declared symbol, the JSON reader rejects what the
hot path does not cover (no fallback), ensemble packing, header/ctypes layout agreement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import CONFIG_NAMES, ROOT, config_case, load_json_gz
from posidonius_b200 import abi
from posidonius_b200.case import InvalidCaseError, UnsupportedCaseError, case_from_dict


def test_library_loads_and_exports_every_header_symbol():
    from posidonius_b200._lib import exported_symbols, lib
    L = lib()
    header = open(os.path.join(ROOT, "include", "posidonius_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)  # strip comments
    declared = set(re.findall(r"^[a-z_0-9 \*]*\b(pb200_[a-z0-9_]+)\s*\(", header, flags=re.M))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), "libposidonius_b200.so does not export %s" % name
    assert declared == set(exported_symbols())
    assert b"sm_100a" in L.pb200_version()


def test_ctypes_layout_matches_header_sizes():
    # sizes implied by the header: doubles/ints only, natural alignment
    assert C.sizeof(abi.Body) == 8 * (5 + 3 * 7 + 9) + 4 * 10 + 8 * 8
    assert C.sizeof(abi.Case) == 8 * 9 + 8 * 3 + 4 * 14 + C.sizeof(abi.Body) * 10 + 8 * 30 * 2 + 8 * 100 * 2
    assert C.sizeof(abi.StateView) == 8 * 11


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_config_cases_parse_and_validate_without_gpu(name):
    from posidonius_b200.ensemble import validate_case
    case, tables = case_from_dict(config_case(name))
    validate_case(case, tables)  # host-only check, no device needed
    assert case.n_particles >= 2 and case.coordinates_type in (0, 1, 2)


def test_rejects_out_of_scope_cases(manifest):
    for name, fx in manifest["reject"].items():
        with pytest.raises(UnsupportedCaseError):
            case_from_dict(load_json_gz(fx["case"]))


def test_rejects_kaula_and_creep_models():
    d = config_case("c4_trappist1")
    model = d["universe"]["particles"][1]["tides"]["effect"]["OrbitingBody"]
    d["universe"]["particles"][1]["tides"]["effect"]["OrbitingBody"] = {"Kaula": model["ConstantTimeLag"]}
    with pytest.raises(UnsupportedCaseError):
        case_from_dict(d)
    d = config_case("c4_trappist1")
    model = d["universe"]["particles"][1]["rotational_flattening"]["effect"]["OrbitingBody"]
    d["universe"]["particles"][1]["rotational_flattening"]["effect"]["OrbitingBody"] = {"CreepCoplanar": model["OblateSpheroid"]}
    with pytest.raises(UnsupportedCaseError):
        case_from_dict(d)


def test_validate_rejects_bad_structure():
    from posidonius_b200.ensemble import validate_case
    case, tables = case_from_dict(config_case("c1_example"))
    case.consider_disk = 1
    with pytest.raises(UnsupportedCaseError):
        validate_case(case, tables)
    case, tables = case_from_dict(config_case("c1_example"))
    case.bodies[0].moment_of_inertia = 0.0
    with pytest.raises(InvalidCaseError):
        validate_case(case, tables)
    case, tables = case_from_dict(config_case("c1_example"))
    case.host_tides = 1
    with pytest.raises(UnsupportedCaseError):
        validate_case(case, tables)


def test_perturbed_ensemble_member_zero_is_the_base_case():
    from posidonius_b200.perturb import make_ensemble_cases
    case, _ = case_from_dict(config_case("c4_trappist1"))
    cases = make_ensemble_cases(case, 128, 20261017 + 4)
    for b in range(case.n_particles):
        assert cases[0].bodies[b].inertial_position[:] == case.bodies[b].inertial_position[:]
        assert cases[0].bodies[b].inertial_velocity[:] == case.bodies[b].inertial_velocity[:]
    p0 = np.array(case.bodies[3].heliocentric_position[:])
    p9 = np.array(cases[9].bodies[3].heliocentric_position[:])
    assert 0 < np.max(np.abs(p9 / p0 - 1.0)) <= 1e-3
    # barycentric: total momentum vanishes
    mom = sum(cases[9].bodies[b].mass * np.array(cases[9].bodies[b].inertial_velocity[:]) for b in range(case.n_particles))
    assert np.max(np.abs(mom)) < 1e-18


def test_no_device_reports_error_not_fallback():
    """Without a GPU the ensemble constructor must fail loudly (PB200_E_CUDA), never integrate on the CPU."""
    from posidonius_b200._lib import lib
    if lib().pb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from posidonius_b200.ensemble import Ensemble, EnsembleError
    case, tables = case_from_dict(config_case("c1_example"))
    with pytest.raises(EnsembleError):
        Ensemble(case, tables)


def test_validate_rejects_out_of_range_and_duplicate_particle_ids():
    """A malformed recovery image must not index the flattened pair-dependent-factor map out of bounds (pb200_ensemble_create)."""
    from posidonius_b200.ensemble import validate_case
    case, tables = case_from_dict(config_case("c4_trappist1"))
    case.bodies[3].id = 10
    with pytest.raises(InvalidCaseError):
        validate_case(case, tables)
    case.bodies[3].id = -1
    with pytest.raises(InvalidCaseError):
        validate_case(case, tables)
    case.bodies[3].id = 2
    with pytest.raises(InvalidCaseError):
        validate_case(case, tables)
    case.bodies[3].id = 3
    validate_case(case, tables)


def test_step_kernel_selection_is_a_pure_function_of_the_case():
    """Which build of the step kernel integrates a case (pb200_case_step_kernel, no device needed): every BASELINE configuration
    at its BASELINE ensemble size lands on its own compile-time build; other effect subsets of 2- / 3-body systems on the
    catch-all lane = planet builds; everything else (other body counts, WHDS, host not at index 0, Anderson / Newhall GR, pure
    gravity) on the run-time-geometry kernel."""
    from posidonius_b200.ensemble import step_kernel_for

    def kernel(name, n_sys, edit=None, **kw):
        d = config_case(name)
        if edit:
            edit(d)
        case, _ = case_from_dict(d)
        return step_kernel_for(case, n_sys, **kw)

    assert kernel("c1_example", 65536) == "s2"
    assert kernel("c2_case3", 4096) == "s2t"
    assert kernel("c3_case7", 16384) == "s3"
    assert kernel("c3_case7_evolving", 16384) == "s3e"
    assert kernel("c4_trappist1", 65536) == "n8w"
    assert kernel("c4_trappist1", 8192) == "n8w"                                  # the 8-GPU shard: 171 CTAs of 384 threads, time-sliced
    assert kernel("c4_trappist1", 4096) == "n8"                                   # fewer than one 384-thread CTA per SM: 64-thread CTAs
    assert kernel("c4_trappist1", 65536, arithmetic=abi.ARITH_STRICT) == "n8"
    assert kernel("c5_circumbinary", 65536) == "s3p"                              # passive planet, one thread per system
    assert kernel("c5_circumbinary", 8192) == "s3j"                               # small shard: two lanes per system
    off = lambda *effects: (lambda d: [d["universe"]["consider_effects"].__setitem__(e, False) for e in effects])
    assert kernel("c1_example", 1000, off("general_relativity")) == "s2any"
    assert kernel("c3_case7_evolving", 1000, off("rotational_flattening")) == "s3any"
    assert kernel("c5_circumbinary", 65536, off("general_relativity")) == "s3jany"
    assert kernel("c1_example", 1000, off("tides", "rotational_flattening", "general_relativity")) == "generic"   # pure gravity: no spin to integrate

    def whds(d):
        d["alternative_coordinates_type"] = "WHDS"
    assert kernel("c1_example", 1000, whds) == "generic"

    def anderson(d):
        d["universe"]["general_relativity_implementation"] = "Anderson1975"
    assert kernel("c3_case7", 1000, anderson) == "generic"
    # the 5-body golden fixtures of the reference: run-time geometry
    import json
    with open(os.path.join(ROOT, "tests", "golden", "manifest.json")) as f:
        fx = json.load(f)["fixtures"]["test_integrator-whfast_jacobi"]
    case, _ = case_from_dict(load_json_gz(fx["case"]))
    assert case.n_particles == 5 and step_kernel_for(case, 1000) == "generic"
