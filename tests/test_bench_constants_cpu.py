"""bench.py's per-workload constants against the oracle's counting build, and the work split of the strong-scaling run."""
import importlib.util
import os

import pytest

from conftest import ROOT, config_case


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", ["c1_example", "c2_case3", "c3_case7", "c3_case7_evolving", "c4_trappist1", "c5_circumbinary"])
def test_flops_per_system_step_table_reproduces(name):
    """WORKLOADS[...][2] (the roofline numerator) = exact operation count of the reference's arithmetic as written, counted by
    the oracle's counting build over the first 100 steps of the unperturbed case (SURVEY §8d counting rule), within 0.1 %."""
    from oracle.binding import count_flops
    from posidonius_b200.case import case_from_dict
    bench = _bench()
    case, tables = case_from_dict(config_case(name))
    counted = count_flops(case, tables, 100)["flops_per_step"]
    assert abs(counted - bench.WORKLOADS[name][2]) <= 1e-3 * counted, (name, counted, bench.WORKLOADS[name][2])


def test_strong_scaling_split_covers_the_ensemble_once():
    """bench.py --gpus N: 65536 systems in total, contiguous ranges, every member on exactly one rank."""
    from posidonius_b200.shard import shard_range
    for world in (1, 2, 4, 8, 3, 7):
        ranges = [shard_range(65536, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == 65536
        for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
            assert a1 == b0 and a1 > a0
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
    assert [b - a for a, b in (shard_range(65536, r, 8) for r in range(8))] == [8192] * 8


def test_reference_arm_and_repo_arm_share_the_workload_definition():
    bench = _bench()
    assert bench.DEFAULT_WORKLOAD == "c4_trappist1" and bench.WORKLOADS["c4_trappist1"][1] == 65536
    case, _ = bench.load_case("c4_trappist1", False)
    assert case.n_particles == 8 and case.time_step == 0.08
    # run-length knobs only: physics fields equal the committed configuration
    from posidonius_b200.case import case_from_dict
    ref, _ = case_from_dict(config_case("c4_trappist1"))
    for b in range(8):
        assert case.bodies[b].mass == ref.bodies[b].mass and case.bodies[b].inertial_position[:] == ref.bodies[b].inertial_position[:]
