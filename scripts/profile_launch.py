#!/usr/bin/env python
"""One warm-up launch and one (profiled) launch of the step kernel on a device-built ensemble — the target of the ncu runs:

    ncu --set full --import-source on --clock-control none -k regex:whfast_steps -s 1 -c 1 -o gpurun_out/step \
        python scripts/profile_launch.py --workload c4_trappist1 --systems 65536 --steps 100 --arithmetic hybrid
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4_trappist1")
    ap.add_argument("--systems", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--launches", type=int, default=2)
    ap.add_argument("--arithmetic", default="hybrid")
    args = ap.parse_args()
    import bench
    from posidonius_b200.ensemble import Ensemble
    idx = bench.WORKLOADS[args.workload][0]
    case, tables = bench.load_case(args.workload, False)
    with Ensemble.perturbed(case, tables, args.systems, bench.SEED + idx, bench.AMPLITUDE, arithmetic=bench.ARITH[args.arithmetic]) as ens:
        ens.initialize_physical_values()
        for _ in range(args.launches):
            ens.iterate(args.steps, synchronize=True)
            print("launch: %.3f ms, %d time slice(s), %.4g system-steps/s" % (ens.last_step_ms(), ens.last_pieces(),
                                                                              args.systems * args.steps / (ens.last_step_ms() * 1e-3)))


if __name__ == "__main__":
    main()
