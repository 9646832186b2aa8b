# A/B runs of experimental builds of the library (PB200_LIB): throughput of the bench workload per arithmetic mode, one line per build.
# usage: bash scripts/sweep_variants.sh "hybrid fast" v384 v192 ...   (the shipped build is measured last)
modes="$1"; shift
run() { for m in $modes; do python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-other-workloads --arithmetic $m 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   $m', '%.4g' % d['value'], '%.3f' % d['roofline']['frac'], 'slices', d['config']['time_slices_per_launch'], 'alive', d['config']['systems_alive'], 'dE', '%.3e' % d['config']['ensemble_summary']['max_abs_dE_over_E'])"; done; }
for v in "$@"; do echo "== $v"; PB200_LIB=$PWD/posidonius_b200/libpb200_$v.so run; done
echo "== ship"; run
