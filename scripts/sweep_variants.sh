# A/B runs of experimental builds of the library (PB200_LIB): throughput of the bench workload, one line per build
run() { python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'])"; }
for v in "$@"; do echo "== $v"; PB200_LIB=$PWD/posidonius_b200/libpb200_$v.so run; done
echo "== ship"; run
