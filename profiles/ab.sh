# A/B of library builds inside ONE gpurun call (box-to-box and run-to-run noise is ~3%):
#   bash profiles/ab.sh libA.so libB.so [...]   -> alternates the given builds (paths relative to the repo root), three rounds
cd "$(dirname "$0")/.."
for rep in 1 2 3; do
  for v in "$@"; do
    FMOV_LIB=$PWD/$v timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']; print('$v', round(d['value']), round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items() if b>0.3})"
  done
done
