# A/B of two library builds inside ONE gpurun call (box-to-box and run-to-run noise is ~3%):
#   bash profiles/ab.sh  -> alternates fmov_pose_b200/libfmov_A.so and libfmov_B.so
for v in A B A B A B; do
  FMOV_LIB=/root/repo/fmov_pose_b200/libfmov_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']; print('$v', round(d['value']), round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items() if b>0.5})"
done
