set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp_core.py tests/test_gpu_render.py tests/test_gpu_train_step.py -m gpu -q -x -s > gpurun_out/r2_pp_pytest.log 2>&1; echo "new-lib parity rc=$?"; grep -E "passed|failed|FAILED|ERROR|bench-config" gpurun_out/r2_pp_pytest.log | tail -6
for rep in 1 2; do
for v in old noshare b200; do
  FMOV_LIB="$PWD/fmov_pose_b200/libfmov_$v.so" timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_extras 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']; print('$v', round(d['value']), round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items() if b>0.5})" | tee -a gpurun_out/r2_pingpong_ab.txt
done
done
