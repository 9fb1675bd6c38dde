python -m pytest tests/test_gpu_mlp_core.py tests/test_gpu_render.py tests/test_gpu_train_step.py -x -q 2>&1 | tail -4 > gpurun_out/pytest_ab.log
python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_extras > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
tail -3 gpurun_out/pytest_ab.log; tail -3 gpurun_out/bench_ab.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print(round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']))
print({k:round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()})
PY
