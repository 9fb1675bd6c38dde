#!/bin/bash
# Round-2 evidence capture on ONE B200 (run through gpurun; every step only after its own command ran without ncu):
#   bash profiles/r2_capture.sh tests     pytest -m gpu (full suite)
#   bash profiles/r2_capture.sh bench     bench.py defaults -> gpurun_out/${TAG}_bench.json
#   bash profiles/r2_capture.sh launches  ncu launch list (gpu__time_duration) of the bench command
#   bash profiles/r2_capture.sh full      ncu --set full of every kernel of one 8192-ray step + marching cubes, summarised
#   bash profiles/r2_capture.sh fullstep  (the step part of `full` only)
# Summaries are written to gpurun_out/ (copied into profiles/ by hand after review); the .ncu-rep files stay in /tmp.
set -u
TAG=${TAG:-r2b}          # r2 = first session of round 2, r2b = final tree of round 2 (CTA-pair engine)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREGEX='regex:fine_|dw_kernel|colsum|composite_|sample_round|sample_coarse|sdf_query|raygen|pose_gf|pose_|adam_step|grad_gather|weight_norm|ray_reduce|loss_fwd|grad_amax|pack_'
for what in "$@"; do
case "$what" in
tests)
  timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log ;;
bench)
  timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; head -c 1500 gpurun_out/${TAG}_bench.json; echo ;;
launches)
  timeout 600 python bench.py --steps 2 --warmup 3 --no_extras --no_cpu_baseline > /dev/null 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no_extras --no_cpu_baseline > gpurun_out/${TAG}_launches_run.log 2>&1
  echo "launches rc=$?"; python profiles/summarize.py launches gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_8192rays.txt 2>&1; head -30 gpurun_out/${TAG}_launches_8192rays.txt ;;
fullstep)
  timeout 300 python profiles/run_kernels.py step 8192 > gpurun_out/${TAG}_step_plain.log 2>&1 && \
  timeout 1500 ncu --set full --clock-control none --import-source on -k "$KREGEX" --launch-skip 40 -c 45 -o /tmp/${TAG}_step8192 -f \
    python profiles/run_kernels.py step 8192 > gpurun_out/${TAG}_ncu_step.log 2>&1
  echo "ncu step rc=$?"; tail -2 gpurun_out/${TAG}_ncu_step.log
  python profiles/summarize.py full /tmp/${TAG}_step8192.ncu-rep > gpurun_out/${TAG}_ncu_full_8192rays.txt 2>&1 ;;
full)
  timeout 300 python profiles/run_kernels.py step 8192 > gpurun_out/${TAG}_step_plain.log 2>&1 && \
  timeout 1500 ncu --set full --clock-control none --import-source on -k "$KREGEX" --launch-skip 40 -c 45 -o /tmp/${TAG}_step8192 -f \
    python profiles/run_kernels.py step 8192 > gpurun_out/${TAG}_ncu_step.log 2>&1
  echo "ncu step rc=$?"; tail -2 gpurun_out/${TAG}_ncu_step.log
  python profiles/summarize.py full /tmp/${TAG}_step8192.ncu-rep > gpurun_out/${TAG}_ncu_full_8192rays.txt 2>&1
  ls -la /tmp/${TAG}_step8192.ncu-rep
  timeout 300 python bench.py --mc_only > gpurun_out/${TAG}_mc_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none -k regex:mc_ -c 9 -o /tmp/${TAG}_mc -f python bench.py --mc_only > gpurun_out/${TAG}_ncu_mc.log 2>&1
  echo "ncu mc rc=$?"; python profiles/summarize.py full /tmp/${TAG}_mc.ncu-rep > gpurun_out/${TAG}_ncu_full_mc.txt 2>&1
  timeout 600 ncu --set full --clock-control none -k regex:sdf_query -c 3 -o /tmp/${TAG}_grid -f python profiles/run_kernels.py grid > gpurun_out/${TAG}_ncu_grid.log 2>&1
  echo "ncu grid rc=$?"; python profiles/summarize.py full /tmp/${TAG}_grid.ncu-rep > gpurun_out/${TAG}_ncu_full_grid.txt 2>&1 ;;
*) echo "unknown step $what"; exit 2 ;;
esac
done
