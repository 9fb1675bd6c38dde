#!/bin/bash
# Round-2 A/B of the fine-stage HBM-traffic experiments that were written at the end of round 1 without GPU time left
# (compiled and spill-checked only: profiles/spills_by_line.py).  Variants of libfmov_b200.so:
#   base      default build
#   relu      -DFMOV_RELU_BITS                       colour backward reads ReLU sign words (16 -> 1 block per tile)
#   rq        -DFMOV_RECOMPUTE_Q                     q_l rebuilt in the backward pass from V-bar, delta, sigma (-32 blocks)
#   relu_rq   both                                   fine_bwd 265 -> 218 blocks per 128-point tile (-18 %)
#   relu_rq1 / relu_rq3   same with FMOV_RQ_PH=1 / 3 (prefetch distance of the three operand streams, in 8-column pieces)
#   l2 / relu_rq_l2       -DFMOV_L2_HINTS: fine_bwd reads H twice; first read evict_last, single-use traffic evict_first
#   l2fwd / all           -DFMOV_L2_HINTS_FWD: fine_fwd writes H1..H7 evict_last (read back by its reverse sweep), delta stores
#                         and the last-use H loads evict_first
#
#   bash profiles/r2_fine_variants.sh build        here (no GPU): builds fmov_pose_b200/libfmov_<name>.so (they travel with gpurun)
#   gpurun --timeout 1500 -- 'bash profiles/r2_fine_variants.sh run'      parity subset per variant, then alternating benches
#   gpurun --timeout 1500 -- 'bash profiles/r2_fine_variants.sh ncu relu_rq'   ncu --set full of the fine/dw kernels + marching cubes
set -u
cd "$(dirname "$0")/.."
declare -A FLAGS=( [base]="" [relu]="-DFMOV_RELU_BITS" [rq]="-DFMOV_RECOMPUTE_Q" [relu_rq]="-DFMOV_RELU_BITS -DFMOV_RECOMPUTE_Q"
                   [relu_rq1]="-DFMOV_RELU_BITS -DFMOV_RECOMPUTE_Q -DFMOV_RQ_PH=1" [relu_rq3]="-DFMOV_RELU_BITS -DFMOV_RECOMPUTE_Q -DFMOV_RQ_PH=3"
                   [l2]="-DFMOV_L2_HINTS" [relu_rq_l2]="-DFMOV_RELU_BITS -DFMOV_RECOMPUTE_Q -DFMOV_L2_HINTS"
                   [l2fwd]="-DFMOV_L2_HINTS_FWD" [all]="-DFMOV_RELU_BITS -DFMOV_RECOMPUTE_Q -DFMOV_L2_HINTS -DFMOV_L2_HINTS_FWD" )
ORDER="${VARIANTS:-base relu rq relu_rq relu_rq1 relu_rq3 l2 relu_rq_l2 l2fwd all}"        # VARIANTS="base relu_rq all" trims the run (~2.5 GPU-min per variant)
case "${1:-}" in
build)
  for v in $ORDER; do
    FMOV_NVCC_EXTRA="${FLAGS[$v]}" FMOV_LIB_OUT="$PWD/fmov_pose_b200/libfmov_$v.so" python -m fmov_pose_b200.build || exit 1
    echo "built libfmov_$v.so  (${FLAGS[$v]})"
  done ;;
run)
  mkdir -p gpurun_out
  for v in $ORDER; do
    # parity first: a variant that is not green is not timed
    if FMOV_LIB="$PWD/fmov_pose_b200/libfmov_$v.so" timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_train_step.py -x -q \
         > "gpurun_out/r2_variant_$v.pytest.log" 2>&1; then echo "$v parity ok"; else echo "$v PARITY FAILED"; tail -15 "gpurun_out/r2_variant_$v.pytest.log"; FLAGS[$v]="FAILED"; fi
  done
  for rep in 1 2; do
    for v in $ORDER; do
      [ "${FLAGS[$v]}" = "FAILED" ] && continue
      FMOV_LIB="$PWD/fmov_pose_b200/libfmov_$v.so" timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_extras 2>/dev/null | tail -1 \
        | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']; print('$v', round(d['value']), round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items() if b>0.5})" \
        | tee -a gpurun_out/r2_variants.txt
    done
  done ;;
ncu)
  # one --set full capture of the step's kernels with the chosen variant (default relu_rq) and of the marching-cubes passes;
  # read back here with `python profiles/summarize.py full gpurun_out/<file>.ncu-rep`
  v="${2:-relu_rq}"
  mkdir -p gpurun_out
  FMOV_LIB="$PWD/fmov_pose_b200/libfmov_$v.so" timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'fine_|dw_kernel|composite_|sample_round' -c 14 -o "gpurun_out/r2_step_$v" -f python profiles/run_kernels.py step 2048 > "gpurun_out/r2_ncu_$v.log" 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mc_' -c 9 -o gpurun_out/r2_mc -f \
    python bench.py --mc_only > gpurun_out/r2_ncu_mc.log 2>&1
  tail -3 "gpurun_out/r2_ncu_$v.log" gpurun_out/r2_ncu_mc.log ;;
*) echo "usage: $0 build|run|ncu [variant]"; exit 2 ;;
esac
