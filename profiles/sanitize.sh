#!/bin/bash
# compute-sanitizer evidence for the two hot kernels: memcheck, racecheck (shared-memory hazards:
# both kernels order their shared-memory rings with __syncwarp only) and synccheck, at small shapes (the tools slow a
# kernel down ~100x).  Usage (GPU box): bash profiles/sanitize.sh <tag>  ->  gpurun_out/<tag>_sanitize_*.log
S=${1:-x}
O=gpurun_out
mkdir -p $O
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # name, tool, driver args...
  local name=$1 tool=$2; shift 2
  timeout 900 $CS --tool $tool --print-limit 20 python profiles/prof_driver.py "$@" > $O/${S}_sanitize_${name}_${tool}.log 2>&1
  echo "$name $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/${S}_sanitize_${name}_${tool}.log | tail -1) | $(grep -c 'envs=' $O/${S}_sanitize_${name}_${tool}.log) run(s) completed"
}
for tool in memcheck racecheck synccheck; do
  run window_t16 $tool --envs 300 --K 96 --reps 1 --presteps 50 --window 16
  run window_t32 $tool --envs 130 --K 96 --reps 1 --presteps 50 --window 32
  run step_g1 $tool --envs 700 --K 48 --reps 1 --presteps 50 --window 1 --lanes 1
  run step_g8 $tool --envs 300 --K 48 --reps 1 --presteps 50 --window 1 --lanes 8
  run step_g1_hard $tool --envs 500 --K 32 --reps 1 --presteps 50 --window 1 --lanes 1 --hard
done
# round 2: the host-path kernels (compact_frames_kernel, history_rows_kernel), the fresh-maps regeneration and the policy
# kernel, through their own GPU tests
for tool in memcheck racecheck; do
  timeout 1200 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scengen.py tests/test_gpu_adapters.py -x -q \
      -k "two_engine_split or fresh_maps or policy_kernel" > $O/${S}_sanitize_round2_${tool}.log 2>&1
  echo "round-2 kernels $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/${S}_sanitize_round2_${tool}.log | tail -1) | $(grep -E 'passed|failed' $O/${S}_sanitize_round2_${tool}.log | tail -1)"
done
