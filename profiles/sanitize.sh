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
