#!/bin/bash
# The measurement sequence behind profiles/r01_<stage>_*: GPU tests, quick timings, the bench line, the reference arm,
# the ncu launch list of a short bench run and one full ncu capture of each kernel.  Usage: bash profiles/final_run.sh <stage>
S=${1:-x}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/${S}_pytest.log; cat $O/${S}_pytest.log
for i in 1 2 3; do python profiles/prof_driver.py --envs 4096 --K 1000 --reps 20 | cut -c1-75; done
python profiles/prof_driver.py --envs 65536 --K 100 --reps 10 --window 1 --hard | cut -c1-75
python profiles/prof_driver.py --envs 1048576 --K 32 --reps 10 --window 1 | cut -c1-75
python bench.py --steps 200 --warmup 5 > $O/${S}_bench_n1.json 2> $O/${S}_bench.err; cut -c1-200 $O/${S}_bench_n1.json
python bench.py --impl reference --steps 10 --warmup 2 > $O/${S}_bench_reference_arm.json 2>> $O/${S}_bench.err; cut -c1-120 $O/${S}_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${S}_launches.csv python bench.py --steps 2 --warmup 1 > $O/${S}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:window_kernel -s 5 -c 1 -f -o $O/${S}_win python profiles/prof_driver.py --envs 4096 --K 1000 --reps 3 > $O/${S}_ncu_win.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 1 -f -o $O/${S}_step python profiles/prof_driver.py --envs 1048576 --K 32 --reps 3 --window 1 > $O/${S}_ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 1 -f -o $O/${S}_hard python profiles/prof_driver.py --envs 65536 --K 100 --reps 3 --window 1 --hard > $O/${S}_ncu_hard.log 2>&1
ls -la $O | tail -12
