#!/bin/bash
# A/B of kernel variants on the GPU box: build alternates with `nvcc ... -D<MACRO> -o variants/lib_<name>.so` (the
# .so files are git-ignored but travel with gpurun), then `bash profiles/ab_variants.sh`; SHIPSIM_LIB selects the library.
for round in 1 2; do
  for f in base variants/lib_*.so; do
    if [ $f = base ]; then unset SHIPSIM_LIB; else export SHIPSIM_LIB=$PWD/$f; fi
    echo "$(basename $f) headline: $(python profiles/prof_driver.py --envs 4096 --K 1000 --reps 20 | cut -c1-70)"
    echo "$(basename $f) hard:     $(python profiles/prof_driver.py --envs 65536 --K 100 --reps 10 --window 1 --hard | cut -c1-70)"
    echo "$(basename $f) 1M:       $(python profiles/prof_driver.py --envs 1048576 --K 32 --reps 10 --window 1 | cut -c1-70)"
  done
done
