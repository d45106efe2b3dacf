#!/bin/bash
# A/B of kernel variants: variants/lib_*.so against the in-tree library, two rounds
for round in 1 2; do
  python profiles/prof_driver.py --envs 4096 --K 1000 --reps 20 | sed "s/^/base /"
  for f in variants/lib_*.so; do
    SHIPSIM_LIB=$PWD/$f python profiles/prof_driver.py --envs 4096 --K 1000 --reps 20 | sed "s|^|$(basename $f) |"
  done
done
