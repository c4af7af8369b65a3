#!/bin/bash
# round 2, session AC: rows per thread of the C3 angle-addition loop: 2 / 4 / 8
for v in rows2 "" rows8; do
  if [ -n "$v" ]; then export IB200_LIB=$PWD/build_exp/lib_$v.so; else unset IB200_LIB; fi
  echo "variant ${v:-rows4}"; for i in 1 2; do python scripts/profile_kernel.py fixed genz_oscillatory 65536 2>&1 | tail -1; done
done
