#!/bin/bash
# round 2, session AE: A/B of the round-based kernel: cold paths out of line / cp.async key ring / padded records / bisection by selects
for v in ${VARIANTS:-base cold all}; do
  if [ "$v" != base ]; then export IB200_LIB=$PWD/build_exp/lib_$v.so; else unset IB200_LIB; fi
  timeout 300 python scripts/probe_variants.py 32768 2>&1 | tail -4
done
