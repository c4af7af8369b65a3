#!/bin/bash
# select-pass variants: 16 keys double buffered (library), 8 keys double buffered, 8 keys not double buffered
mkdir -p gpurun_out
for v in main ku8 ku8nodb; do
  if [ $v = main ]; then unset IB200_LIB; else export IB200_LIB=build_exp/lib_$v.so; fi
  timeout 300 python scripts/probe_rb.py 65536 1 genz_product_peak,genz_oscillatory,genz_gaussian,genz_c0 > gpurun_out/r2i_probe_$v.log 2>&1
  echo "== $v"; grep -E "mode 1 \{" gpurun_out/r2i_probe_$v.log | sed -E "s/'integrals_per_s.*gevals_per_s': ([0-9.]+).*/gevals \1/"
done
