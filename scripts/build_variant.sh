#!/bin/bash
# Builds an alternative libintegrals_b200.so for kernel experiments: recompiles ONE translation unit with extra nvcc flags and links it
# with the objects of the regular build.   scripts/build_variant.sh <name> <unit, e.g. hcubature_rb_d4> <fmad true|false> [extra flags...]
# -> build_exp/lib_<name>.so   (use with IB200_LIB=build_exp/lib_<name>.so)
set -e
cd "$(dirname "$0")/../integrals.jl_b200/csrc"
name=$1; unit=$2; fmad=$3; shift 3
out=../../build_exp
mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O2 -Xptxas -v -fmad=$fmad "$@" -c $unit.cu -o $out/${unit}_$name.o 2> $out/${unit}_$name.ptxas.log
objs=$(ls ../../gpurun_out/_build/*.o | grep -v "_build/$unit.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/lib_$name.so $objs $out/${unit}_$name.o
echo built $out/lib_$name.so
