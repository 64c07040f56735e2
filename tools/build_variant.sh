#!/bin/bash
# tools/build_variant.sh <name> [nvcc -D flags...]  ->  acoss_b200/csrc/variants/lib_<name>.so
# A/B builds of the same ABI: k2_fast.cu is recompiled with the given flags, the other objects come from the
# main build (run acoss_b200/csrc/build.py first).  Select a variant at run time with ACOSS_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../acoss_b200/csrc"
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c k2_fast.cu -o variants/k2_fast_$name.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/lib_$name.so api.o k0_onramp.o k1_oti.o k2_exact.o k3_dp.o k4_knn.o k5_earlyfusion.o variants/k2_fast_$name.o
rm -f variants/k2_fast_$name.o
echo variants/lib_$name.so
