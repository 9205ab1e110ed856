#!/bin/bash
# usage: tools/build_variant.sh <dir under polar_b200/> [extra nvcc flags, e.g. -DPOLAR_TM_PIPE=0]
# builds libpolar_b200.so with the extra flags into polar_b200/<dir>/ and links libpolar_host.so next to it
# (A/B runs: POLAR_B200_LIB_DIR=$PWD/polar_b200/<dir>, see tools/ab.sh)
set -e
d=polar_b200/$1; shift
mkdir -p $d
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Iinclude -Ipolar_b200/csrc "$@" polar_b200/csrc/polar_b200.cu -o $d/libpolar_b200.so 2>/dev/null
nvcc -O2 -std=c++17 -Xcompiler -fPIC -shared -Iinclude -Ipolar_b200/csrc polar_b200/csrc/PolarCode.cpp -L$d -lpolar_b200 -Xlinker -rpath -Xlinker '$ORIGIN' -o $d/libpolar_host.so 2>/dev/null
ls -la $d
