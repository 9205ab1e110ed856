#!/bin/bash
# usage: tools/build_variant.sh <dir under polar_b200/> [extra nvcc flags, e.g. -DPOLAR_TM_PIPE=0]
# builds both libraries with the extra flags into polar_b200/<dir>/
# (A/B runs: POLAR_B200_LIB_DIR=$PWD/polar_b200/<dir>, see tools/ab_dirs.sh)
set -e
d=$1; shift
python -m polar_b200.build --force --dir $d -- "$@"
