#!/bin/bash
# usage: tools/ab_lists.sh "<lib dirs>" "<list sizes>" n K crc B : device-resident rate per list size and build
for L in $2; do for d in $1; do echo -n "$d: "; POLAR_B200_LIB_DIR=$PWD/polar_b200/$d python tools/list_rate.py $3 $4 $5 $L $6 2>&1 | tail -1; done; done
