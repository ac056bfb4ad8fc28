#!/bin/bash
# usage: scripts/profile_layer.sh <tag> <kernel-regex> B H W cin cout k s [res] [f32]
TAG=$1; K=$2; shift 2
OUT=gpurun_out
CMD="python scripts/one_layer.py $*"
$CMD && \
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 2 -c 1 -o $OUT/prof_$TAG $CMD > $OUT/ncu_$TAG.log 2>&1
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/prof_${TAG}_src.csv 2>/dev/null
rm -f $OUT/prof_$TAG.ncu-rep
