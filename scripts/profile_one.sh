#!/bin/bash
# usage: scripts/profile_one.sh <tag> <kernel-regex> <skip> <count>
TAG=$1; K=$2; S=$3; C=$4
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
rm -f $OUT/prof_$TAG.ncu-rep
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $C -o $OUT/prof_$TAG $CMD > $OUT/ncu_$TAG.log 2>&1
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv --print-source sass > $OUT/prof_${TAG}_src.csv 2>/dev/null
rm -f $OUT/prof_$TAG.ncu-rep
ls -la $OUT | grep $TAG
