#!/bin/bash
# ncu evidence for one round (run under gpurun, ONE GPU).  Usage: scripts/profile.sh r01
# Each ncu pass only after the identical plain command exited 0.  Outputs are kept small
# (gpurun_out/ is capped at 64 MiB): a launch list and --set full captures of a handful of launches.
set -u
TAG=${1:-r01}
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p $OUT
rm -f $OUT/*.ncu-rep
$CMD > $OUT/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 204 -c 140 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 186 -c 5 -o $OUT/prof_conv_$TAG $CMD > $OUT/ncu_conv_$TAG.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"nms_kernel|head_decode|stem_conv" -s 9 -c 3 -o $OUT/prof_post_$TAG $CMD > $OUT/ncu_post_$TAG.log 2>&1
for f in conv post; do
  ncu -i $OUT/prof_${f}_$TAG.ncu-rep --page raw --csv > $OUT/prof_${f}_${TAG}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_${f}_$TAG.ncu-rep --page details --csv > $OUT/prof_${f}_${TAG}_details.csv 2>/dev/null
done
du -sh $OUT; ls -la $OUT | tail -14
