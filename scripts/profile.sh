#!/bin/bash
# ncu evidence for one round (run under gpurun, ONE GPU).  Usage: scripts/profile.sh r01
# Every ncu pass runs only after the identical plain command exited 0.  The capture range is exactly ONE steady-state
# step of bench.py (cudaProfilerStart/Stop, --profile-from-start off).  Outputs stay small (gpurun_out/ is capped at
# 64 MiB): CSV exports only, .ncu-rep files are deleted on the box.
set -u
TAG=${1:-r01}
OUT=gpurun_out
rm -f $OUT/tune_$TAG.txt $OUT/tune_ms_$TAG.txt
# the plain run of step 1 times the per-layer kernel variants and writes its choices; every ncu pass (whose timings would be perturbed) reads them
CMD="python bench.py --profile-step --warmup 3 --tune-cache $OUT/tune_$TAG.txt"
mkdir -p $OUT
# 1. launch list of one steady-state step: duration + DRAM bytes per launch
$CMD > $OUT/plain_$TAG.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
# 2. --set full on a few launches of each hot kernel (same capture range)
prof() {  # name regex skip count
  $CMD > /dev/null 2>&1 && \
  ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o $OUT/prof_$1_$TAG $CMD > $OUT/ncu_$1_$TAG.log 2>&1
  ncu -i $OUT/prof_$1_$TAG.ncu-rep --page raw --csv > $OUT/prof_$1_${TAG}_raw.csv 2>/dev/null
  rm -f $OUT/prof_$1_$TAG.ncu-rep
}
prof gemm "conv_gemm_(pair_)?kernel" 0 6           # c2f_2.conv1 (1x1 64->64 @160), c2f_2.conv2, conv3 (3x3/s2), c2f_4.conv1, c2f_4.conv2, conv5
prof conv3 "conv3x3_(pair_)?kernel" 0 8            # conv1 (pair-line s2), the two 160x160 bottleneck convs, the 80x80 ones (CTA-pair kernel where the autotuner picked it)
prof post 'nms_kernel|head_decode|stem_t' 0 3
prof dec "conv_gemm_(pair_)?kernel" 26 9           # the LAST conv_gemm launches of the step: the six decode-fused final head convs (75 registers)
                                         # + the head's 20x20 3x3 layers the autotuner gave to the generic kernel
du -sh $OUT; ls $OUT | grep $TAG
# 3. MS-Block variant: launch list of one step + --set full of the fused layer kernel (160x160 pw1->dw->pw2, 80x80, 40x40 dw->pw2 / dw, 20x20)
CMDMS="python bench.py --block ms --profile-step --warmup 3 --tune-cache $OUT/tune_ms_$TAG.txt"
$CMDMS > /dev/null 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file $OUT/launches_ms_$TAG.csv $CMDMS > $OUT/ncu_launches_ms_$TAG.log 2>&1
$CMDMS > /dev/null 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:ms_layer_kernel" -s 0 -c 20 -o $OUT/prof_dw_$TAG $CMDMS > $OUT/ncu_dw_$TAG.log 2>&1
ncu -i $OUT/prof_dw_$TAG.ncu-rep --page raw --csv > $OUT/prof_dw_${TAG}_raw.csv 2>/dev/null
rm -f $OUT/prof_dw_$TAG.ncu-rep
