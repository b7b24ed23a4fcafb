#!/bin/bash
# On the GPU box: full-set ncu capture of selected kernels of one eager step + SASS hot spots (tools/sass_hot.py).
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
cap() {   # name regex count kernel-indices
    timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$2" -c $3 \
        -f -o /tmp/${TAG}_$1 python tools/profile_step.py supervised > $OUT/${TAG}_ncu_$1.log 2>&1
    ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_raw.csv 2>/dev/null
    ncu -i /tmp/${TAG}_$1.ncu-rep --page source --csv > /tmp/${TAG}_$1_src.csv 2>/dev/null
    python tools/sass_hot.py /tmp/${TAG}_$1_src.csv $4 > $OUT/${TAG}_$1_sass_hot.txt 2>&1
}
cap gru2 "gru_fwd|gru_bwd" 4 "0,2"
cap bnglu2 "bnglu_bwd" 7 "3,5,6"
cap bnglu2f "bnglu_fwd" 7 "3,6"
du -sh $OUT
