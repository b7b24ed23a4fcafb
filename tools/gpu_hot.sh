#!/bin/bash
# On the GPU box: full-set ncu capture (with SASS-level stall sampling) of the kernels furthest from their roofline.
set -u
TAG=${1:-r1b}
REGEX=${2:-"gru_fwd|gru_bwd|bnglu_bwd|logmel"}
COUNT=${3:-12}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$REGEX" -c $COUNT \
    -f -o /tmp/${TAG}_hot python tools/profile_step.py supervised > $OUT/${TAG}_ncu_hot.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/${TAG}_hot.ncu-rep --page raw --csv > $OUT/${TAG}_hot_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_hot.ncu-rep --page source --csv > /tmp/${TAG}_hot_src.csv 2>/dev/null
python tools/sass_hot.py /tmp/${TAG}_hot_src.csv > $OUT/${TAG}_hot_sass.txt 2>&1
gzip -9 -c /tmp/${TAG}_hot_src.csv > $OUT/${TAG}_hot_src.csv.gz
ls -la /tmp/${TAG}_hot.ncu-rep $OUT
