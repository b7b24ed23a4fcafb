#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full-set ncu capture of one eager training step.
# Keeps gpurun_out small: raw-page CSVs always, the .ncu-rep only when it is < 45 MB.
set -u
WL=${1:-supervised}
TAG=${2:-r1}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches_${WL}.csv python tools/profile_step.py $WL > $OUT/${TAG}_ncu_launch.log 2>&1
cap() {   # name regex count
    timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$2" -c $3 \
        -f -o /tmp/${TAG}_$1 python tools/profile_step.py $WL > $OUT/${TAG}_ncu_$1.log 2>&1
    ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_raw.csv 2>/dev/null
    ncu -i /tmp/${TAG}_$1.ncu-rep --page details --csv > $OUT/${TAG}_$1_details.csv 2>/dev/null
    sz=$(stat -c %s /tmp/${TAG}_$1.ncu-rep 2>/dev/null || echo 0)
    if [ "$sz" -gt 0 ] && [ "$sz" -lt 20000000 ]; then cp /tmp/${TAG}_$1.ncu-rep $OUT/; fi
}
cap gru "gru_fwd|gru_bwd" 4
cap heads "heads_fwd|heads_bwd" 2
cap bnglu "bnglu_bwd|bnglu_fwd" 14
cap conv "conv3x3|conv_wgrad|conv0" 26
du -sh $OUT
