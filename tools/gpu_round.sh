#!/bin/bash
# One gpurun call = the whole round-end evidence set, most important first, every leg under its own timeout:
#   GPU parity tests, smoke, bench (ours N=1), ncu launch list, ncu full-set capture of the hot kernels of one eager step,
#   reference arm, mean-teacher bench.  Everything lands in gpurun_out/<TAG>_*.
set -u
TAG=${1:-r1}
LEGS=${2:-"tests smoke bench launches hot ref mt"}
OUT=gpurun_out
mkdir -p $OUT
has() { case " $LEGS " in *" $1 "*) return 0;; *) return 1;; esac; }
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $OUT/${TAG}_timeline.txt; }

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if has tests; then
    timeout 600 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1
    stamp "pytest -m gpu exit $? : $(tail -1 $OUT/${TAG}_pytest_gpu.log)"
fi
if has smoke; then
    timeout 200 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
    stamp "smoke exit $? : $(tail -1 $OUT/${TAG}_smoke.log)"
fi
if has bench; then
    timeout 420 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
    stamp "bench exit $?"
fi
if has launches; then
    timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file $OUT/${TAG}_launches_supervised.csv python tools/profile_step.py supervised > $OUT/${TAG}_ncu_launch.log 2>&1
    stamp "ncu launch list exit $?"
fi
if has hot; then
    # full-set capture (with SASS-level stall sampling) of the kernels that carry the step: GRU recurrence, tcgen05 BN+GLU,
    # register-resident BN+GLU, front end, tcgen05 convolutions and GEMMs
    timeout 420 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:"gru_fwd_v2|gru_bwd_v2|bnglu_tc5|bnglu_small|logmel|gemm_tc5|conv3x3_tc5|conv_wgrad_tc5|conv0" -c 48 \
        -f -o /tmp/${TAG}_hot python tools/profile_step.py supervised > $OUT/${TAG}_ncu_hot.log 2>&1
    stamp "ncu hot capture exit $?"
    ncu -i /tmp/${TAG}_hot.ncu-rep --page raw --csv > $OUT/${TAG}_hot_raw.csv 2>/dev/null
    ncu -i /tmp/${TAG}_hot.ncu-rep --page source --csv > /tmp/${TAG}_hot_src.csv 2>/dev/null
    python tools/sass_hot.py /tmp/${TAG}_hot_src.csv > $OUT/${TAG}_hot_sass.txt 2>&1
    stamp "ncu export done"
fi
if has ref; then
    timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
    stamp "reference arm exit $?"
fi
if has mt; then
    timeout 300 python bench.py --workload mean_teacher --steps 20 > $OUT/${TAG}_bench_mt.json 2> $OUT/${TAG}_bench_mt.err
    stamp "mean-teacher bench exit $?"
fi
du -sh $OUT
cat $OUT/${TAG}_timeline.txt
