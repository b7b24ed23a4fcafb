#!/bin/bash
# round-2 GPU session: legs selected by name, every leg under its own timeout; outputs in gpurun_out/<TAG>_*
set -u
TAG=${1:-r2}
LEGS=${2:-"tests ab bench"}
OUT=gpurun_out
mkdir -p $OUT
has() { case " $LEGS " in *" $1 "*) return 0;; *) return 1;; esac; }
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $OUT/${TAG}_timeline.txt; }
if has tests; then
    timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest_gpu.log 2>&1
    stamp "pytest -m gpu exit $? : $(tail -1 $OUT/${TAG}_pytest_gpu.log)"
fi
if has newtests; then
    timeout 900 python -m pytest tests/test_trainer_gpu.py tests/test_variants_gpu.py -x -q -m gpu > $OUT/${TAG}_pytest_new.log 2>&1
    stamp "pytest new exit $? : $(tail -1 $OUT/${TAG}_pytest_new.log)"
fi
if has fe; then
    timeout 200 python tools/bench_frontend.py --ab > $OUT/${TAG}_frontend.txt 2>&1
    stamp "front-end A/B exit $?"
fi
if has ab; then
    timeout 300 python tools/bench_ab.py gru_v3 gru_v3=2 > $OUT/${TAG}_ab.txt 2>&1
    stamp "A/B exit $?"
fi
if has smoke; then
    timeout 200 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1
    stamp "smoke exit $? : $(tail -1 $OUT/${TAG}_smoke.log)"
fi
for wl in supervised mean_teacher inference dcase2024; do
    if has bench_$wl || has benchq; then
        timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --quick > $OUT/${TAG}_bench_${wl}.json 2> $OUT/${TAG}_bench_${wl}.err
        stamp "bench $wl (quick) exit $? : $(cut -c1-150 $OUT/${TAG}_bench_${wl}.json)"
    fi
done
if has bench; then
    timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
    stamp "bench full exit $?"
fi
for wl in mean_teacher inference; do
    if has full_$wl; then
        timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 > $OUT/${TAG}_bench_${wl}.json 2> $OUT/${TAG}_bench_${wl}.err
        stamp "bench $wl (full) exit $? : $(cut -c1-150 $OUT/${TAG}_bench_${wl}.json)"
    fi
done
if has ref; then
    timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
    stamp "reference arm exit $?"
fi
if has launches; then
    timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file $OUT/${TAG}_launches_supervised.csv python tools/profile_step.py supervised > $OUT/${TAG}_ncu_launch.log 2>&1
    stamp "ncu launch list exit $?"
fi
if has hot; then
    timeout 800 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -c 105 \
        -f -o /tmp/${TAG}_hot python tools/profile_step.py supervised > $OUT/${TAG}_ncu_hot.log 2>&1
    stamp "ncu hot capture exit $?"
    ncu -i /tmp/${TAG}_hot.ncu-rep --page raw --csv > $OUT/${TAG}_hot_raw.csv 2>/dev/null
    ncu -i /tmp/${TAG}_hot.ncu-rep --page source --csv > /tmp/${TAG}_hot_src.csv 2>/dev/null
    python tools/sass_hot.py /tmp/${TAG}_hot_src.csv > $OUT/${TAG}_hot_sass.txt 2>&1
    stamp "ncu export done"
fi
cat $OUT/${TAG}_timeline.txt
