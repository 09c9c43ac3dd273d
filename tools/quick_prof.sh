#!/bin/bash
# quick profiling pass: launch list of one resident sweep + full captures of selected eval launches (second pass)
# usage: tools/quick_prof.sh <tag> <launch-index:name> ...
tag=$1; shift
python tools/profile_sweep.py --width 1920 --height 1080 --passes 4 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv python tools/profile_sweep.py --width 1920 --height 1080 --passes 2 > /dev/null 2>&1
python tools/launch_agg.py gpurun_out/launches_$tag.csv
for k in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rmd_eval_kernel -s ${k%%:*} -c 1 -o gpurun_out/prof_${k##*:}_$tag -f \
    python tools/profile_sweep.py --width 1920 --height 1080 --passes 2 > gpurun_out/ncu_full_${k##*:}_$tag.log 2>&1; echo "ncu ${k##*:} rc=$?"
done
