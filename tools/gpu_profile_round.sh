#!/bin/bash
# The ncu part of tools/gpu_round.sh alone (parity tests and bench are run separately): launch list of the bench command, DRAM traffic of
# the evaluation launches, full captures of the dominant kernels, text summaries for profiles/.  usage: tools/gpu_profile_round.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --resident-only > /dev/null 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 3 --resident-only > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu list rc=$?"
python tools/launch_agg.py gpurun_out/launches_$tag.csv > gpurun_out/launch_agg_$tag.txt 2>&1
python tools/profile_sweep.py --width 1920 --height 1080 --passes 2 > gpurun_out/prof_plain_$tag.log 2>&1 || exit 1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:rmd_eval_kernel -s 33 -c 33 \
  --csv --log-file gpurun_out/eval_traffic_$tag.csv python tools/profile_sweep.py --width 1920 --height 1080 --passes 2 > /dev/null 2>&1 && \
  python tools/make_traffic.py gpurun_out/eval_traffic_$tag.csv $tag
for k in 33:eval8x8ang 34:eval16x16ang 36:eval32x8ang; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:rmd_eval_kernel -s ${k%%:*} -c 1 -o gpurun_out/prof_${k##*:}_$tag -f \
    python tools/profile_sweep.py --width 1920 --height 1080 --passes 2 > gpurun_out/ncu_full_${k##*:}_$tag.log 2>&1; echo "ncu ${k##*:} rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rmd_lists_kernel -s 1 -c 1 -o gpurun_out/prof_lists_$tag -f \
  python tools/profile_sweep.py --width 1920 --height 1080 --passes 2 > gpurun_out/ncu_full_lists_$tag.log 2>&1; echo "ncu lists rc=$?"
python tools/profile_tu.py --width 1920 --height 1080 --passes 2 > gpurun_out/prof_tu_plain_$tag.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dq_kernel|tu_eval_kernel" -s 3 -c 3 -o gpurun_out/prof_tu_$tag -f \
  python tools/profile_tu.py --width 1920 --height 1080 --passes 2 > gpurun_out/ncu_full_tu_$tag.log 2>&1; echo "ncu tu rc=$?"
for k in eval8x8ang eval16x16ang eval32x8ang lists tu; do
  [ -f gpurun_out/prof_${k}_$tag.ncu-rep ] && python tools/ncu_summary.py gpurun_out/prof_${k}_$tag.ncu-rep > gpurun_out/summary_${k}_$tag.txt 2>&1
done
python tools/ncu_lines.py gpurun_out/prof_eval8x8ang_$tag.ncu-rep 0 ILi3ELi0ELi1 40 > gpurun_out/lines_eval8x8ang_$tag.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_eval16x16ang_$tag.ncu-rep 0 ILi3ELi0ELi0 40 > gpurun_out/lines_eval16x16ang_$tag.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_lists_$tag.ncu-rep 0 rmd_lists_kernel 30 rmd_lists_kernel > gpurun_out/lines_lists_$tag.txt 2>&1
python tools/ncu_opmix.py gpurun_out/prof_eval16x16ang_$tag.ncu-rep 0 > gpurun_out/opmix_eval16x16ang_$tag.txt 2>&1
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
