#!/bin/bash
# ncu captures of the quantiser and rate kernels: in the TU sweep (bulk) and inside a served encode (walk-sized batches).  usage: tools/gpu_profile_tu_round.sh <tag>
tag=${1:-r2b}
mkdir -p gpurun_out
bash tools/served_launches.sh > gpurun_out/served_launch_agg_$tag.txt 2>&1; echo "served list rc=$?"
python tools/profile_tu.py --width 1920 --height 1080 --passes 2 > gpurun_out/prof_tu_plain_$tag.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dq_kernel|rate_kernel" -s 2 -c 2 -o gpurun_out/prof_tu_$tag -f \
  python tools/profile_tu.py --width 1920 --height 1080 --passes 2 > gpurun_out/ncu_full_tu_$tag.log 2>&1; echo "ncu tu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_tu_$tag.ncu-rep > gpurun_out/summary_tu_$tag.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_tu_$tag.ncu-rep 0 dq_kernel 30 dq_kernel > gpurun_out/lines_dq_$tag.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_tu_$tag.ncu-rep 0 rate_kernel 30 rate_kernel > gpurun_out/lines_rate_$tag.txt 2>&1
rm -f gpurun_out/*.ncu-rep
ls gpurun_out | head -40
