#!/bin/bash
# ncu --set full captures of the dependent-quantisation kernel inside a real served encode (direct mode, one CU per call); GPU box only.
set -e
cd /root/repo
D=$(mktemp -d /tmp/vvcdq_XXXX)
python - "$D" <<'P'
import sys, os
sys.path[:0] = ['/root/repo', '/root/repo/tools']
import bench
bench.write_crop(sys.argv[1], [bench.synth_luma(k) for k in range(2)], 1)
P
cd $D
K=${1:-dq_kernel}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip ${2:-3000} -c ${3:-24} -o /root/repo/gpurun_out/served_$K -f \
  /root/repo/oracle/_ref/EncoderAppServe -c /root/repo/oracle/_ref/encoder_intra.cfg -i in.yuv -wdt 128 -hgt 128 -fr 30 -f 1 -q 32 --InputBitDepth=10 --InternalBitDepth=10 --OutputBitDepth=10 -b ncu.bin > /root/repo/gpurun_out/served_$K.log 2>&1 || echo ncu rc=$?
ls -la /root/repo/gpurun_out/served_$K.ncu-rep
