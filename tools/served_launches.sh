set -e
python tools/serve_direct_probe.py 32 2>&1 | tail -4 | cut -c1-600
D=$(ls -d /tmp/vvcdirect_* | head -1)
cd $D
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 40000 -c 6000 --csv --log-file /root/repo/gpurun_out/served_launches.csv /root/repo/oracle/_ref/EncoderAppServe -c /root/repo/oracle/_ref/encoder_intra.cfg -i in.yuv -wdt 128 -hgt 128 -fr 30 -f 1 -q 32 --InputBitDepth=10 --InternalBitDepth=10 --OutputBitDepth=10 -b ncu.bin > /root/repo/gpurun_out/served_ncu.log 2>&1 || echo ncu rc=$?
python /root/repo/tools/served_launch_agg.py /root/repo/gpurun_out/served_launches.csv
