set -e
python tools/serve_direct_probe.py 32 2>&1 | tail -4 | cut -c1-600
D=$(ls -d /tmp/vvcdirect_* | head -1)
cd $D
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 40000 -c 6000 --csv --log-file /root/repo/gpurun_out/served_launches.csv /root/repo/oracle/_ref/EncoderAppServe -c /root/repo/oracle/_ref/encoder_intra.cfg -i in.yuv -wdt 128 -hgt 128 -fr 30 -f 1 -q 32 --InputBitDepth=10 --InternalBitDepth=10 --OutputBitDepth=10 -b ncu.bin > /root/repo/gpurun_out/served_ncu.log 2>&1 || echo ncu rc=$?
python - <<'P'
import csv,collections
rows=[r for r in csv.reader(l for l in open('/root/repo/gpurun_out/served_launches.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(',','')))
    except: pass
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print('%-62s n=%5d mean %8.1f us max %8.1f share %.3f'%(k,len(v),sum(v)/len(v)/1e3,max(v)/1e3,sum(v)/tot))
P
