# A/B of the scratch-plane layout: build the other layout first,
#   nvcc ... -DVVCB_SCRATCH_VISIT_MAJOR=0 -o vvc_intra_b200/libvvc_sm.so vvc_intra_b200/csrc/vvcb_api.cu
for v in "" vvc_intra_b200/libvvc_sm.so; do
  echo "variant: ${v:-default}"
  VVCB_LIBRARY_PATH=${v:+$PWD/$v} python bench.py --no-bitexact --no-strong 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d.get('kernel_ms'), round(d['e2e']['value']))"
done
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"rmd_eval_kernel|rmd_lists_kernel" -s 34 -c 34 --csv --log-file gpurun_out/traffic_vm.csv python tools/profile_sweep.py --width 1920 --height 1080 --passes 2 > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/traffic_vm.csv') if l.startswith('"'))]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
tot={}
for r in rows[1:]:
    k='lists' if 'lists' in r[ki] else 'eval'
    v=float(r[vi].replace(',',''))
    u=r[ui]
    if 'byte' in u.lower():
        mult={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(u,1)
        v*=mult
    tot[(k,r[mi])]=tot.get((k,r[mi]),0)+v
for k,v in sorted(tot.items()): print(k, round(v/1e6,1) if 'bytes' in k[1] else v)
P
