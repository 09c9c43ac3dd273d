#!/usr/bin/env python
"""Average duration per kernel (template instantiation) from an ncu launch list (--metrics gpu__time_duration.sum --csv).
usage: launch_agg.py <csv>"""
import csv,re,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i; break
c={k:j for j,k in enumerate(h)}
agg=collections.Counter(); cnt=collections.Counter()
for r in rows[start+1:]:
    if len(r)<len(h): continue
    if r[c['Metric Name']]!='gpu__time_duration.sum': continue
    name=r[c['Kernel Name']]
    m=re.search(r'rmd_eval_kernel<\(int\)(\d), \(int\)(\d)(?:, \(bool\)(\d))?>',name)
    key='eval<%s,%s,%s>'%m.groups() if m else name.split('(')[0][-40:]
    v=float(r[c['Metric Value']].replace(',',''))
    u=r[c['Metric Unit']]
    v*= {'ns':1e-3,'us':1,'ms':1e3}.get(u,1)
    agg[key]+=v; cnt[key]+=1
tot=sum(agg.values())
for k,v in sorted(agg.items()): print('%-45s n=%3d per launch %8.1f us'%(k,cnt[k],v/cnt[k]))
