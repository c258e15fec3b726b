#!/bin/bash
mkdir -p gpurun_out
echo "== launch list (bench --steps 2 --warmup 1)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_under_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_bench_launches.csv')))
hdr=None;agg=collections.Counter();cnt=collections.Counter()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        try: v=float(d['Metric Value'].replace(',',''))
        except: continue
        u=d.get('Metric Unit','')
        v_us = v/1000.0 if u in ('ns','nsecond') else (v if u in ('us','usecond') else v*1000.0 if u in ('ms','msecond') else v)
        name=d['Kernel Name'].split('(')[0][:70]
        agg[name]+=v_us; cnt[name]+=1
tot=sum(agg.values())
out=open('gpurun_out/r02_bench_launch_list_summary.txt','w')
def P(*a):
    print(*a); print(*a,file=out)
P('# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 1 (serialised, cold-cache launch times: compare SHARES)')
P('total kernel time us', round(tot), 'launches', sum(cnt.values()))
for k,v in agg.most_common(30): P('%-72s %10.0f us %6.1f%% n=%d avg %.1f us'%(k,v,100*v/tot,cnt[k],v/cnt[k]))
PY
echo "== ncu full: k_mac_dense_umma"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mac_dense_umma -s 2 -c 1 -o gpurun_out/r02_mac_umma python tools/e2e_timeline.py 3 > /dev/null 2>&1
ls -la gpurun_out/r02_mac_umma.ncu-rep
