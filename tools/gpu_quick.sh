#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; (timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
echo "== e2e timeline MS=1"; MS=1 timeout 300 python tools/e2e_timeline.py 24 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['ms_per_step']);[print(r) for r in d['forward_export_import_wait_ms']]"
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_cryptonets.json; python -c "
import json;d=json.load(open('gpurun_out/r02_bench_cryptonets.json'));print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['roofline']['frac'])"
echo "== cifar timing"; timeout 300 python tools/cifar_once.py 3
echo "== cifar launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_cifar_launches.csv python tools/cifar_once.py 2 2>&1 | tail -3
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_cifar_launches.csv')))
hdr=None;agg=collections.Counter();cnt=collections.Counter()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        try: v=float(d['Metric Value'].replace(',',''))
        except: continue
        u=d.get('Metric Unit','')
        v_us = v/1000.0 if u in ('ns','nsecond') else (v if u in ('us','usecond') else v*1000.0 if u in ('ms','msecond') else v)
        name=d['Kernel Name'].split('(')[0][:60]
        agg[name]+=v_us; cnt[name]+=1
tot=sum(agg.values())
print('total kernel time us', round(tot), 'launches', sum(cnt.values()))
for k,v in agg.most_common(25): print('%-62s %10.0f us %6.1f%% n=%d'%(k,v,100*v/tot,cnt[k]))
PY
