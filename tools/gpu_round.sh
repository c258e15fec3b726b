#!/bin/bash
# one-GPU round job: full GPU test suite, smoke, headline bench, LoLa workloads, microbench, reference arm, launch list, ncu of the top kernel
mkdir -p gpurun_out
echo "== pytest -m gpu"; (timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench cryptonets"; timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_cryptonets.json; cut -c1-1200 gpurun_out/r02_bench_cryptonets.json
echo "== bench lola_small"; timeout 600 python bench.py --workload lola_small --steps 20 --warmup 3 2>>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_lola_small.json; cut -c1-700 gpurun_out/r02_bench_lola_small.json
echo "== bench lola_cifar"; timeout 900 python bench.py --workload lola_cifar --steps 3 --warmup 1 2>>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_lola_cifar.json; cut -c1-700 gpurun_out/r02_bench_lola_cifar.json
echo "== bench microbench"; timeout 600 python bench.py --workload microbench 2>>gpurun_out/r02_bench.err > gpurun_out/r02_bench_microbench.jsonl; cut -c1-260 gpurun_out/r02_bench_microbench.jsonl
cp profiles/r02_opcounts_*.json gpurun_out/ 2>/dev/null
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_reference.json; cut -c1-1500 gpurun_out/r02_bench_reference.json
tail -5 gpurun_out/r02_bench.err
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
echo "== ncu full: top kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_forward_digits_fp -s 3 -c 1 -o gpurun_out/r02_ntt_fwd_digits python tools/e2e_timeline.py 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mac_umma -s 2 -c 1 -o gpurun_out/r02_mac_umma_dense python tools/e2e_timeline.py 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
