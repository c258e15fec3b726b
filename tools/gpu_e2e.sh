#!/bin/bash
mkdir -p gpurun_out
show='import json,sys;d=json.loads(sys.stdin.read());r=d["forward_export_import_wait_ms"];import statistics as st;ss=[sum(x) for x in r[10:38]];print("steady ms/step", round(st.mean(ss),2), "min", round(min(ss),2), "max", round(max(ss),2))'
echo "== resident input + a dummy 1 GB H2D copy per step on a side stream"; DUMMYCOPY=1 GC=freeze MS=1 timeout 300 python tools/e2e_timeline.py 40 2>/dev/null | python -c "$show"
echo "== upload every step, nvidia-smi clocks during the loop"; (for i in 1 2 3 4 5 6; do sleep 2; nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv,noheader; done) & GC=freeze MS=1 timeout 300 python tools/e2e_timeline.py 400 2>/dev/null | python -c "$show"; wait
