"""Run N un-synchronised forward passes with CNHE_TRACE_SLOW set and report wall time per step: locates host-side stalls."""
import json, os, sys, time
os.environ.setdefault("CNHE_TRACE_SLOW", "2.0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from cryptonets_b200.he import B200BfvFactory
from cryptonets_b200.interfaces import EMatrixFormat
from cryptonets_b200.networks import CRYPTONETS_PRIMES, synthetic_mnist

f = B200BfvFactory(CRYPTONETS_PRIMES, bench.BATCH, seed=1)
eng = f.engine
layers = bench.build_network(f)
x = np.rint(synthetic_mnist(bench.BATCH, seed=7) / 256.0 * 16.0)
xm = f.GetEncryptedMatrix(x, EMatrixFormat.ColumnMajor, 1)
xm.RegisterScale(16.0)
eng.set_option("multi_stream", 0)
for _ in range(3):
    bench.forward(layers, xm).Dispose()
eng.sync()
print("--- timed", file=sys.stderr, flush=True)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
issue = []
eng.timer_start()
t00 = time.perf_counter()
for _ in range(steps):
    t0 = time.perf_counter()
    bench.forward(layers, xm).Dispose()
    issue.append((time.perf_counter() - t0) * 1e3)
ms = eng.timer_stop_ms()
wall = (time.perf_counter() - t00) * 1e3
print(json.dumps({"steps": steps, "gpu_ms_per_step": round(ms / steps, 2), "wall_ms_per_step": round(wall / steps, 2),
                  "issue_ms": [round(v, 1) for v in issue]}))
