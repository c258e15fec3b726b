"""Square-activation micro-benchmark: multiply+relinearise of n ciphertexts through the raw ABI (N=8192, SEAL default q,
dbc=10), per-family device times.  Used for A/B runs of kernel variants (env knobs) and as the ncu target for those kernels."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cryptonets_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
eng = Engine([549764251649], 8192, 10, 20)
eng.keygen(1)
eng.set_option("multi_stream", 0)
k, N = eng.k, 8192
rng = np.random.default_rng(0)
q = np.array(eng.q, dtype=np.uint64)
host = (rng.integers(0, 1 << 62, (n, 2, k, N), dtype=np.uint64) % q[None, None, :, None]).astype(np.uint64)
a = eng.dev_from(host)
out = eng.dev_alloc(n * 2 * k * N)
for _ in range(2):
    eng.raw_multiply_relin(0, a, a, n, out)
eng.sync()
eng.prof_enable(True)
t0 = time.perf_counter()
eng.timer_start()
for _ in range(iters):
    eng.raw_multiply_relin(0, a, a, n, out)
host_ms = (time.perf_counter() - t0) * 1e3 / iters
ms = eng.timer_stop_ms() / iters
prof = eng.prof_collect()
print(json.dumps({"n": n, "ms": round(ms, 3), "host_issue_ms": round(host_ms, 3), "us_per_ct": round(ms * 1e3 / n, 2),
                  "families_ms": {k_: round(v["ms"] / iters, 3) for k_, v in prof.items() if v["ms"] > 0}}))
