"""Per-layer host-issue time, wall time and summed kernel time of one CryptoNets-MNIST forward pass: finds host-side stalls."""
import json, os, sys, time
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
eng.set_option("multi_stream", int(os.environ.get("MS", "0")))
for _ in range(2):
    bench.forward(layers, xm).Dispose()
eng.sync()
for rep in range(2):
    m = xm
    for i, layer in enumerate(layers):
        eng.prof_enable(True)
        t0 = time.perf_counter()
        nxt = layer.Apply(m)
        t1 = time.perf_counter()
        eng.sync()
        t2 = time.perf_counter()
        prof = eng.prof_collect()
        eng.prof_enable(False)
        td = time.perf_counter()
        if m is not xm:
            m.Dispose()
        td2 = time.perf_counter()
        m = nxt
        print(json.dumps({"layer": i, "type": type(layer).__name__, "issue_ms": round((t1 - t0) * 1e3, 2), "wall_ms": round((t2 - t0) * 1e3, 2),
                          "kernels_ms": round(sum(v["ms"] for v in prof.values()), 2), "dispose_ms": round((td2 - td) * 1e3, 2)}), flush=True)
    m.Dispose()
