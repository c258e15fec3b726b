"""Sweep the multiply/key-switch chunk size (ciphertexts per kernel wave) on the CryptoNets-MNIST step: small chunks keep the
digit-NTT -> key-MAC and NTT -> tensor -> INTT intermediates inside the 126 MB L2."""
import json, os, sys
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
for chunk in [int(a) for a in sys.argv[1:]] or [128, 64, 32, 16, 8, 4]:
    eng.set_option("chunk", chunk)
    for _ in range(2):
        bench.forward(layers, xm).Dispose()
    eng.sync()
    eng.prof_enable(True)
    import time
    t0 = time.perf_counter()
    eng.timer_start()
    for _ in range(3):
        bench.forward(layers, xm).Dispose()
    host_ms = (time.perf_counter() - t0) * 1e3 / 3
    ms = eng.timer_stop_ms() / 3
    prof = eng.prof_collect()
    eng.prof_enable(False)
    print(json.dumps({"chunk": chunk, "ms_per_step": round(ms, 2), "host_issue_ms": round(host_ms, 2), "families_ms": {k: round(v["ms"] / 3, 2) for k, v in prof.items()}}), flush=True)
