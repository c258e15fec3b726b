"""NTT micro-benchmark (BASELINE.json config 5): achieved algorithmic GB/s = 16*N bytes per residue polynomial / time."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cryptonets_b200.engine import Engine


def run(N, count, n_polys, iters=10):
    t = {4096: 40961, 8192: 65537, 16384: 65537}[N]
    eng = Engine([t], N, 10, 20, count)
    k = eng.k
    rng = np.random.default_rng(0)
    words = n_polys * N
    host = rng.integers(0, 1 << 35, words, dtype=np.uint64)
    d = eng.dev_from(host)
    out = {}
    for inverse in (False, True):
        for _ in range(3):
            eng.raw_ntt(d, d, n_polys, 0, k, inverse)
        eng.timer_start()
        for _ in range(iters):
            eng.raw_ntt(d, d, n_polys, 0, k, inverse)
        ms = eng.timer_stop_ms() / iters
        gbs = 16.0 * N * n_polys / (ms * 1e-3) / 1e9
        out["inv" if inverse else "fwd"] = dict(ms=ms, gbs=gbs, polys_per_s=n_polys / (ms * 1e-3))
    eng.dev_free(d)
    eng.close()
    return out


if __name__ == "__main__":
    peak = 6580.3
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    cases = [(4096, 3, 32768), (8192, 5, 16384), (16384, 6, 8192), (8192, 2, 16384)]
    if len(sys.argv) > 3:
        cases = [(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))]
    for N, count, n in cases:
        r = run(N, count, n)
        for d, v in r.items():
            print(json.dumps(dict(N=N, k=count, n_polys=n, dir=d, ms=round(v["ms"], 4), algo_GBs=round(v["gbs"], 1),
                                  frac_of_measured_hbm=round(v["gbs"] / peak, 3))))
