"""Per-operation noise-budget trace of the shipped LoLa topologies at the reference's own parameters (runs on the GPU box).

For every topology the network is applied layer by layer next to the Raw (plaintext) backend; with the library option
"trace_noise" on, every evaluator-level operation (the granularity of the reference's OperationsCount / CryptoTracker,
`HE Wrapper/AtomicSealBfvVector.cs:211-294`, `HE Wrapper/CryptoTracker.cs:41-52`) records the invariant noise budget of its first
output ciphertext.  The raw trace goes to a JSON file; tools/noise_model.py renders it next to the analytic model (offline).

usage: python tools/noise_trace.py [--out gpurun_out/noise_trace.json] [--only lola_small,lola_cifar] [--mtilde 0|1]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from cryptonets_b200 import networks as nets  # noqa: E402
from cryptonets_b200.he import B200BfvFactory  # noqa: E402
from cryptonets_b200.raw import RawFactory  # noqa: E402

# name -> (builder, primes, N, decomposition bit count, the reference's SmallModulusCount, reference line, image maker)
TOPOLOGIES = {
    "lola_small": (nets.lola_small, nets.LOLA_SMALL_PRIMES, 8192, 40, 3, "LoLaCryptonets.cs:285", nets.synthetic_mnist),
    "lola": (nets.lola, nets.LOLA_PRIMES, 8192, None, 5, "LoLaCryptonets.cs:208", nets.synthetic_mnist),
    "lola_dense": (nets.lola_dense, nets.LOLA_DENSE_PRIMES, 16384, 60, 7, "LoLaCryptonets.cs:123", nets.synthetic_mnist),
    "lola_large": (nets.lola_large, nets.LOLA_LARGE_PRIMES, 16384, 60, 7, "LoLaCryptonets.cs:336", nets.synthetic_mnist),
    "lola_cifar": (nets.lola_cifar, nets.CIFAR_PRIMES, 16384, 60, 8, "LolaCifarCryptoNet.cs:35", nets.synthetic_cifar),
}


def layer_chain(net):
    out, p = [], net
    while p is not None and hasattr(p, "Source"):
        out.append(p)
        p = p.Source
    return out[::-1]


def mac_gain(layer, primes):
    """sqrt(sum w^2) of output 0 of a scalar-MAC layer, per plaintext modulus (weights as centred residues)."""
    if getattr(layer, "Weights", None) is None or not hasattr(layer, "kernelSize"):
        return None
    ks = layer.kernelSize
    w = np.rint(np.asarray(layer.Weights[:ks], dtype=np.float64)[: len(layer.Offsets)] * layer.WeightsScale).astype(object)
    out = []
    for t in primes:
        r = [int(x) % t for x in w]
        r = [x - t if x > t // 2 else x for x in r]
        out.append(float(np.sqrt(float(sum(x * x for x in r)))))
    return out


def run(name, k, mtilde, seed=5, weights="shipped"):
    builder, primes, N, dbc, k_ref, ref_line, imgs_of = TOPOLOGIES[name]
    kw = {} if dbc is None else dict(DecompositionBitCount=dbc, GaloisDecompositionBitCount=dbc)
    f = B200BfvFactory(primes, N, SmallModulusCount=k, seed=seed, **kw)
    eng = f.engine
    rec = dict(topology=name, N=N, k=k, k_reference=k_ref, reference=ref_line, primes=[int(p) for p in primes], q=[int(x) for x in eng.q],
               dbc=dbc or 10, dbc_galois=dbc or 20, mtilde_centered=mtilde, weights=weights, layers=[])
    try:
        eng.set_option("behz_centered_mtilde", mtilde)
        imgs = imgs_of(1, seed=6) if imgs_of is nets.synthetic_mnist else imgs_of(1)
        wts = None
        if weights == "synthetic" and name == "lola_cifar":
            wts = nets.cifar_weights(synthetic=True)
        if weights == "synthetic" and name == "lola_large":
            wts = nets.lola_large_weights(synthetic=True)
        net, rd = builder(f, imgs, weights=wts) if wts is not None else builder(f, imgs)
        net.PrepareNetwork()
        raw_net, rrd = builder(RawFactory(N), imgs, weights=wts) if wts is not None else builder(RawFactory(N), imgs)
        raw_net.PrepareNetwork()
        eng.op_counts(reset=True)
        eng.trace_noise(True)
        ma, mb = rd.GetNext(), rrd.GetNext()
        for A, B in list(zip(layer_chain(net), layer_chain(raw_net)))[1:]:
            t0 = time.time()
            ma, mb = A.Apply(ma), B.Apply(mb)
            eng.sync()
            dt = time.time() - t0
            ops = eng.trace_read(clear=True)
            budgets = [eng.noise_budget(v.vec, ch, b) for v in ma.vectors for ch in range(eng.P) for b in range(v.vec.blocks)]
            got, want = np.asarray(ma.Decrypt(), dtype=np.float64), np.asarray(mb.Decrypt(), dtype=np.float64)
            equal = bool(got.shape == want.shape and np.allclose(got, want, rtol=1e-9, atol=1e-9))
            eng.trace_read(clear=True)  # decryptions are not evaluator operations
            rec["layers"].append(dict(layer=type(A).__name__, seconds=dt, ops=ops, out_budget_min=int(min(budgets)), out_budget_max=int(max(budgets)),
                                      n_out_ct=len(budgets), equals_raw=equal, mac_gain=mac_gain(A, primes), counts=eng.op_counts(reset=True)))
            print("%-12s k=%d %-24s budget %3d..%3d  == Raw: %s  (%d ops, %.2fs)" % (name, k, type(A).__name__, min(budgets), max(budgets), equal,
                                                                                  len(ops), dt), flush=True)
        rec["final_equals_raw"] = rec["layers"][-1]["equals_raw"]
        rec["final_argmax_equal"] = bool(np.asarray(got).reshape(-1).argmax() == np.asarray(want).reshape(-1).argmax())
    finally:
        f.Dispose()
    return rec


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/noise_trace.json")
    ap.add_argument("--only", default="")
    ap.add_argument("--extra-prime", action="store_true", help="also run each topology with one more coefficient prime")
    args = ap.parse_args()
    names = [n for n in args.only.split(",") if n] or list(TOPOLOGIES)
    out = []
    for n in names:
        k_ref = TOPOLOGIES[n][4]
        for mt in (0, 1):
            out.append(run(n, k_ref, mt))
        if args.extra_prime and n != "lola":
            out.append(run(n, k_ref + 1, 0))
        if n in ("lola_cifar", "lola_large"):
            out.append(run(n, k_ref, 0, weights="synthetic"))
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(out, open(args.out, "w"))
    print("wrote", args.out)
