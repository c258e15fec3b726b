"""Generates tests/golden/*.npz from the reference's shipped model constants (run in the build container, where
/root/reference exists; the GPU box only sees the generated fixtures).

Sources: CryptoNets/Weights.cs (Weights_0/1/3, Biases_2/3; `CryptoNets/CryptoNets.cs:33-72` says which layer uses which)
and LowLatencyCryptoNets/SmallModel.cs (`LoLaCryptonets.cs:280-329`)."""
import os
import re
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def arrays(path):
    src = open(path, encoding="utf-8-sig").read()
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"double\[\]\s+(\w+)\s*\{\s*get;\s*\}\s*=\s*new\s+double\[\]\s*\{(.*?)\};", src, re.S):
        vals = [float(x) for x in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", m.group(2))]
        out[m.group(1)] = np.array(vals, dtype=np.float64)
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    a = arrays(os.path.join(REF, "CryptoNets", "Weights.cs"))
    print({k: v.shape for k, v in a.items()})
    np.savez_compressed(os.path.join(OUT, "cryptonets_mnist_weights.npz"), **a)
    b = arrays(os.path.join(REF, "LowLatencyCryptoNets", "SmallModel.cs"))
    print({k: v.shape for k, v in b.items()})
    np.savez_compressed(os.path.join(OUT, "lola_small_weights.npz"), **b)
