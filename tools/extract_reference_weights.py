"""Generates cryptonets_b200/models/*.npz from the reference's shipped model constants (run in the build container, where
/root/reference exists; the GPU box only sees the generated files).  These are the trained parameters the reference apps embed or
read at start-up -- data, not code: nothing is hand-copied, this script is the only way they enter the repo.

Sources: CryptoNets/Weights.cs (Weights_0/1/3, Biases_2/3; `CryptoNets/CryptoNets.cs:33-72` says which layer uses which),
LowLatencyCryptoNets/SmallModel.cs (`LoLaCryptonets.cs:280-329`), CifarCryptoNet/CifarWeight.csv + CifarBias.csv (one layer per
line, `NeuralNetworks/WeightsReader.cs:24-36`, used by `LolaCifarCryptoNet.cs:27,74-75,102-103,120-121`) and
LowLatencyCryptoNets/MnistLargeWeight.csv + MnistLargeBias.csv (`LoLaCryptonets.cs:333,369-370,390-391,402-403`).
The CSV values are float32 numbers printed as doubles; they are stored as float32 when that round-trips exactly."""
import os
import re
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cryptonets_b200", "models")


def arrays(path):
    src = open(path, encoding="utf-8-sig").read()
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"double\[\]\s+(\w+)\s*\{\s*get;\s*\}\s*=\s*new\s+double\[\]\s*\{(.*?)\};", src, re.S):
        vals = [float(x) for x in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", m.group(2))]
        out[m.group(1)] = np.array(vals, dtype=np.float64)
    return out


def csv_layers(weights_csv, biases_csv):
    """WeightsReader: line i of each file is layer i."""
    out = {}
    for prefix, path in (("Weights", weights_csv), ("Biases", biases_csv)):
        with open(path, encoding="utf-8-sig") as f:
            for i, line in enumerate(l for l in f if l.strip()):
                v = np.array([float(x) for x in line.strip().split(",")], dtype=np.float64)
                v32 = v.astype(np.float32)
                out["%s_%d" % (prefix, i)] = v32 if np.array_equal(v32.astype(np.float64), v) else v
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    a = arrays(os.path.join(REF, "CryptoNets", "Weights.cs"))
    print({k: v.shape for k, v in a.items()})
    np.savez_compressed(os.path.join(OUT, "cryptonets_mnist_weights.npz"), **a)
    b = arrays(os.path.join(REF, "LowLatencyCryptoNets", "SmallModel.cs"))
    print({k: v.shape for k, v in b.items()})
    np.savez_compressed(os.path.join(OUT, "lola_small_weights.npz"), **b)
    c = csv_layers(os.path.join(REF, "CifarCryptoNet", "CifarWeight.csv"), os.path.join(REF, "CifarCryptoNet", "CifarBias.csv"))
    print({k: (v.shape, v.dtype) for k, v in c.items()})
    np.savez_compressed(os.path.join(OUT, "lola_cifar_weights.npz"), **c)
    d = csv_layers(os.path.join(REF, "LowLatencyCryptoNets", "MnistLargeWeight.csv"), os.path.join(REF, "LowLatencyCryptoNets", "MnistLargeBias.csv"))
    print({k: (v.shape, v.dtype) for k, v in d.items()})
    np.savez_compressed(os.path.join(OUT, "lola_large_weights.npz"), **d)
