"""cProfile of the host side of one CryptoNets-MNIST forward pass (issue only; the GPU runs asynchronously)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from cryptonets_b200.he import B200BfvFactory
from cryptonets_b200.interfaces import EMatrixFormat
from cryptonets_b200.networks import CRYPTONETS_PRIMES, synthetic_mnist

f = B200BfvFactory(CRYPTONETS_PRIMES, bench.BATCH, seed=1)
layers = bench.build_network(f)
x = np.rint(synthetic_mnist(bench.BATCH, seed=7) / 256.0 * 16.0)
xm = f.GetEncryptedMatrix(x, EMatrixFormat.ColumnMajor, 1)
xm.RegisterScale(16.0)
for _ in range(2):
    bench.forward(layers, xm).Dispose()
f.engine.sync()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    bench.forward(layers, xm).Dispose()
pr.disable()
f.engine.sync()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
