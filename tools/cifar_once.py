"""Two LoLa-CIFAR inferences at the reference's parameters (the ncu launch-list target: `ncu --metrics gpu__time_duration.sum ...`)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cryptonets_b200 import networks as nets
from cryptonets_b200.he import B200BfvFactory

f = B200BfvFactory(nets.CIFAR_PRIMES, 16384, DecompositionBitCount=60, GaloisDecompositionBitCount=60, SmallModulusCount=8, seed=1)
f.engine.set_option("multi_stream", int(os.environ.get("MS", "0")))
net, rd = nets.lola_cifar(f, nets.synthetic_cifar(1))
net.PrepareNetwork()
chain = []
p = net
while p is not None and hasattr(p, "Source"):
    chain.append(p)
    p = p.Source
chain = chain[::-1]
x = chain[1].Apply(chain[0].GetNext())
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    t0 = time.time()
    m = x
    for layer in chain[2:]:
        nxt = layer.Apply(m)
        if m is not x:
            m.Dispose()
        m = nxt
    f.engine.sync()
    print("inference %d: %.3f s" % (rep, time.time() - t0), flush=True)
    m.Dispose()
f.Dispose()
