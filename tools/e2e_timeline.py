"""Host-side breakdown of the pipelined end-to-end loop of bench.py (import / forward issue / export issue / wait), per step."""
import gc, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from cryptonets_b200.he import B200BfvFactory, B200BfvMatrix, B200BfvVector
from cryptonets_b200.interfaces import EMatrixFormat
from cryptonets_b200.networks import CRYPTONETS_PRIMES, synthetic_mnist

f = B200BfvFactory(CRYPTONETS_PRIMES, bench.BATCH, seed=1)
eng = f.engine
layers = bench.build_network(f)
x = np.rint(synthetic_mnist(bench.BATCH, seed=7) / 256.0 * 16.0)
xm = f.GetEncryptedMatrix(x, EMatrixFormat.ColumnMajor, 1)
xm.RegisterScale(16.0)
eng.set_option("multi_stream", int(os.environ.get("MS", "1")))
if os.environ.get("CHUNK"):
    eng.set_option("chunk", int(os.environ["CHUNK"]))
host_in = torch.empty(eng.P * 784 * eng.ct_words, dtype=torch.int64).pin_memory()
DEPTH = int(os.environ.get("DEPTH", "1"))  # batches queued ahead of the one the host waits for
host_outs = [torch.empty(eng.P * 10 * eng.ct_words, dtype=torch.int64).pin_memory() for _ in range(DEPTH + 1)]
eng.export_raw_many([v.vec for v in xm.vectors], host_in.data_ptr())


_dummy = None
_side = None
def imp():
    global _dummy, _side
    if os.environ.get("DUMMYCOPY"):  # the same bytes cross PCIe every step on a side stream, but nothing depends on them:
        if _dummy is None:           # separates hardware interference (DMA vs kernels) from scheduling / dependencies
            _dummy = torch.empty(host_in.numel(), dtype=torch.int64, device="cuda")
            _side = torch.cuda.Stream()
        with torch.cuda.stream(_side):
            _dummy.copy_(host_in, non_blocking=True)
        return None
    if os.environ.get("NOIMPORT"):  # isolate the upload: every step reuses the resident batch
        return None
    vecs = eng.import_raw_many(host_in.data_ptr(), 784, 1, bench.BATCH, 16.0)
    return B200BfvMatrix(f, [B200BfvVector(f, v) for v in vecs], EMatrixFormat.ColumnMajor, CopyVectors=False)


steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
# GC=off disables the cyclic collector in the loop, GC=freeze parks everything allocated so far in the permanent generation; every
# collection is timed (a full collection over the heap that torch and the network leave behind costs tens of milliseconds)
_gc_t0 = [0.0]
gc_log = []
def _gc_cb(phase, info):
    if phase == "start":
        _gc_t0[0] = time.perf_counter()
    else:
        gc_log.append((info["generation"], round((time.perf_counter() - _gc_t0[0]) * 1e3, 2)))
gc.callbacks.append(_gc_cb)
if os.environ.get("GC") == "off":
    gc.disable()
elif os.environ.get("GC") == "freeze":
    gc.collect()
    gc.freeze()
rows = []
t_start = time.perf_counter()
nxt = imp()
pending = []
for s in range(steps):
    t0 = time.perf_counter()
    cur = nxt
    out = bench.forward(layers, cur if cur is not None else xm)
    if cur is not None:
        cur.Dispose()
    t1 = time.perf_counter()
    ticket = eng.export_raw_many_async([v.vec for v in out.vectors], host_outs[s % (DEPTH + 1)].data_ptr())
    out.Dispose()
    t2 = time.perf_counter()
    if s + 1 < steps:
        nxt = imp()
    t3 = time.perf_counter()
    pending.append(ticket)
    if len(pending) > DEPTH:
        eng.export_wait(pending.pop(0))
    t4 = time.perf_counter()
    rows.append([round((b - a) * 1e3, 1) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4))])
for ticket in pending:
    eng.export_wait(ticket)
total = (time.perf_counter() - t_start) * 1e3
print("gc (generation, ms):", [g for g in gc_log if g[1] > 1.0], "collections:", len(gc_log), file=sys.stderr)
print(json.dumps({"steps": steps, "ms_per_step": round(total / steps, 2), "forward_export_import_wait_ms": rows}))
