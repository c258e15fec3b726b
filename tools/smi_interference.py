"""Does clock sampling disturb the launch path?  N un-synchronised steps (a) alone, (b) with `nvidia-smi -lms 100` running, (c) with
an in-process NVML sampler thread.  Prints GPU ms/step for each."""
import json, os, subprocess, sys, threading, time
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
eng.set_option("multi_stream", 0)
for _ in range(3):
    bench.forward(layers, xm).Dispose()
eng.sync()


def run(steps=8):
    eng.timer_start()
    for _ in range(steps):
        bench.forward(layers, xm).Dispose()
    return eng.timer_stop_ms() / steps


out = {"alone": [round(run(), 2) for _ in range(3)]}
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + bench.ClockSampler.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                     stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
time.sleep(0.5)
out["nvidia_smi_lms100"] = [round(run(), 2) for _ in range(3)]
p.terminate()
p.wait()
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    stop = False
    n = [0]

    def loop():
        while not stop:
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            n[0] += 1
            time.sleep(0.1)

    th = threading.Thread(target=loop, daemon=True)
    th.start()
    out["pynvml_100ms"] = [round(run(), 2) for _ in range(3)]
    stop = True
    out["pynvml_samples"] = n[0]
except Exception as e:
    out["pynvml"] = "unavailable: %s" % e
out["alone_again"] = [round(run(), 2) for _ in range(3)]
print(json.dumps(out))
