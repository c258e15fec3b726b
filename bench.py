#!/usr/bin/env python
"""bench.py -- encrypted images/s on CryptoNets-MNIST (N=8192), the headline metric of BASELINE.json.

A step = one pass of the reference's timed region ("Batch-Time": after EncryptLayer, before Decrypt,
`CryptoNets/CryptoNets.cs:31,74`) over one 8192-image batch of synthetic MNIST-shaped inputs:
conv 5x5/2 (845 outputs) -> square -> dense 845->100 -> square -> dense 100->10, P plaintext moduli (default 2, the
reference's configuration, `CryptoNets.cs:17`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--plain-moduli 1|2]      B200 arm (one process per GPU under torchrun)
  python bench.py --impl reference ...                                          CPU arm: the in-repo C++ oracle (the
        reference's C#/SEAL path cannot be built here) on all host cores, each step a bounded sample of the same workload.

Prints ONE JSON line (rank 0).  `value` is the whole-job aggregate with inputs resident in HBM; `e2e` goes through the
public API with host (pinned) ciphertext buffers; `roofline` is the NTT kernel family measured live with CUDA events on
the library's stream."""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 8192
WORKLOAD = "CryptoNets-MNIST N=8192 k=5: conv5x5s2(845) > square > dense845x100 > square > dense100x10, 8192 images/batch"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(primes, threads, seed=0):
    """One bounded sample of the workload on the CPU oracle; returns (estimated seconds per full batch, description)."""
    from oracle.oracle_py import Oracle
    from cryptonets_b200.layers import ConvolutionEngine
    from cryptonets_b200.networks import cryptonets_weights, transpose
    w = cryptonets_weights()
    rng = np.random.default_rng(seed)
    ce = ConvolutionEngine()
    ce.InputShape, ce.KernelShape, ce.Stride, ce.Upperpadding, ce.MapCount = [28, 28], [5, 5], [2, 2], [1, 1], [5, 1]
    ce.Prepare()
    total = 0.0
    s_conv = max(1, 845 // max(64, 2 * threads))
    s_sq1 = max(1, 845 // max(32, 2 * threads))
    s_d3 = max(1, 100 // max(16, min(100, threads)))
    s_sq2 = max(1, 100 // max(16, min(100, 2 * threads)))
    for t in primes:
        o = Oracle(t, 8192, -1, 10, 20)
        o.keygen(1)
        ctw = o.ct_words
        q = np.array(o.q, dtype=np.uint64)

        def rand_cts(n):
            a = rng.integers(0, 1 << 43, (n, 2, o.k, 8192), dtype=np.uint64)
            return (a % q[None, None, :, None]).reshape(n, ctw)

        def lift(x):
            x = np.rint(x)
            return np.where(x < 0, x + t, x).astype(np.uint64)

        x = rand_cts(784)
        gather = np.array([[ce.Location(c, off, ce.InputShape) for off in ce.Offsets] for c in ce.Corners] * 5, dtype=np.int32)
        w0 = np.array([[w["Weights_0"][m * 26 + ce.Location(None, off, ce.KernelShape)] for off in ce.Offsets] for m in range(5)]) * 32
        wconv = lift(np.repeat(w0, 169, axis=0))
        bconv = lift(np.repeat(np.array([w["Weights_0"][(m + 1) * 26 - 1] for m in range(5)]) * 512, 169))
        t0 = time.perf_counter()
        o.mac_layer(x, gather, wconv, bconv, 845, 25, threads=threads, m_begin=0, m_step=s_conv)
        total += (time.perf_counter() - t0) * s_conv
        a1 = rand_cts(845)
        t0 = time.perf_counter()
        o.square_layer(a1, threads=threads, begin=0, step=s_sq1)
        total += (time.perf_counter() - t0) * s_sq1
        w1 = lift(transpose(w["Weights_1"], 845, 100).reshape(100, 845) * 1024)
        b1 = lift(np.rint(w["Biases_2"] * 1000.0) % t)
        t0 = time.perf_counter()
        o.mac_layer(a1, None, w1, b1, 100, 845, threads=threads, m_begin=0, m_step=s_d3)
        total += (time.perf_counter() - t0) * s_d3
        a2 = rand_cts(100)
        t0 = time.perf_counter()
        o.square_layer(a2, threads=threads, begin=0, step=s_sq2)
        total += (time.perf_counter() - t0) * s_sq2
        w3 = lift(w["Weights_3"].reshape(10, 100) * 32)
        b3 = lift(np.rint(w["Biases_3"] * 1000.0) % t)
        t0 = time.perf_counter()
        o.mac_layer(a2, None, w3, b3, 10, 100, threads=threads)
        total += time.perf_counter() - t0
    desc = ("per plaintext modulus: conv outputs every %d-th of 845, square1 every %d-th of 845, dense3 every %d-th of 100, square2 every "
            "%d-th of 100, dense5 all 10; layer times scaled by the sampling stride" % (s_conv, s_sq1, s_d3, s_sq2))
    return total, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cryptonets_b200.networks import CRYPTONETS_PRIMES
    primes = CRYPTONETS_PRIMES[: args.plain_moduli]
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_sample(primes, threads)
    times = []
    desc = ""
    for i in range(args.steps):
        s, desc = cpu_sample(primes, threads, seed=i)
        times.append(s)
    sec = float(np.mean(times))
    value = BATCH / sec
    print(json.dumps({
        "impl": "reference", "metric": "encrypted images/sec (CryptoNets-MNIST, N=8192)", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic (uniform residues; shipped CryptoNets weights)",
        "config": {"workload": WORKLOAD, "plain_moduli": len(primes), "note": "in-repo C++ oracle of the SEAL 3.2 path; the C#/SEAL reference cannot be built here"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------------------- B200 arm
def build_network(factory):
    """The CryptoNets-MNIST layer chain without its reader/encrypt layers (those sit before the timer)."""
    from cryptonets_b200.layers import PoolLayer, SquareActivation
    from cryptonets_b200.networks import cryptonets_weights, transpose

    class Src:
        Factory = factory

        def GetOutputScale(self):
            return 16.0

        def PrepareNetwork(self):
            pass

    w = cryptonets_weights()
    conv1 = PoolLayer(Source=Src(), InputShape=[28, 28], KernelShape=[5, 5], Upperpadding=[1, 1], Stride=[2, 2], MapCount=[5, 1], WeightsScale=32,
                      Weights=w["Weights_0"])
    act2 = SquareActivation(Source=conv1)
    dense3 = PoolLayer(Source=act2, InputShape=[845], KernelShape=[845], Stride=[1000], MapCount=[100], Weights=transpose(w["Weights_1"], 845, 100),
                       Bias=w["Biases_2"], WeightsScale=1024)
    act4 = SquareActivation(Source=dense3)
    dense5 = PoolLayer(Source=act4, InputShape=[100], KernelShape=[100], Stride=[1000], MapCount=[10], Weights=w["Weights_3"], Bias=w["Biases_3"],
                       WeightsScale=32)
    layers = [conv1, act2, dense3, act4, dense5]
    dense5.PrepareNetwork()
    return layers


def forward(layers, m):
    for layer in layers:
        nxt = layer.Apply(m)
        if layer is not layers[0]:
            m.Dispose()
        m = nxt
    return m


def run_b200(args):
    import torch
    import torch.distributed as dist
    from cryptonets_b200.he import B200BfvFactory, B200BfvMatrix, B200BfvVector
    from cryptonets_b200.interfaces import EMatrixFormat
    from cryptonets_b200.networks import CRYPTONETS_PRIMES, synthetic_mnist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    primes = CRYPTONETS_PRIMES[: args.plain_moduli]
    f = B200BfvFactory(primes, BATCH, seed=1 + rank, device=local)
    eng = f.engine
    layers = build_network(f)
    # every rank owns one batch (replicas over batches: the only split the slot packing allows)
    imgs = synthetic_mnist(BATCH, seed=20240917 + rank)
    x_raw = np.rint(imgs / 256.0 * 16.0)
    xm = f.GetEncryptedMatrix(x_raw, EMatrixFormat.ColumnMajor, 1)
    xm.RegisterScale(16.0)
    eng.sync()

    def barrier():
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the device-resident number is taken on ONE stream so that the per-kernel CUDA-event times are not inflated by kernels of the
    # other plaintext-modulus channel running concurrently; warm up in the same mode (the block recycler is per stream)
    eng.set_option("multi_stream", 0)
    for _ in range(args.warmup):
        forward(layers, xm).Dispose()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:  # one nvidia-smi loop per job, on the rank that reports
        sampler.start()
    launches0 = eng.launch_count()
    eng.prof_enable(True)
    eng.timer_start()
    last = None
    for _ in range(args.steps):
        if last is not None:
            last.Dispose()
        last = forward(layers, xm)
    ms = eng.timer_stop_ms()
    barrier()
    prof = eng.prof_collect()
    eng.prof_enable(False)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop()
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
        # the one exchange of the path: all-gather of the score ciphertexts (10 ct x P) over NVLink
        per = 10 * eng.ct_words
        mine = torch.empty(eng.P * per, dtype=torch.int64, device="cuda")
        for ch in range(eng.P):
            for j, v in enumerate(last.vectors):
                p, wds = v.vec.device_ptr(ch)
                eng.dev_copy(mine.data_ptr() + 8 * (ch * per + j * eng.ct_words), p, wds)
        eng.sync()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        torch.cuda.synchronize()
    value = BATCH * args.steps * world / (ms * 1e-3)

    # informational: the same K steps with one CUDA stream per plaintext-modulus channel (the two channels' kernels overlap each other's
    # tails; per-launch event times are then inflated by the concurrency, which is why the roofline above is taken on one stream)
    eng.set_option("multi_stream", 1)
    for _ in range(2):
        forward(layers, xm).Dispose()
    barrier()
    eng.timer_start()
    for _ in range(args.steps):
        forward(layers, xm).Dispose()
    ms2 = eng.timer_stop_ms()
    barrier()
    if world > 1:
        t2 = torch.tensor([ms2], dtype=torch.float64, device="cuda")
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms2 = float(t2.item())
    value_two_streams = BATCH * args.steps * world / (ms2 * 1e-3)

    # ---- e2e: host (pinned) ciphertexts in, score ciphertexts out, through the public API
    host_in = torch.empty(eng.P * 784 * eng.ct_words, dtype=torch.int64).pin_memory()
    host_out = torch.empty(eng.P * 10 * eng.ct_words, dtype=torch.int64).pin_memory()
    raw = eng.export_raw_many([v.vec for v in xm.vectors], host_in.data_ptr())
    del raw

    def e2e_import():
        vecs = eng.import_raw_many(host_in.data_ptr(), 784, 1, BATCH, 16.0)  # asynchronous: runs on the library's upload stream
        return B200BfvMatrix(f, [B200BfvVector(f, v) for v in vecs], EMatrixFormat.ColumnMajor, CopyVectors=False)

    host_outs = [host_out, torch.empty_like(host_out).pin_memory()]

    def e2e_run(steps):
        """`steps` batches: host ciphertexts in, score ciphertexts back on the host, pipelined one batch deep the way a serving loop
        is: batch i+1 is uploaded and queued while batch i computes; the host only ever waits for batch i-1's scores."""
        nxt = e2e_import()
        pending = None
        for s_ in range(steps):
            cur = nxt
            out = forward(layers, cur)
            cur.Dispose()
            ticket = eng.export_raw_many_async([v.vec for v in out.vectors], host_outs[s_ & 1].data_ptr())
            out.Dispose()  # stream-ordered: released after the copies above
            if s_ + 1 < steps:
                nxt = e2e_import()
            if pending is not None:
                eng.export_wait(pending)
            pending = ticket
        eng.export_wait(pending)

    eng.set_option("multi_stream", int(os.environ.get("CNHE_E2E_MULTI_STREAM", "1")))
    e2e_run(max(8, args.warmup))  # reaches the steady state of the upload slots and of the block recycler at pipeline depth 2
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
    e2e_value = BATCH * args.steps * world / e2e_s

    out = None
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        fam = prof["ntt_forward"]
        # dominant family: forward NTT (incl. the digit-decomposing variant of relinearisation)
        achieved = fam["bytes"] / (fam["ms"] * 1e-3) / 1e9 if fam["ms"] > 0 else 0.0
        roof = {"bound": "hbm", "kernel": "k_ntt_forward_fp / k_ntt_forward_digits_fp (N=8192), 16*N algorithmic bytes per transform", "achieved": achieved,
                "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "peak_source": peak_kind + " copy bandwidth (MEASURED_PEAKS.json)",
                # ncu (profiles/r01_square_path_v2_ncu.txt): one k_ntt_forward_digits_fp launch of 16000 transforms moved 42.3 MB + 995.9 MB of
                # DRAM traffic against 2097.2 MB algorithmic (16N per transform; the digit source is shared by 125 transforms through
                # L2) -- ratio 0.495, applied to this run's mean launch (waves are larger than the captured one)
                "traffic": 0.495 * fam["bytes"] / max(1, fam["launches"]),
                "traffic_source": "ncu dram bytes / algorithmic bytes = 0.495 for k_ntt_forward_digits_fp<13,1> (profiles/r01_square_path_v2_ncu.txt), scaled to this run's mean launch",
                "launches_timed": fam["launches"], "algorithmic_bytes_per_launch": fam["bytes"] / max(1, fam["launches"]),
                "avg_launch_ms": fam["ms"] / max(1, fam["launches"]), "share_of_step": fam["ms"] / ms if ms else None,
                "families_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()}}
        cpu_threads = os.cpu_count() or 1
        cpu_sec, cpu_desc = cpu_sample(primes, cpu_threads)
        out = {
            "metric": "encrypted images/sec (CryptoNets-MNIST, N=8192)", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic MNIST-shaped uint8 images (80% zeros), shipped CryptoNets weights, device-generated keys",
            "config": {"workload": WORKLOAD, "plain_moduli": len(primes), "parallelism": "replica-per-gpu x%d" % world,
                       "l2": "inputs larger than L2 (784 ct x 640 KiB per modulus)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(host_in.numel() * 8), "d2h_bytes_per_step": int(host_out.numel() * 8)},
            "value_two_streams": {"value": value_two_streams, "unit": "images/s", "ms_per_step": ms2 / args.steps,
                                  "note": "same steps, one CUDA stream per plaintext modulus; not used for the roofline"},
            "roofline": roof,
            "cpu_baseline": {"value": BATCH / cpu_sec, "unit": "images/s", "cores": cpu_threads, "kind": "port", "sample": cpu_desc},
            "readme_anchor_images_per_s": 320.0,
        }
    last.Dispose()
    xm.Dispose()
    f.Dispose()
    if world > 1:
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def run_microbench(args):
    """BASELINE config 5: NTT and multiply+relinearise micro-benchmark, GPU (libcnhe) next to the CPU oracle (this leg is a
    cpu_baseline: the oracle is timed, never used for results) on the box's host cores.  N in {4096, 8192, 16384}, k = 2..6
    coefficient moduli (prefixes of SEAL's default tables).  One JSON line per case."""
    from cryptonets_b200.engine import Engine
    from oracle.oracle_py import Oracle
    PEAK = 6580.3
    try:
        PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    threads = os.cpu_count() or 1
    cases = [(4096, 2), (4096, 3), (8192, 2), (8192, 4), (8192, 5), (16384, 4), (16384, 6)]
    t = {4096: 40961, 8192: 65537, 16384: 65537}
    rng = np.random.default_rng(0)
    for N, k in cases:
        eng = Engine([t[N]], N, 10, 20, k)
        eng.keygen(1)
        eng.set_option("multi_stream", 0)
        q = np.array(eng.q, dtype=np.uint64)
        n_polys = (1 << 27) // N // k * k  # 1 GiB of residue polynomials
        d = eng.dev_from(rng.integers(0, 1 << 35, n_polys * N, dtype=np.uint64))
        res = {"N": N, "k": k}
        for inverse in (False, True):
            for _ in range(2):
                eng.raw_ntt(d, d, n_polys, 0, k, inverse)
            eng.timer_start()
            for _ in range(5):
                eng.raw_ntt(d, d, n_polys, 0, k, inverse)
            ms = eng.timer_stop_ms() / 5
            res["gpu_%s_Mpolys_s" % ("intt" if inverse else "ntt")] = round(n_polys / ms / 1e3, 2)
            res["gpu_%s_frac_hbm" % ("intt" if inverse else "ntt")] = round(16.0 * N * n_polys / (ms * 1e-3) / 1e9 / PEAK, 3)
        eng.dev_free(d)
        # multiply + relinearise of n ciphertexts
        n = 256 if N <= 8192 else 128
        cts = (rng.integers(0, 1 << 62, (n, 2, k, N), dtype=np.uint64) % q[None, None, :, None]).astype(np.uint64)
        a = eng.dev_from(cts)
        out = eng.dev_alloc(n * 2 * k * N)
        for _ in range(2):
            eng.raw_multiply_relin(0, a, a, n, out)
        eng.timer_start()
        for _ in range(3):
            eng.raw_multiply_relin(0, a, a, n, out)
        res["gpu_square_relin_us_per_ct"] = round(eng.timer_stop_ms() / 3 * 1e3 / n, 2)
        eng.close()
        # CPU oracle, all host threads
        orc = Oracle(t[N], N, k, 10, 20)
        orc.keygen(1)
        cpu_polys = threads * 64 // k * k
        host = rng.integers(0, 1 << 35, cpu_polys * N, dtype=np.uint64)
        hp = host.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
        orc.L.orc_ntt_batch(orc.h, 0, hp, cpu_polys, 0, threads)  # in place; first call warms the pages and the thread pool
        t0 = time.perf_counter()
        for _ in range(3):
            orc.L.orc_ntt_batch(orc.h, 0, hp, cpu_polys, 0, threads)
        res["cpu_ntt_Mpolys_s"] = round(3 * cpu_polys / (time.perf_counter() - t0) / 1e6, 3)
        m = min(n, 2 * threads)
        t0 = time.perf_counter()
        orc.square_layer(cts[:m].reshape(m, -1), threads=threads)
        res["cpu_square_relin_us_per_ct"] = round((time.perf_counter() - t0) * 1e6 / m, 1)
        res["cpu_threads"] = threads
        res["speedup_ntt"] = round(res["gpu_ntt_Mpolys_s"] / res["cpu_ntt_Mpolys_s"], 1)
        res["speedup_square_relin"] = round(res["cpu_square_relin_us_per_ct"] / res["gpu_square_relin_us_per_ct"], 1)
        print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--plain-moduli", type=int, default=2, choices=[1, 2])
    ap.add_argument("--microbench", action="store_true", help="NTT / multiply+relinearise micro-benchmark (BASELINE config 5), one JSON line per case")
    args = ap.parse_args()
    if args.microbench:
        run_microbench(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
