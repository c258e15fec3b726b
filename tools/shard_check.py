"""Row-sharded LoLa-CIFAR on N GPUs of one box against the Raw (plaintext) backend: every rank holds the same keys and input ciphertexts,
computes its slice of the 5488-row dense layer, the partial ciphertexts are all-gathered (NCCL) and summed locally
(cryptonets_b200/parallel.py).  Run under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py
k = 9 coefficient primes (one more than the reference's 8) so that the last layer decrypts (profiles/r02_noise_trace.md)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cryptonets_b200 import networks as nets  # noqa: E402
from cryptonets_b200.he import B200BfvFactory  # noqa: E402
from cryptonets_b200.raw import RawFactory  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
k = int(sys.argv[1]) if len(sys.argv) > 1 else 9
f = B200BfvFactory(nets.CIFAR_PRIMES, 16384, DecompositionBitCount=60, GaloisDecompositionBitCount=60, SmallModulusCount=k, seed=1, device=local)
imgs = nets.synthetic_cifar(1)
net, _ = nets.lola_cifar(f, imgs, shard=(rank, world, None) if world > 1 else None)
net.PrepareNetwork()
t0 = time.time()
out = net.GetNext()
f.engine.sync()
dt = time.time() - t0
got = np.asarray(out.Decrypt()).reshape(-1)
res = {"rank": rank, "world": world, "k": k, "seconds_first_inference": dt, "scores": [float(x) for x in got]}
if rank == 0:
    raw_net, _ = nets.lola_cifar(RawFactory(16384), imgs)
    raw_net.PrepareNetwork()
    want = np.asarray(raw_net.GetNext().Decrypt()).reshape(-1)
    res["equals_raw"] = bool(np.allclose(got, want, rtol=1e-9, atol=1e-9))
    res["budget"] = min(f.engine.noise_budget(v.vec, ch, 0) for v in out.vectors for ch in range(2))
if world > 1:
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        res["all_ranks_identical_scores"] = all(r["scores"] == res["scores"] for r in allres)
if rank == 0:
    print(json.dumps(res))
    assert res["equals_raw"] and res.get("all_ranks_identical_scores", True)
f.Dispose()
if world > 1:
    dist.destroy_process_group()
