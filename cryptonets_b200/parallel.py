"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU tests).

What the path allows:
  * replicas -- the 8192 images of a CryptoNets batch share every ciphertext (SIMD slot packing, `CryptoNets/CryptoNets.cs:15-26`), so a
    batch cannot be split by image: every rank owns whole batches (LoLa: whole images) and the only exchange is the all-gather of the
    final score ciphertexts (10 x P per batch) at the decrypt/score step -- `gather_score_ciphertexts` / `ScoreGatherer`;
  * row shards inside ONE inference -- a row-major dense layer (CIFAR: 5488 rows) is split over the ranks; with ForceDenseFormat every
    rank's partial result is one ciphertext whose rows sit at their global columns, and the ranks' partials ADD UP to the product
    (`HE Wrapper/EncryptedSealBfvMatrix.cs:92-116` sums the masked rows the same way).  Modular addition is not an NCCL reduction, so the
    exchange is an all-gather of the partial ciphertexts followed by a local `ct_add` chain -- `allreduce_ciphertext_sum`;
    without ForceDense the slices are concatenated -- `allgather_sparse_elements`.
Also here: `bind_to_gpu_numa` (host threads and pinned buffers next to the GPU they feed)."""
import os

import numpy as np
import torch
import torch.distributed as dist


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_initialized() else 1


def rank_of(group=None):
    return dist.get_rank(group) if dist.is_initialized() else 0


def batches_of_rank(n_batches, rank, world):
    """Round-robin assignment of batch indices to ranks (weak scaling: every rank gets the same count +-1)."""
    return list(range(rank, n_batches, world))


def row_slice(n_rows, rank, world):
    """Contiguous, balanced slice [first, first + count) of n_rows for `rank` (5488 rows over 4 ranks -> 1372 each)."""
    base, extra = divmod(n_rows, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def gather_score_ciphertexts(local_words, group=None):
    """all-gather a rank's score ciphertexts (1-D int64 tensor of raw words, any device) -> list of tensors, one per rank."""
    world = world_size(group)
    if world == 1:
        return [local_words]
    out = [torch.empty_like(local_words) for _ in range(world)]
    dist.all_gather(out, local_words, group=group)
    return out


def max_over_ranks(value, device="cpu", group=None):
    """device time of a step is the max over ranks"""
    if world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class _DevView:
    """zero-copy torch view of raw device words (the library owns the memory)"""

    def __init__(self, ptr, words):
        self.__cuda_array_interface__ = {"shape": (int(words),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def device_view(ptr, words, device):
    return torch.as_tensor(_DevView(ptr, words), device=device)


class ScoreGatherer:
    """All-gather of a batch's score ciphertexts over NVLink -- the one exchange of the replicated path -- ordered on the device, without a
    host synchronisation and without coupling the plaintext-modulus channels to each other: every channel packs its own score ciphertexts
    on ITS stream (torch.cuda.ExternalStream over cnhe_context_stream), a side stream waits for those copies (events) and runs the NCCL
    collective, so the vectors may be disposed right after and the next batch's kernels never wait for it.  Two buffer sets alternate; a
    channel touching a set again first waits for the collective that read it two batches ago."""

    def __init__(self, eng, n_vectors, device, group=None):
        self.eng, self.group, self.device = eng, group, device
        self.per = n_vectors * eng.ct_words
        self.mine = [torch.empty(eng.P * self.per, dtype=torch.int64, device=device) for _ in range(2)]
        self.all = [torch.empty(world_size(group) * eng.P * self.per, dtype=torch.int64, device=device) for _ in range(2)]
        self.side = torch.cuda.Stream(device=device)
        self.done = [None, None]
        self.turn = 0

    def gather(self, vecs):
        """vecs: the score vectors of the batch just queued (views of one slab per channel handed out by the layer call)."""
        i = self.turn
        self.turn ^= 1
        mine, out = self.mine[i], self.all[i]
        chans = {}
        for ch in range(self.eng.P):
            chans.setdefault(self.eng.stream(ch), []).append(ch)  # single-stream mode: every channel reports stream 0
        for sptr, chs in chans.items():
            ext = torch.cuda.ExternalStream(sptr, device=self.device)
            if self.done[i] is not None:
                ext.wait_event(self.done[i])
            with torch.cuda.stream(ext):
                for ch in chs:
                    for j, v in enumerate(vecs):
                        p, wds = v.device_ptr(ch)
                        mine[ch * self.per + j * wds: ch * self.per + (j + 1) * wds].copy_(device_view(p, wds, self.device), non_blocking=True)
            self.side.wait_stream(ext)
        with torch.cuda.stream(self.side):
            if world_size(self.group) > 1:
                dist.all_gather_into_tensor(out, mine, group=self.group)
            else:
                out.copy_(mine, non_blocking=True)
            self.done[i] = torch.cuda.Event()
            self.done[i].record(self.side)
        return out

    def finish(self):
        """host-side wait for the collectives queued so far (end of a timed region)"""
        self.side.synchronize()


def allreduce_ciphertext_sum(factory, vec, group=None):
    """Sum of one dense single-block encrypted vector per rank (the partial products of a row-sharded ForceDense layer): all-gather of the
    P x 2kN raw words, then local homomorphic additions in rank order (identical on every rank).  Returns a new B200BfvVector."""
    from .he import B200BfvVector
    eng = factory.engine
    world = world_size(group)
    if world == 1:
        return vec
    blocks = vec.vec.blocks
    words = eng.P * blocks * eng.ct_words
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    mine = torch.empty(words, dtype=torch.int64, device=dev)
    eng.sync()  # the partial product is complete before another stream reads it
    for ch in range(eng.P):
        p, wds = vec.vec.device_ptr(ch)
        mine[ch * wds: (ch + 1) * wds].copy_(device_view(p, wds, dev))
    parts = torch.empty(world * words, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(parts, mine, group=group)
    torch.cuda.synchronize()
    total = None
    for r in range(world):
        piece = B200BfvVector(factory, eng.import_raw_ptr(parts.data_ptr() + 8 * r * words, blocks, vec.Dim, vec.Scale, int(vec.Format)))
        if total is None:
            total = piece
        else:
            nxt = total.Add(piece)
            total.Dispose()
            piece.Dispose()
            total = nxt
    return total


def allgather_sparse_elements(factory, vec, counts, group=None):
    """Concatenation of the ranks' sparse encrypted slices (row-sharded layer without ForceDense); counts[r] = elements of rank r."""
    from .he import B200BfvVector
    eng = factory.engine
    world = world_size(group)
    if world == 1:
        return vec
    dev = torch.device("cuda", torch.cuda.current_device())
    cap = max(counts) * eng.ct_words
    mine = torch.zeros(eng.P * cap, dtype=torch.int64, device=dev)
    eng.sync()
    for ch in range(eng.P):
        p, wds = vec.vec.device_ptr(ch)
        mine[ch * cap: ch * cap + wds].copy_(device_view(p, wds, dev))
    parts = torch.empty(world * eng.P * cap, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(parts, mine, group=group)
    total = sum(counts)
    packed = torch.empty(eng.P * total * eng.ct_words, dtype=torch.int64, device=dev)
    for ch in range(eng.P):
        off = 0
        for r in range(world):
            n = counts[r] * eng.ct_words
            src = (r * eng.P + ch) * cap
            packed[ch * total * eng.ct_words + off: ch * total * eng.ct_words + off + n].copy_(parts[src: src + n])
            off += n
    torch.cuda.synchronize()
    return B200BfvVector(factory, eng.import_raw_ptr(packed.data_ptr(), total, total, vec.Scale, int(vec.Format)))


def gpu_numa_node(index):
    """NUMA node of GPU `index` from sysfs (-1 unknown)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = "/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]
        return int(open(path).read())
    except Exception:
        return -1


def bind_to_gpu_numa(index):
    """Pin this process (and what it allocates from now on, first-touch: the pinned staging buffers) to the CPUs of the GPU's NUMA node.
    Returns a description for the bench record (with the previous affinity under "previous_cpus": restore it with os.sched_setaffinity
    before CPU-heavy work such as the CPU baseline leg); a no-op when the topology cannot be read."""
    node = gpu_numa_node(index)
    if node < 0:
        return {"numa_node": None, "bound": False}
    try:
        spec = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return {"numa_node": node, "bound": False}
        previous = sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "bound": True, "cpus": len(allowed), "previous_cpus": previous}
    except Exception:
        return {"numa_node": node, "bound": False}
