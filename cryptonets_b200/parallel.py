"""Multi-GPU plumbing: one process per GPU, replicas over batches, one all-gather of the score ciphertexts.

The 8192 images of a CryptoNets batch share every ciphertext (SIMD slot packing, `CryptoNets/CryptoNets.cs:15-26`), so a batch
cannot be split by image; the shardable unit is the batch.  Each rank therefore owns whole batches and the only
exchange of the path is the gather of the final score ciphertexts (10 x P ciphertexts per batch) at the decrypt/score step.
Works with any torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def batches_of_rank(n_batches, rank, world):
    """Round-robin assignment of batch indices to ranks (weak scaling: every rank gets the same count +-1)."""
    return list(range(rank, n_batches, world))


def gather_score_ciphertexts(local_words, world=None):
    """all-gather a rank's score ciphertexts (1-D int64 tensor of raw words, any device) -> list of tensors, one per rank."""
    world = world or (dist.get_world_size() if dist.is_initialized() else 1)
    if world == 1:
        return [local_words]
    out = [torch.empty_like(local_words) for _ in range(world)]
    dist.all_gather(out, local_words)
    return out


def max_over_ranks(value, device="cpu"):
    """device time of a step is the max over ranks"""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
