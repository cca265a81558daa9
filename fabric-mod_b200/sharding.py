"""Multi-GPU plumbing for the verify path: one process per GPU (torch.distributed), the signature batch split into
contiguous 32-aligned ranges, one all-gather of the validity bitmask words (NCCL over NVLink on GPUs, gloo in the
CPU tests).  There is no other data-path collective: signatures are independent units."""
import torch
import torch.distributed as dist


def shard_words(n_total: int, world: int) -> int:
    """Mask words per rank: ranges are padded so that no uint32 mask word straddles two ranks."""
    words = (n_total + 31) // 32
    return (words + world - 1) // world


def shard_range(n_total: int, rank: int, world: int):
    """[begin, end) of the signatures rank `rank` verifies."""
    per = shard_words(n_total, world) * 32
    return min(n_total, rank * per), min(n_total, (rank + 1) * per)


def allgather_mask(local_words: torch.Tensor, n_total: int, world: int, group=None) -> torch.Tensor:
    """local_words: int32[shard_words] (this rank's mask, zero-padded).  Returns int32[ceil(n_total/32)] on every rank."""
    per = shard_words(n_total, world)
    assert local_words.numel() == per, (local_words.numel(), per)
    full = torch.empty(per * world, dtype=local_words.dtype, device=local_words.device)
    if world == 1:
        full.copy_(local_words)
    else:
        dist.all_gather_into_tensor(full, local_words.contiguous(), group=group)
    return full[: (n_total + 31) // 32]
