"""Multi-GPU plumbing for the verify path: one process per GPU (torch.distributed), the signature batch split into
contiguous 32-aligned ranges, and the validity bitmask words reassembled on every rank.  Signatures are independent units
(SURVEY.md section 8e): the bitmask is the only thing that crosses GPUs.  Two ways to move it:

  * allgather_mask: one all_gather_into_tensor (NCCL over NVLink on GPUs, gloo in the CPU tests);
  * PeerMaskExchange: the library's own exchange over peer memory (include/fabgpu_ecdsa.h, fabgpu_peer_mask_*): the verify
    kernel's epilogue stores every ballot word into the buffer of every rank (P2P writes over NVLink / NVSwitch) and a wait
    kernel on the same stream returns when all ranks' words have landed -- no collective launch on the data path.  torch is
    used only to hand the CUDA IPC handles around once, at set-up."""
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


class _DeviceWords:
    """__cuda_array_interface__ view of `n` int32 words at a raw device pointer (what the C library hands back)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


class PeerMaskExchange:
    """Set-up and per-step driver of fabgpu_verify_p256_device_keyed_allgather for this rank's context `ctx`."""

    def __init__(self, ctx, n_total: int, world: int, rank: int, device, group=None):
        """Collective over `group`.  self.ok is False on EVERY rank when any rank could not create or map a buffer (no peer access
        between some pair of GPUs, IPC disabled ...): the caller then keeps the NCCL all-gather."""
        self.ctx, self.world, self.rank, self.device = ctx, world, rank, device
        self.n_total = n_total
        self.words = shard_words(n_total, world)
        self.step = 0
        self.error = None

        def agree(flag: bool) -> bool:
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(int(t.item()))

        mine = torch.zeros(64, dtype=torch.uint8, device=device)
        try:
            mine = torch.from_numpy(ctx.peer_mask_create(world, rank, self.words)).to(device)
            created = True
        except Exception as e:                                   # noqa: BLE001 -- any failure means "use NCCL"
            created, self.error = False, str(e)
        self.ok = agree(created)
        if not self.ok:
            if created:
                ctx.peer_mask_close()
            return
        handles = torch.empty(64 * world, dtype=torch.uint8, device=device)
        if world == 1:
            handles.copy_(mine)
        else:
            dist.all_gather_into_tensor(handles, mine, group=group)
        try:
            ctx.peer_mask_open(handles.cpu().numpy())
            opened = True
        except Exception as e:                                   # noqa: BLE001
            opened, self.error = False, str(e)
        self.ok = agree(opened)                                  # also the barrier: every rank has mapped every buffer before the first store
        if not self.ok:
            ctx.peer_mask_close()

    def verify(self, all_cached, d_key_slot, d_qx, d_qy, d_e, d_r, d_s, n, stream=0) -> torch.Tensor:
        """Enqueues verify + exchange on `stream`; returns the assembled mask (int32[ceil(n_total/32)], device memory owned by the
        library, valid after the stream has run and until two more steps have been enqueued)."""
        self.step += 1
        ptr = self.ctx.verify_p256_device_keyed_allgather(all_cached, d_key_slot, d_qx, d_qy, d_e, d_r, d_s, n, self.step, stream)
        full = torch.as_tensor(_DeviceWords(ptr, self.words * self.world), device=self.device)
        return full[: (self.n_total + 31) // 32]

    def close(self):
        self.ctx.peer_mask_close()
