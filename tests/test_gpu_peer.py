"""The library's own bitmask exchange over peer memory (fabgpu_peer_mask_*, fabric-mod_b200/sharding.py PeerMaskExchange) on two
or more GPUs, one process per GPU: tampered batches, exact mask on every rank, fused epilogue (all keys tabled) and scatter path
(mixed key-table / generic batch).  Skipped on single-GPU boxes (the gloo tests cover the host logic there)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    import importlib
    import torch
    import torch.distributed as dist
    from oracle import fast
    from tools import workload
    pkg = importlib.import_module("fabric-mod_b200")
    sharding = importlib.import_module("fabric-mod_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    w = workload.Workload(n_total, 12, seed=909, nthreads=4)            # every rank builds the same global batch, verifies its range
    w.tamper_r(0.07)
    b, e = sharding.shard_range(n_total, rank, world)
    ctx = pkg.binding.Context(max_batch=e - b + 64, device_ids=[rank])
    slots = ctx.keys_register(w.keys_xy[:8]) & 0xFFF                     # keys 8..11 stay without a table
    slot_of = np.full(12, -1, np.int32)
    slot_of[:8] = slots
    ex = sharding.PeerMaskExchange(ctx, n_total, world, rank, dev)
    st = torch.cuda.current_stream(dev)
    res = {}
    for name, idx in (("tabled", w.key_idx % 8), ("mixed", w.key_idx)):
        # "tabled": signatures re-keyed onto the 8 tabled keys would not verify, so instead keep the true keys and only pick signatures of tabled keys
        sel = np.arange(b, e)
        ks = slot_of[w.key_idx[sel]]
        if name == "tabled":
            keep = ks >= 0
            r_ = w.r[sel].copy(); r_[~keep] = 0                          # signatures of untabled keys: r = 0 -> invalid at once
            ks = np.where(keep, ks, slots[0]).astype(np.int32)
        else:
            r_ = w.r[sel]
        t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (w.qx()[sel], w.qy()[sel], w.digest[sel], r_, w.s[sel])]
        tk = torch.from_numpy(np.ascontiguousarray(ks)).to(dev)
        for rep in range(3):                                             # several steps: generations and flags are reused
            full = ex.verify(name == "tabled", tk.data_ptr(), t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), t[4].data_ptr(), e - b, st.cuda_stream)
            st.synchronize()
            res[name] = full.cpu().numpy().view(np.uint32).copy()
            dist.barrier()
    np.savez(os.path.join(out_dir, "peer_%d.npz" % rank), **res)
    dist.barrier()
    ex.close()
    ctx.close()
    dist.destroy_process_group()


def test_peer_memory_mask_exchange_two_ranks(tmp_path):
    import torch
    import torch.multiprocessing as mp
    from oracle import fast
    from tools import workload
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    n_total = 20000 + 13                                                 # ragged
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    w = workload.Workload(n_total, 12, seed=909, nthreads=4)
    w.tamper_r(0.07)
    st = fast.verify_batch(w.keys_xy, w.key_idx, w.digest, w.dig_off(), w.sigs, w.sig_off, nthreads=8)
    exp_mixed = fast.valid_mask(st)
    st_t = st.copy(); st_t[w.key_idx >= 8] = 1
    exp_tabled = fast.valid_mask(st_t)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "peer_%d.npz" % r))
        assert (z["mixed"] == exp_mixed).all(), r
        assert (z["tabled"] == exp_tabled).all(), r
    assert 1000 < int(n_total - np.unpackbits(exp_mixed.view(np.uint8)).sum()) < 2000
