"""N > 1 path on CPU: world_size-2 gloo run of the batch split + bitmask all-gather that bench.py and the
multi-GPU configs use (fabric-mod_b200/sharding.py).  The per-rank verifier here is the oracle (no GPU on this box);
on the GPU box the same plumbing carries the CUDA kernel's mask over NCCL."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fast
from tools import workload

sharding = importlib.import_module("fabric-mod_b200.sharding")


def test_shard_ranges_cover_and_align():
    for n in (0, 1, 31, 32, 33, 1000, 65536, 262144, 1048576, 1048577):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                b, e = sharding.shard_range(n, r, world)
                assert b == prev and b <= e <= n
                assert b % 32 == 0 or b == n
                prev = e
            assert prev == n
            assert sharding.shard_words(n, world) * world * 32 >= n


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = workload.Workload(n, 8, seed=seed, nthreads=2)
    w.tamper_r(0.1)
    b, e = sharding.shard_range(n, rank, world)
    st = fast.verify_batch(w.keys_xy, w.key_idx[b:e], w.digest[b:e], (np.arange(e - b + 1) * 32).astype(np.uint32),
                           w.sigs[w.sig_off[b]:w.sig_off[e]], w.sig_off[b:e + 1] - w.sig_off[b], nthreads=2)
    words = sharding.shard_words(n, world)
    local = np.zeros(words, np.uint32)
    m = fast.valid_mask(st)
    local[: m.shape[0]] = m
    full = sharding.allgather_mask(torch.from_numpy(local.view(np.int32)), n, world)
    np.save(os.path.join(out_dir, "mask_%d.npy" % rank), full.numpy().view(np.uint32))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_allgather_matches_single_process(tmp_path):
    n, seed, world = 5000, 77, 2          # ragged: not a multiple of 32 * world
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, seed, str(tmp_path)), nprocs=world, join=True)
    w = workload.Workload(n, 8, seed=seed, nthreads=2)
    w.tamper_r(0.1)
    exp = fast.valid_mask(fast.verify_batch(w.keys_xy, w.key_idx, w.digest, w.dig_off(), w.sigs, w.sig_off, nthreads=4))
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), "mask_%d.npy" % r))
        assert got.shape == exp.shape and (got == exp).all()
    assert 300 < int(n - np.unpackbits(exp.view(np.uint8)).sum()) < 700


def _parity_worker(rank, world, port, n_total, keys, out_dir):
    """The bench's named-shape parity leg with the oracle standing in for the GPU: shard -> statuses -> mask words -> all-gather."""
    from tools import parity_workload as pw
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = pw.Shard(rank, world, n_total, keys, nthreads=2)
    st = pw.oracle_status(sh, 2)
    words = sharding.shard_words(n_total, world)
    local = np.zeros(words, np.uint32)
    m = fast.valid_mask(st)
    local[: m.shape[0]] = m
    full = sharding.allgather_mask(torch.from_numpy(local.view(np.int32)), n_total, world)
    st_all = torch.empty(n_total, dtype=torch.uint8)
    dist.all_gather_into_tensor(st_all, torch.from_numpy(st))
    if rank == 0:
        np.save(os.path.join(out_dir, "pmask.npy"), full.numpy().view(np.uint32))
        np.save(os.path.join(out_dir, "pstatus.npy"), st_all.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_world2_named_shape_parity_plumbing(tmp_path):
    from tools import parity_workload as pw
    n_total, keys, world = 4096, 8, 2
    port = _free_port()
    mp.spawn(_parity_worker, args=(world, port, n_total, keys, str(tmp_path)), nprocs=world, join=True)
    exp = np.concatenate([pw.oracle_status(pw.Shard(r, world, n_total, keys, nthreads=2), 2) for r in range(world)])
    got_mask = np.load(os.path.join(str(tmp_path), "pmask.npy"))
    got_st = np.load(os.path.join(str(tmp_path), "pstatus.npy"))
    assert (got_st == exp).all() and (got_mask == fast.valid_mask(exp)).all()
    last = pw.Shard(world - 1, world, n_total, keys, nthreads=2)
    assert len(last.tail) > 30 and len(set(int(x) for x in exp)) >= 6          # the adversarial tail brought several error kinds
    assert 100 < int((exp != 0).sum()) < 400
