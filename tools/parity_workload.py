"""Parity workloads of BASELINE.json configs[3] / configs[4] at their named multi-GPU shape (not product code).

The global batch of `n_total` signatures is split into `world` contiguous 32-aligned ranges (fabric-mod_b200/sharding.py).  Every
rank builds ITS shard from the seed (so no signature bytes cross ranks); rank 0 can rebuild any shard for the oracle.  Each
shard: seeded low-S signatures over K keys, a Bernoulli(5 %) subset with one bit of r flipped, and --
on the last rank -- the adversarial tail of tests/vectors.py (high-S, r >= n, r = 0, s = 0, e = 0, u1 G = +-u2 Q, x(R) >= n,
off-curve keys, malformed DER ...) written over the shard's last entries.
"""
import os
import sys

import numpy as np

from tools import workload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARITY_SEED = 0xC0F1655


def named_shape(world):
    """(n_total, K, name) for `world` GPUs: configs[4] on 4 GPUs, configs[3] on 8, the 64k-per-GPU analogue otherwise."""
    if world == 8:
        return 1048576, 256, "configs[3]: 1 048 576 signatures over 8 GPUs (K = 256 keys)"
    if world == 4:
        return 262144, 64, "configs[4]: 262 144 signatures over 4 GPUs (K = 64 keys)"
    return 65536 * world, 64, "configs[1] shape: 65 536 signatures per GPU over %d GPU(s) (K = 64 keys)" % world


def _tail_cases():
    tdir = os.path.join(ROOT, "tests")
    if tdir not in sys.path:
        sys.path.insert(0, tdir)
    import vectors
    cases = [c for c in vectors.build()]
    return cases, vectors


class Shard:
    """Arrays in the layout fabgpu_bccsp_verify_batch takes: keys_xy uint8[K',64], key_idx int32[n], digest blob + uint32[n+1]
    offsets, DER blob + uint32[n+1] offsets; `tampered` = indices whose r was flipped; `tail` = names of the tail cases."""

    def __init__(self, rank, world, n_total, K, nthreads=None, with_tail=True):
        per = ((n_total + 31) // 32 + world - 1) // world * 32
        begin, end = min(n_total, rank * per), min(n_total, (rank + 1) * per)
        n = end - begin
        w = workload.Workload(n, K, seed=PARITY_SEED + 7919 * rank, nthreads=nthreads)
        self.tampered = w.tamper_r(0.05)
        self.n, self.begin = n, begin
        keys = [w.keys_xy]
        key_idx = w.key_idx.copy()
        dig_blob, dig_off = w.digest.reshape(-1), w.dig_off()
        sig_blob, sig_off = w.sigs, w.sig_off
        self.tail, self.tail_expected = [], []
        if with_tail and rank == world - 1:
            cases, vectors = _tail_cases()
            m = min(len(cases), n)
            cases = cases[:m]
            extra = np.zeros((m, 64), np.uint8)
            for j, c in enumerate(cases):
                extra[j, :32] = np.frombuffer((c["qx"] % (1 << 256)).to_bytes(32, "big"), np.uint8)
                extra[j, 32:] = np.frombuffer((c["qy"] % (1 << 256)).to_bytes(32, "big"), np.uint8)
            keys.append(extra)
            head = n - m
            key_idx[head:] = K + np.arange(m, dtype=np.int32)
            tail_dig = b"".join(c["digest"] for c in cases)
            tail_sig = b"".join(c["sig"] for c in cases)
            dig_blob = np.concatenate([dig_blob[: 32 * head], np.frombuffer(tail_dig, np.uint8)])
            dl = np.array([len(c["digest"]) for c in cases], np.uint32)
            dig_off = np.concatenate([dig_off[: head + 1], (32 * head + np.cumsum(dl)).astype(np.uint32)])
            sl = np.array([len(c["sig"]) for c in cases], np.uint32)
            sig_head = int(sig_off[head])
            sig_blob = np.concatenate([sig_blob[:sig_head], np.frombuffer(tail_sig, np.uint8)])
            sig_off = np.concatenate([sig_off[: head + 1], (sig_head + np.cumsum(sl)).astype(np.uint32)])
            self.tail = [c["name"] for c in cases]
            self.tail_expected = [vectors.expected_status(c) for c in cases]       # the Python oracle's verdict on each tail case
            self.tampered = self.tampered[self.tampered < head]
        self.keys_xy = np.ascontiguousarray(np.concatenate(keys))
        self.key_idx = np.ascontiguousarray(key_idx)
        self.dig_blob = np.ascontiguousarray(dig_blob, dtype=np.uint8)
        self.dig_off = np.ascontiguousarray(dig_off, dtype=np.uint32)
        self.sig_blob = np.ascontiguousarray(sig_blob, dtype=np.uint8)
        self.sig_off = np.ascontiguousarray(sig_off, dtype=np.uint32)

    def args(self):
        return (self.keys_xy, self.key_idx, self.dig_blob, self.dig_off, self.sig_blob, self.sig_off)


def oracle_status(shard, nthreads):
    """ORACLE side (tests / bench checker only): the C port's three-valued status for every signature of the shard; the tail is
    cross-checked against the Python restatement."""
    from oracle import fast
    st = fast.verify_batch(*shard.args(), nthreads=nthreads)
    if shard.tail:
        m = len(shard.tail)
        got = st[shard.n - m:].tolist()
        assert got == list(shard.tail_expected), "C and Python oracles disagree on the adversarial tail: %r" % (
            [(nm, a, b) for nm, a, b in zip(shard.tail, got, shard.tail_expected) if a != b],)
    return st
