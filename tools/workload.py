"""Synthetic workload generator for the BASELINE.json configs.  Not the oracle and
not product code: it only manufactures (pubkey, digest, DER signature) tuples the way Fabric signers do
(reference bccsp/sw/ecdsa.go:27-39) using tools/libfabgpu_siggen.so (OpenSSL arithmetic, seeded DRBG)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
DEFAULT_SEED = 0xFAB51C


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libfabgpu_siggen.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.siggen_der.restype = ctypes.c_long
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Workload:
    """n signatures over K keys.  Arrays: keys_xy uint8[K,64], priv uint8[K,32], key_idx int32[n],
    digest uint8[n,32], r/s uint8[n,32] (big-endian), sigs uint8 blob + sig_off uint32[n+1] (DER)."""

    def __init__(self, n, K, seed=DEFAULT_SEED, msg_len=1024, nthreads=None, key_idx=None):
        nthreads = nthreads or min(os.cpu_count() or 1, 64)
        L = lib()
        self.n, self.K, self.seed = n, K, seed
        self.priv = np.zeros((K, 32), np.uint8)
        self.keys_xy = np.zeros((K, 64), np.uint8)
        L.siggen_keys(ctypes.c_uint64(seed), ctypes.c_int(K), _p(self.priv), _p(self.keys_xy))
        rng = np.random.default_rng(seed)
        self.key_idx = rng.integers(0, K, size=n, dtype=np.int32) if key_idx is None else np.ascontiguousarray(key_idx, dtype=np.int32)
        self.digest = np.zeros((n, 32), np.uint8)
        L.siggen_digests(ctypes.c_uint64(seed), ctypes.c_int(n), ctypes.c_int(msg_len), _p(self.digest))
        self.r = np.zeros((n, 32), np.uint8)
        self.s = np.zeros((n, 32), np.uint8)
        L.siggen_sign_batch(_p(self.priv), _p(self.key_idx), _p(self.digest), ctypes.c_int(n), ctypes.c_uint64(seed),
                            _p(self.r), _p(self.s), ctypes.c_int(nthreads))
        self._der()

    def _der(self):
        L = lib()
        self.sig_off = np.zeros(self.n + 1, np.uint32)
        blob = np.zeros(72 * self.n + 8, np.uint8)
        tot = L.siggen_der(_p(self.r), _p(self.s), ctypes.c_int(self.n), _p(blob), _p(self.sig_off))
        self.sigs = blob[:tot].copy()

    def tamper_r(self, frac=0.05, seed=None):
        """Config #5: for a seeded Bernoulli(frac) subset flip one seeded bit of r (kept non-zero), re-DER."""
        rng = np.random.default_rng(self.seed + 5 if seed is None else seed)
        pick = np.nonzero(rng.random(self.n) < frac)[0]
        byte = rng.integers(1, 32, size=pick.size)      # never the top byte: keeps r < 2^248.. fine vs n is not guaranteed; oracle decides
        bit = rng.integers(0, 8, size=pick.size)
        self.r[pick, byte] ^= (1 << bit).astype(np.uint8)
        zero = np.nonzero(~self.r.any(axis=1))[0]
        self.r[zero, 31] = 1
        self._der()
        return pick

    def qx(self):
        return np.ascontiguousarray(self.keys_xy[self.key_idx, :32])

    def qy(self):
        return np.ascontiguousarray(self.keys_xy[self.key_idx, 32:])

    def dig_off(self):
        return (np.arange(self.n + 1, dtype=np.uint32) * 32).astype(np.uint32)

    def sig(self, i):
        return bytes(self.sigs[self.sig_off[i]:self.sig_off[i + 1]])
