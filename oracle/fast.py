"""ORACLE (test infrastructure, NOT product code): ctypes binding of oracle/c/libecdsa_oracle.so,
the fast C restatement of bccsp/sw's verifier (see the header of oracle/c/ecdsa_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "c", "libecdsa_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_verify_batch.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def verify_batch(keys_xy, key_idx, digests, dig_off, sigs, sig_off, nthreads=1):
    """keys_xy uint8[K,64]; key_idx int32[n]; digests/sigs uint8 blobs with uint32[n+1] offsets.
    Returns uint8[n] status codes (oracle.bccsp_sw.VALID, INVALID, ERR_*)."""
    keys_xy = np.ascontiguousarray(keys_xy, dtype=np.uint8).reshape(-1, 64)
    key_idx = np.ascontiguousarray(key_idx, dtype=np.int32)
    digests = np.ascontiguousarray(digests, dtype=np.uint8).reshape(-1)
    sigs = np.ascontiguousarray(sigs, dtype=np.uint8).reshape(-1)
    dig_off = np.ascontiguousarray(dig_off, dtype=np.uint32)
    sig_off = np.ascontiguousarray(sig_off, dtype=np.uint32)
    n = key_idx.shape[0]
    assert dig_off.shape[0] == n + 1 and sig_off.shape[0] == n + 1
    status = np.full(n, 255, dtype=np.uint8)
    if digests.size == 0:
        digests = np.zeros(1, np.uint8)
    if sigs.size == 0:
        sigs = np.zeros(1, np.uint8)
    rc = lib().oracle_verify_batch(_p(keys_xy), ctypes.c_int(keys_xy.shape[0]), _p(key_idx), _p(digests), _p(dig_off),
                                   _p(sigs), _p(sig_off), ctypes.c_int(n), _p(status), ctypes.c_int(nthreads))
    if rc != 0:
        raise RuntimeError("oracle_verify_batch failed")
    return status


def valid_mask(status):
    """uint32 little-endian bitmask words: signature i -> bit i%32 of word i//32, 1 = VALID."""
    bits = (np.asarray(status) == 0).astype(np.uint8)
    pad = (-len(bits)) % 32
    if pad:
        bits = np.concatenate([bits, np.zeros(pad, np.uint8)])
    return np.packbits(bits.reshape(-1, 32), axis=1, bitorder="little").view("<u4").reshape(-1)
