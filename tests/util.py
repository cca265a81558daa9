import ctypes
import importlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pkg():
    return importlib.import_module("fabric-mod_b200")


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def be32(x: int) -> np.ndarray:
    return np.frombuffer(int(x).to_bytes(32, "big"), np.uint8).copy()


def from_be(a) -> int:
    return int.from_bytes(bytes(a), "big")


def hash_to_e32(digest: bytes) -> np.ndarray:
    d = digest[:32]
    return np.frombuffer(b"\x00" * (32 - len(d)) + d, np.uint8).copy()


_HS = None


def hostsim():
    """tests/host_sim/libhostsim.so: the device verify code compiled for the host (test helper, not product)."""
    global _HS
    if _HS is None:
        d = os.path.join(ROOT, "tests", "host_sim")
        so = os.path.join(d, "libhostsim.so")
        src = os.path.join(d, "hostsim.cpp")
        hdrs = [os.path.join(ROOT, "fabric-mod_b200", "csrc", h) for h in ("p256_fe.cuh", "p256_point.cuh", "p256_modinv.cuh", "ecdsa_verify.cuh", "ecdsa_batchaffine.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(f) for f in [src] + hdrs):
            # host build of the table would take minutes at the product's 16-bit G windows; the algorithm is window-size generic
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-DFAB_WG=8", "-DFAB_WQ=8", "-DFAB_Q_TWO_LEVEL=1", "-shared", "-fPIC", "-o", so, src])
        _HS = ctypes.CDLL(so)
        _HS.hostsim_gtable.restype = ctypes.c_size_t
    return _HS


def hostsim_verify(qx, qy, e, r, s, cached=False, ba=False, small=False):
    arrs = [np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32) for a in (qx, qy, e, r, s)]
    n = arrs[0].shape[0]
    out = np.zeros(n, np.uint8)
    fn = hostsim().hostsim_verify_batch_small if small else hostsim().hostsim_verify_batch_ba if ba else (hostsim().hostsim_verify_batch_cached if cached else hostsim().hostsim_verify_batch)
    fn(*[a.ctypes.data_as(ctypes.c_void_p) for a in arrs], ctypes.c_int(n), out.ctypes.data_as(ctypes.c_void_p))
    return out


def mask_bits(mask_words, n):
    return np.unpackbits(np.asarray(mask_words, dtype="<u4").view(np.uint8), bitorder="little")[:n]
