"""The verify algorithm the CUDA kernel runs (fabric-mod_b200/csrc/ecdsa_verify.cuh: fixed-base table, Booth
windows, complete Jacobian group law, inversion-free final check) compiled for the host and compared with the oracle.
Runs on the GPU-less build box; the PTX limb primitives themselves are covered by tests/test_gpu_*.py."""
import numpy as np
import pytest

from oracle import bccsp_sw as o
from oracle import fast
from tools import workload
from util import be32, hash_to_e32, hostsim_verify, pkg
import vectors

V_INVALID, V_VALID, V_OFFCURVE = 0, 1, 2


@pytest.mark.parametrize("cached", [False, True, "ba", "small"])
def test_constructed_edge_cases(cached):
    b = pkg().binding
    for c in vectors.build():
        exp = vectors.expected_status(c)
        if c["expect"] is not None:
            assert exp == c["expect"], "oracle disagrees with the construction: " + c["name"]
        if len(c["sig"]) == 0:
            continue
        st, r, s = b.gate_signature(c["sig"])
        if st != o.VALID:
            assert st == exp, c["name"]
            continue
        out = hostsim_verify(be32(c["qx"] % (1 << 256)), be32(c["qy"] % (1 << 256)), hash_to_e32(c["digest"]),
                             np.frombuffer(r, np.uint8), np.frombuffer(s, np.uint8), cached=cached is True, ba=cached == "ba", small=cached == "small")[0]
        got = {V_VALID: o.VALID, V_INVALID: o.INVALID, V_OFFCURVE: o.ERR_OFF_CURVE}[int(out)]
        assert got == exp, c["name"]


def test_config1_1024_tuples():
    # BASELINE.json configs[0]: 1 024 synthetic tuples, K = 16 keys, all valid
    w = workload.Workload(1024, 16, seed=workload.DEFAULT_SEED, nthreads=4)
    exp = fast.verify_batch(w.keys_xy, w.key_idx, w.digest, w.dig_off(), w.sigs, w.sig_off, nthreads=4)
    assert (exp == o.VALID).all()
    out = hostsim_verify(w.qx(), w.qy(), w.digest, w.r, w.s)
    assert (out == V_VALID).all()


def test_tampered_matches_oracle_bit_for_bit():
    w = workload.Workload(768, 8, seed=7, nthreads=4)
    w.tamper_r(frac=0.25)
    exp = fast.verify_batch(w.keys_xy, w.key_idx, w.digest, w.dig_off(), w.sigs, w.sig_off, nthreads=4)
    out = hostsim_verify(w.qx(), w.qy(), w.digest, w.r, w.s)
    assert ((out == V_VALID) == (exp == o.VALID)).all()
    assert 100 < int((exp != o.VALID).sum()) < 300
    # the per-key-table path gives the same bits
    out_c = hostsim_verify(w.qx(), w.qy(), w.digest, w.r, w.s, cached=True)
    assert (out_c == out).all()
    # and so does the batch-affine accumulation with its shared inversions (groups of 64 with a ragged tail: 768 + 37)
    out_b = hostsim_verify(w.qx(), w.qy(), w.digest, w.r, w.s, ba=True)
    assert (out_b == out).all()
    # the small-table tier (signed windows, no doublings)
    out_s = hostsim_verify(w.qx(), w.qy(), w.digest, w.r, w.s, small=True)
    assert (out_s == out).all()
    k = 37
    out_b2 = hostsim_verify(w.qx()[:k], w.qy()[:k], w.digest[:k], w.r[:k], w.s[:k], ba=True)
    assert (out_b2 == out[:k]).all()


def test_field_inversion_by_division_steps():
    """fe_inv_safegcd (the shared inversion of the batch-affine levels) against the Fermat ladder and Python."""
    import ctypes
    from util import hostsim, from_be
    rng = np.random.default_rng(11)
    p = (1 << 256) - (1 << 224) + (1 << 192) + (1 << 96) - 1
    xs = [1, 2, p - 1, p - 2, (1 << 255), 3] + [int.from_bytes(rng.bytes(32), "big") % (p - 1) + 1 for _ in range(60)]
    a = np.concatenate([be32(x) for x in xs])
    outs = []
    for op in (4, 8):
        out = np.zeros_like(a)
        hostsim().hostsim_fieldop(ctypes.c_int(op), a.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(xs)),
                                  out.ctypes.data_as(ctypes.c_void_p))
        outs.append(out.reshape(-1, 32))
    assert (outs[0] == outs[1]).all()
    R = 1 << 256
    for x, o_ in zip(xs, outs[1]):          # Montgomery in / out: x = a R  ->  a^-1 R = R^2 / x
        assert from_be(o_) == (R * R * pow(x, -1, p)) % p


def test_small_table_signed_recoding_on_crafted_scalars():
    """k * Q through the small table (signed FAB_WS-bit windows; the host build uses the product's width) for scalars that stress the recoding: runs of ones (carry through every
    window), digits exactly at the sign boundary (32 / 33), the largest values below 2^256 (the last window absorbs the carry)."""
    import ctypes
    import random
    from oracle import p256
    from util import hostsim
    hs = hostsim()
    Q = p256.scalar_mult(0x1234567, (p256.GX, p256.GY))
    rng = random.Random(99)
    ks = [1, 2, 31, 32, 33, 63, 64, 65, (1 << 256) - 1, (1 << 256) - 2, p256.N - 1, p256.N, p256.N + 1, (1 << 255), (1 << 255) - 1,
          int("100000" * 43, 2) & ((1 << 256) - 1), int("100001" * 43, 2) & ((1 << 256) - 1), int("011111" * 43, 2) & ((1 << 256) - 1),
          int("111111" * 42, 2), 0,
          int("10000000" * 32, 2), int("10000001" * 32, 2), int("01111111" * 32, 2), int("11111111" * 31 + "11111110", 2), int("10000000" * 31 + "01111111", 2)]
    ks += [rng.getrandbits(256) for _ in range(12)]
    for k in ks:
        ox = (ctypes.c_uint8 * 32)(); oy = (ctypes.c_uint8 * 32)()
        hs.hostsim_small_mul(k.to_bytes(32, "big"), Q[0].to_bytes(32, "big"), Q[1].to_bytes(32, "big"), ox, oy)
        exp = p256.scalar_mult(k % p256.N, Q)
        got = (int.from_bytes(bytes(ox), "big"), int.from_bytes(bytes(oy), "big"))
        assert got == ((0, 0) if exp is p256.INF else exp), hex(k)


@pytest.mark.parametrize("ws", [5, 6, 7])
def test_small_tables_at_other_window_widths(ws):
    """The small-table code is width-generic (FAB_WS; the product builds 8): the host build at 5, 6 and 7 bits multiplies crafted scalars
    correctly (signed recoding with the carry absorbed by the last window) and verifies a tampered batch like the generic path."""
    import ctypes
    import os
    import subprocess
    from oracle import p256
    from util import ROOT
    d = os.path.join(ROOT, "tests", "host_sim")
    so = os.path.join(d, "libhostsim_ws%d.so" % ws)
    src = os.path.join(d, "hostsim.cpp")
    hdr = os.path.join(ROOT, "fabric-mod_b200", "csrc", "ecdsa_verify.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-DFAB_WG=8", "-DFAB_WQ=8", "-DFAB_Q_TWO_LEVEL=1", "-DFAB_WS=%d" % ws, "-shared", "-fPIC", "-o", so, src])
    hs = ctypes.CDLL(so)
    Q = p256.scalar_mult(0x7654321, (p256.GX, p256.GY))
    half, full = 1 << (ws - 1), 1 << ws
    nw = (257 + ws - 1) // ws
    pat = lambda digit: sum(digit << (ws * j) for j in range(nw)) & ((1 << 256) - 1)      # the same digit in every window
    ks = [1, half, half + 1, full - 1, full, pat(half), pat(half + 1), pat(half - 1), pat(full - 1), (1 << 256) - 1, p256.N - 1, p256.N + 1, 1 << 255]
    for k in ks:
        ox = (ctypes.c_uint8 * 32)(); oy = (ctypes.c_uint8 * 32)()
        hs.hostsim_small_mul(k.to_bytes(32, "big"), Q[0].to_bytes(32, "big"), Q[1].to_bytes(32, "big"), ox, oy)
        exp = p256.scalar_mult(k % p256.N, Q)
        assert (int.from_bytes(bytes(ox), "big"), int.from_bytes(bytes(oy), "big")) == ((0, 0) if exp is p256.INF else exp), (ws, hex(k))
    w = workload.Workload(256, 4, seed=70 + ws, nthreads=2)
    w.tamper_r(frac=0.25)
    exp = fast.verify_batch(w.keys_xy, w.key_idx, w.digest, w.dig_off(), w.sigs, w.sig_off, nthreads=2)
    arrs = [np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32) for a in (w.qx(), w.qy(), w.digest, w.r, w.s)]
    out = np.zeros(w.n, np.uint8)
    hs.hostsim_verify_batch_small(*[a.ctypes.data_as(ctypes.c_void_p) for a in arrs], ctypes.c_int(w.n), out.ctypes.data_as(ctypes.c_void_p))
    assert ((out == V_VALID) == (exp == o.VALID)).all() and 30 < int((exp != o.VALID).sum()) < 100
