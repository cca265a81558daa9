"""ORACLE (test infrastructure, NOT product code).

Pure-Python big-integer restatement of the arithmetic the reference reaches at
``bccsp/sw/ecdsa.go:56`` -- Go 1.14.4 ``crypto/ecdsa.Verify`` over
``crypto/elliptic`` P-256 (Go stdlib is not vendored in /root/reference; the
version is pinned at reference ``Makefile:79`` / ``go.mod:3``).  The published
algorithm (FIPS 186-4 section 6.4 as implemented by Go 1.14) is restated in
``ecdsa_verify_go114`` below; see SURVEY.md section 3.2 for the step list.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  Everything here is written for
obviousness (affine coordinates, modular inverse by ``pow``), not speed.
"""

# --- NIST P-256 domain parameters (checked against OpenSSL in tests) ---
P = 0xFFFFFFFF00000001000000000000000000000000FFFFFFFFFFFFFFFFFFFFFFFF
A = P - 3
B = 0x5AC635D8AA3A93E7B3EBBD55769886BC651D06B0CC53B0F63BCE3C3E27D2604B
GX = 0x6B17D1F2E12C4247F8BCE6E563A440F277037D812DEB33A0F4A13945D898C296
GY = 0x4FE342E2FE1A7F9B8EE7EB4A7C0F9E162BCE33576B315ECECBB6406837BF51F5
N = 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551
HALF_N = N >> 1  # bccsp/utils/ecdsa.go:27-32 (curveHalfOrders = N >> 1)

INF = None  # point at infinity


def is_on_curve(x, y):
    if not (0 <= x < P and 0 <= y < P):
        return False
    return (y * y - (x * x * x + A * x + B)) % P == 0


def point_add(p1, p2):
    """Complete affine addition (handles infinity, P == Q, P == -Q)."""
    if p1 is INF:
        return p2
    if p2 is INF:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return INF
        lam = (3 * x1 * x1 + A) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    y3 = (lam * (x1 - x3) - y1) % P
    return (x3, y3)


def point_neg(p):
    if p is INF:
        return INF
    return (p[0], (-p[1]) % P)


def scalar_mult(k, pt):
    """Left-to-right double-and-add; k any non-negative integer."""
    acc = INF
    for i in range(k.bit_length() - 1, -1, -1):
        acc = point_add(acc, acc)
        if (k >> i) & 1:
            acc = point_add(acc, pt)
    return acc


def hash_to_int(digest: bytes) -> int:
    """Go 1.14 crypto/ecdsa hashToInt for a 256-bit order: leftmost min(len,32)
    bytes as a big-endian integer; no reduction mod N (orderBits == 256 so the
    'excess' shift is zero for <= 32 bytes)."""
    if len(digest) > 32:
        digest = digest[:32]
    return int.from_bytes(digest, "big")


def ecdsa_verify_go114(qx: int, qy: int, digest: bytes, r: int, s: int) -> bool:
    """Go 1.14 crypto/ecdsa.Verify(pub, hash, r, s) restated (reached from the
    reference at bccsp/sw/ecdsa.go:56 and bccsp/pkcs11/ecdsa.go:47).

    Steps: r,s in [1,N-1] else false; e = hashToInt; w = s^-1 mod N;
    u1 = e*w mod N; u2 = r*w mod N; (x,y) = u1*G + u2*Q with complete addition;
    infinity -> false; accept iff x mod N == r.  No on-curve check of Q (Go 1.14
    has none in Verify); callers of this oracle must keep Q on the curve, because
    off-curve behaviour of Go's assembly is formula-specific and not restated.
    """
    if r <= 0 or s <= 0:
        return False
    if r >= N or s >= N:
        return False
    e = hash_to_int(digest)
    w = pow(s, -1, N)
    u1 = e * w % N
    u2 = r * w % N
    pt = point_add(scalar_mult(u1, (GX, GY)), scalar_mult(u2, (qx, qy)))
    if pt is INF:
        return False
    return pt[0] % N == r


def ecdsa_sign_lows(d: int, digest: bytes, k: int):
    """Test-vector signer: plain ECDSA with caller-supplied nonce k, then the low-S
    normalisation signECDSA applies (bccsp/sw/ecdsa.go:27-39 -> utils.ToLowS
    bccsp/utils/ecdsa.go:94-109).  Returns (r, s) or None if k is unusable."""
    k %= N
    if k == 0:
        return None
    pt = scalar_mult(k, (GX, GY))
    r = pt[0] % N
    if r == 0:
        return None
    e = hash_to_int(digest)
    s = pow(k, -1, N) * (e + r * d) % N
    if s == 0:
        return None
    if s > HALF_N:
        s = N - s
    return r, s
