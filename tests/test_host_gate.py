"""Product host gates (fabgpu_gate_signature: DER per Go encoding/asn1, R>0, S>0, low-S, r < 2^256) against the
oracle's restatement of bccsp/utils/ecdsa.go:43-92 + bccsp/sw/ecdsa.go:42-54.  No GPU needed."""
import random

from oracle import bccsp_sw as o
from oracle import goasn1, p256
from util import pkg
import vectors


def oracle_gate(sig):
    """What the gates alone decide: oracle status with the curve check replaced by 'ask the GPU' (VALID) / r >= 2^256."""
    if len(sig) == 0:
        return o.ERR_UNMARSHAL      # fabgpu_gate_signature sees the raw parser outcome; the empty-sig gate sits above it
    try:
        r, s = goasn1.unmarshal_ecdsa_signature(sig)
    except goasn1.Asn1Error:
        return o.ERR_UNMARSHAL
    if r <= 0:
        return o.ERR_R_NOT_POSITIVE
    if s <= 0:
        return o.ERR_S_NOT_POSITIVE
    if s > p256.HALF_N:
        return o.ERR_HIGH_S
    if r >= 1 << 256:
        return o.INVALID
    return o.VALID


def test_gate_on_constructed_cases():
    b = pkg().binding
    for c in vectors.build():
        st, r, s = b.gate_signature(c["sig"])
        assert st == oracle_gate(c["sig"]), c["name"]
        if st == o.VALID:
            rr, ss = goasn1.unmarshal_ecdsa_signature(c["sig"])
            assert int.from_bytes(r, "big") == rr and int.from_bytes(s, "big") == ss


def test_gate_fuzz_mutations():
    b = pkg().binding
    rnd = random.Random(20260922)
    seeds = [goasn1.marshal_ecdsa_signature(rnd.getrandbits(256), rnd.getrandbits(255)) for _ in range(40)]
    seeds += [goasn1.marshal_ecdsa_signature(rnd.getrandbits(8 * k), rnd.getrandbits(8 * j)) for k in (1, 2, 31, 33, 40) for j in (1, 31, 32)]
    seeds += [goasn1.marshal_ecdsa_signature(-rnd.getrandbits(200), 5), goasn1.marshal_ecdsa_signature(5, -rnd.getrandbits(200))]
    n = 0
    for sd in seeds:
        for _ in range(150):
            m = bytearray(sd)
            for _ in range(rnd.choice((1, 1, 2, 3))):
                op = rnd.random()
                if op < 0.6 and m:
                    m[rnd.randrange(len(m))] = rnd.getrandbits(8)
                elif op < 0.8 and m:
                    del m[rnd.randrange(len(m))]
                else:
                    m.insert(rnd.randrange(len(m) + 1), rnd.getrandbits(8))
            sig = bytes(m)
            st, r, s = b.gate_signature(sig)
            assert st == oracle_gate(sig), sig.hex()
            n += 1
    assert n > 5000


def test_gate_long_lengths():
    b = pkg().binding
    big = goasn1.marshal_ecdsa_signature((1 << 2000) + 1, 7)      # long-form lengths (0x82 ..)
    assert b.gate_signature(big)[0] == o.INVALID == oracle_gate(big)
    assert b.gate_signature(goasn1.marshal_ecdsa_signature(7, (1 << 2000) + 1))[0] == o.ERR_HIGH_S
