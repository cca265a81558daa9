"""Pins the oracle (oracle/*.py, oracle/c) before anything else is compared with it:
golden X.509 fixture signatures from the reference tree, the reference's malformed-DER vectors,
its boundary constructions, and agreement with OpenSSL (python `cryptography`)."""
import json
import os

import numpy as np
import pytest

from oracle import bccsp_sw as o
from oracle import fast, goasn1, p256
from tools import workload

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "x509_fixtures.json")
FIX = json.load(open(GOLDEN))
EXP = {"VALID": o.VALID, "ERR_HIGH_S": o.ERR_HIGH_S}

# reference bccsp/sw/impl_test.go:931-959 (same list in bccsp/pkcs11/pkcs11_test.go:875-911)
MALFORMED_DER = [
    bytes([0x30, 0x07, 0x02, 0x01, 0x8F, 0x02, 0x02, 0xFF, 0xF1]),
    bytes([0x30, 0x07, 0x02, 0x01, 0x8F, 0x02, 0x02, 0x00, 0x01]),
    bytes([0x30, 0x07, 0x02, 0x01, 0x8F, 0x02, 0x81, 0x01, 0x01]),
    bytes([0x30, 0x07, 0x02, 0x01, 0x8F, 0x02, 0x81, 0x01, 0x8F]),
    bytes([0x30, 0x0A, 0x02, 0x01, 0x8F, 0x02, 0x05, 0x00, 0x00, 0x00, 0x00, 0x8F]),
]


def _c_status(items):
    """items: list of (qx:int, qy:int, digest:bytes, sig:bytes) -> C-oracle status codes."""
    keys = np.zeros((len(items), 64), np.uint8)
    dig, sig, doff, soff = [], [], [0], [0]
    for i, (qx, qy, d, s) in enumerate(items):
        keys[i] = np.frombuffer(qx.to_bytes(32, "big") + qy.to_bytes(32, "big"), np.uint8)
        dig.append(d); sig.append(s)
        doff.append(doff[-1] + len(d)); soff.append(soff[-1] + len(s))
    return fast.verify_batch(keys, np.arange(len(items), dtype=np.int32), np.frombuffer(b"".join(dig), np.uint8),
                             np.array(doff, np.uint32), np.frombuffer(b"".join(sig), np.uint8), np.array(soff, np.uint32))


def test_constants():
    assert p256.is_on_curve(p256.GX, p256.GY)
    assert p256.scalar_mult(p256.N, (p256.GX, p256.GY)) is p256.INF
    assert p256.scalar_mult(p256.N - 1, (p256.GX, p256.GY)) == (p256.GX, p256.P - p256.GY)
    assert p256.HALF_N == 0x7FFFFFFF800000007FFFFFFFFFFFFFFFDE737D56D38BCF4279DCE5617E3192A8
    assert p256.P - p256.N == 0x4319055358E8617B0C46353D039CDAAE


def test_x509_fixtures_python_oracle():
    assert len(FIX) >= 96
    for x in FIX:
        k = o.P256PublicKey(int(x["qx"], 16), int(x["qy"], 16))
        st = o.status(k, bytes.fromhex(x["sig_der"]), bytes.fromhex(x["digest"]))
        assert st == EXP[x["expect"]], x["cert"]
        valid, err = o.csp_verify(k, bytes.fromhex(x["sig_der"]), bytes.fromhex(x["digest"]))
        if x["expect"] == "VALID":
            assert valid and err is None
        else:
            assert not valid and "Invalid S. Must be smaller than half the order [" in err


def test_x509_fixtures_c_oracle():
    items = [(int(x["qx"], 16), int(x["qy"], 16), bytes.fromhex(x["digest"]), bytes.fromhex(x["sig_der"])) for x in FIX]
    st = _c_status(items)
    assert list(st) == [EXP[x["expect"]] for x in FIX]


def test_high_s_fixtures_become_valid_after_to_low_s():
    # msp/cert.go:76-116 sanitizeECDSASignedCert / msp/cert_test.go:70-97: s -> N - s must verify
    items = []
    for x in FIX:
        if x["expect"] != "ERR_HIGH_S":
            continue
        r, s = int(x["r"], 16), int(x["s"], 16)
        sig = goasn1.marshal_ecdsa_signature(r, p256.N - s)
        k = o.P256PublicKey(int(x["qx"], 16), int(x["qy"], 16))
        assert o.status(k, sig, bytes.fromhex(x["digest"])) == o.VALID
        items.append((k.x, k.y, bytes.fromhex(x["digest"]), sig))
    assert len(items) >= 31
    assert set(_c_status(items)) == {o.VALID}


def test_malformed_der_vectors():
    k = o.P256PublicKey(p256.GX, p256.GY)
    for v in MALFORMED_DER:
        with pytest.raises(goasn1.Asn1Error):
            goasn1.unmarshal_ecdsa_signature(v)
        assert o.status(k, v, b"\x01" * 32) == o.ERR_UNMARSHAL
        valid, err = o.csp_verify(k, v, b"\x01" * 32)
        assert not valid and "Failed unmashalling signature [" in err
    assert set(_c_status([(p256.GX, p256.GY, b"\x01" * 32, v) for v in MALFORMED_DER])) == {o.ERR_UNMARSHAL}


def test_unmarshal_boundaries():
    # bccsp/utils/ecdsa_test.go:19-62
    for raw in (None, b"", b"\x00"):
        r, s, err = o.unmarshal_ecdsa_signature(raw)
        assert err is not None and "failed unmashalling signature [" in err
    for (r, s, msg) in ((-1, 1, "R must be larger than zero"), (0, 1, "R must be larger than zero"),
                        (1, 0, "S must be larger than zero"), (1, -1, "S must be larger than zero")):
        _, _, err = o.unmarshal_ecdsa_signature(goasn1.marshal_ecdsa_signature(r, s))
        assert err == "invalid signature, " + msg
    assert o.unmarshal_ecdsa_signature(goasn1.marshal_ecdsa_signature(1, 1)) == (1, 1, None)
    items = [(p256.GX, p256.GY, b"\x01" * 32, goasn1.marshal_ecdsa_signature(r, s))
             for (r, s) in ((-1, 1), (0, 1), (1, 0), (1, -1), (1, 1))]
    assert list(_c_status(items)) == [o.ERR_R_NOT_POSITIVE, o.ERR_R_NOT_POSITIVE, o.ERR_S_NOT_POSITIVE,
                                      o.ERR_S_NOT_POSITIVE, o.INVALID]


def test_low_s_boundary():
    # bccsp/utils/ecdsa_test.go:64-88, bccsp/sw/ecdsa_test.go:62-73
    assert o.is_low_s(0) and o.is_low_s(p256.HALF_N) and not o.is_low_s(p256.HALF_N + 1)
    k = o.P256PublicKey(p256.GX, p256.GY)
    valid, err = o.verify_ecdsa(k, goasn1.marshal_ecdsa_signature(1, p256.HALF_N + 1), b"hello world")
    assert not valid and "Invalid S. Must be smaller than half the order [" in err
    _, err = o.verify_ecdsa(k, None, b"hello world")
    assert "Failed unmashalling signature [" in err


def test_csp_argument_gates():
    k = o.P256PublicKey(p256.GX, p256.GY)
    assert o.csp_verify(None, b"x", b"y") == (False, "Invalid Key. It must not be nil.")
    assert o.csp_verify(k, b"", b"y") == (False, "Invalid signature. Cannot be empty.")
    assert o.csp_verify(k, b"x", b"") == (False, "Invalid digest. Cannot be empty.")
    assert "Unsupported 'VerifyKey' provided [" in o.csp_verify(object(), b"x", b"y")[1]


def test_der_leniency_go():
    # trailing bytes after/inside the SEQUENCE accepted; long r parses then INVALID
    good = goasn1.marshal_ecdsa_signature(5, 7)
    assert goasn1.unmarshal_ecdsa_signature(good + b"\xde\xad") == (5, 7)
    inner = good[2:] + b"\x02\x01\x09"
    assert goasn1.unmarshal_ecdsa_signature(b"\x30" + bytes([len(inner)]) + inner) == (5, 7)
    big_r = (1 << 300) + 12345
    sig = goasn1.marshal_ecdsa_signature(big_r, 7)
    assert goasn1.unmarshal_ecdsa_signature(sig) == (big_r, 7)
    k = o.P256PublicKey(p256.GX, p256.GY)
    assert o.status(k, sig, b"\x01" * 32) == o.INVALID
    assert list(_c_status([(k.x, k.y, b"\x01" * 32, sig), (k.x, k.y, b"\x01" * 32, good + b"\xde\xad")])) == [o.INVALID, o.INVALID]
    # r == N and r == N-1 (well-formed, low s): (false, nil)
    for r in (p256.N, p256.N + 5):
        assert o.status(k, goasn1.marshal_ecdsa_signature(r, 7), b"\x01" * 32) == o.INVALID


def test_sign_verify_roundtrip_three_way():
    """siggen (OpenSSL signer) -> python oracle, C oracle, and python `cryptography` all agree."""
    from cryptography.hazmat.primitives.asymmetric import ec, utils as au
    from cryptography.hazmat.primitives import hashes
    from cryptography.exceptions import InvalidSignature
    w = workload.Workload(48, 4, seed=1234, msg_len=64, nthreads=1)
    st_c = fast.verify_batch(w.keys_xy, w.key_idx, w.digest, w.dig_off(), w.sigs, w.sig_off)
    assert set(st_c) == {o.VALID}
    for i in range(w.n):
        kx = int.from_bytes(bytes(w.keys_xy[w.key_idx[i], :32]), "big")
        ky = int.from_bytes(bytes(w.keys_xy[w.key_idx[i], 32:]), "big")
        assert o.status(o.P256PublicKey(kx, ky), w.sig(i), bytes(w.digest[i])) == o.VALID
        pub = ec.EllipticCurvePublicNumbers(kx, ky, ec.SECP256R1()).public_key()
        pub.verify(w.sig(i), bytes(w.digest[i]), ec.ECDSA(au.Prehashed(hashes.SHA256())))
    # tamper: python and C oracles agree bit for bit
    w.tamper_r(frac=0.3, seed=99)
    st_c = fast.verify_batch(w.keys_xy, w.key_idx, w.digest, w.dig_off(), w.sigs, w.sig_off)
    st_p = []
    for i in range(w.n):
        kx = int.from_bytes(bytes(w.keys_xy[w.key_idx[i], :32]), "big")
        ky = int.from_bytes(bytes(w.keys_xy[w.key_idx[i], 32:]), "big")
        st_p.append(o.status(o.P256PublicKey(kx, ky), w.sig(i), bytes(w.digest[i])))
    assert list(st_c) == st_p
    assert o.INVALID in st_p and o.VALID in st_p


def test_short_and_long_digest():
    # bccsp/sw/ecdsa_test.go:47-56 signs/verifies the 11-byte "hello world" as the digest; Go hashToInt
    d = 0x1234567
    q = p256.scalar_mult(d, (p256.GX, p256.GY))
    k = o.P256PublicKey(*q)
    for dg in (b"hello world", b"\xab" * 40):
        r, s = p256.ecdsa_sign_lows(d, dg, 0xDEADBEEFCAFE)
        sig = goasn1.marshal_ecdsa_signature(r, s)
        assert o.status(k, sig, dg) == o.VALID
        assert list(_c_status([(k.x, k.y, dg, sig)])) == [o.VALID]
    # a 40-byte digest uses only its leftmost 32 bytes
    r, s = p256.ecdsa_sign_lows(d, b"\xab" * 40, 77)
    assert o.status(k, goasn1.marshal_ecdsa_signature(r, s), b"\xab" * 32 + b"\x00" * 8) == o.VALID


def test_identity_verify_hash_families():
    # msp/identities.go:169-196,216-224; msp/msp_test.go:532-535 (msg[1:], sig[1:] -> error), :569-590 (SHA3)
    import hashlib
    d = 0xBEEF
    k = o.P256PublicKey(*p256.scalar_mult(d, (p256.GX, p256.GY)))
    msg = b"fabric endorsement bytes" * 10
    for fam, h in (("SHA2", hashlib.sha256), ("SHA3", hashlib.sha3_256)):
        r, s = p256.ecdsa_sign_lows(d, h(msg).digest(), 4242)
        sig = goasn1.marshal_ecdsa_signature(r, s)
        assert o.identity_verify(k, msg, sig, fam) is None
        assert o.identity_verify(k, msg[1:], sig, fam) == "The signature is invalid"
        assert o.identity_verify(k, msg, sig[1:], fam).startswith("could not determine the validity of the signature")
    assert "hash familiy not recognized [MD5]" in o.identity_verify(k, msg, b"x", "MD5")
