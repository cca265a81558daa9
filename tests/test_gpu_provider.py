"""GPU tests of the provider-level behaviour around the leaf (b): key types reaching Verify, and
the Go provider's aggregator protocol replayed over the same C-ABI calls from many threads."""
import os
import threading

import numpy as np
import pytest

from oracle import bccsp_sw as o
from oracle import goasn1, p256
from tools import workload
from util import pkg
import vectors

pytestmark = pytest.mark.gpu


def _oracle_verify(k, signature, digest):
    """The embedded software provider of the test: the oracle's restatement of sw.CSP.Verify -> (valid, err)."""
    st = o.status(o.P256PublicKey(k.x, k.y), signature, digest)
    return st == o.VALID, (None if st in (o.VALID, o.INVALID) else "error status %d" % st)


@pytest.fixture(scope="module")
def csp():
    c = pkg().bccsp.GPUCSP(max_batch=8192, fallback=_oracle_verify, flush_seconds=2e-3)
    yield c
    c.close()


def test_private_key_verifies_like_its_public_half(csp):
    # ecdsaPrivateKeyVerifier (bccsp/sw/ecdsa.go:65-69) == ecdsaPublicKeyKeyVerifier on the public key
    from cryptography.hazmat.primitives import serialization
    from cryptography.hazmat.primitives.asymmetric import ec
    b = pkg().bccsp
    sk = ec.generate_private_key(ec.SECP256R1())
    d = sk.private_numbers().private_value
    der = sk.private_bytes(serialization.Encoding.DER, serialization.PrivateFormat.PKCS8, serialization.NoEncryption())
    priv = csp.KeyImport(der, b.ECDSAPrivateKeyImportOpts)
    pub = priv.PublicKey()
    import hashlib
    dg = hashlib.sha256(b"private key as verify key").digest()
    r, s = p256.ecdsa_sign_lows(d, dg, 0xABCDEF0123)
    cases = [goasn1.marshal_ecdsa_signature(r, s), goasn1.marshal_ecdsa_signature(r, p256.N - s), goasn1.marshal_ecdsa_signature(r ^ 4, s),
             goasn1.marshal_ecdsa_signature(r, s)[1:], b""]
    for sig in cases:
        got_priv, got_pub = csp.Verify(priv, sig, dg, None), csp.Verify(pub, sig, dg, None)
        assert got_priv == got_pub
        st = o.status(o.P256PublicKey(pub.x, pub.y), sig, dg)
        assert got_pub[0] == (st == o.VALID) and (got_pub[1] is None) == (st in (o.VALID, o.INVALID))
    assert csp.Verify(priv, cases[0], dg, None) == (True, None)
    # the queued (aggregator) form dispatches on the key type the same way
    assert csp.VerifyQueued(priv, cases[0], dg) == (True, None)
    assert csp.VerifyQueued(priv, cases[2], dg) == (False, None)


def test_x509_imported_key_verifies_the_certificates_it_issued(csp):
    """KeyImport(X509PublicKeyImportOpts) of a CA certificate, then Verify of the issuer signatures in the golden X.509 set that
    this CA made (tests/golden/x509_fixtures.json holds issuer keys as x/y; match on the point)."""
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    ski = json.load(open(os.path.join(here, "golden", "ski_fixtures.json")))
    fx = json.load(open(os.path.join(here, "golden", "x509_fixtures.json")))
    items = fx
    b = pkg().bccsp
    by_point = {}
    for c in ski["certificates"]:
        k = csp.KeyImport(c["cert_pem"].encode(), b.X509PublicKeyImportOpts)
        by_point[(k.x, k.y)] = k
    hits = 0
    for it in items:
        qx, qy = int(it["qx"], 16), int(it["qy"], 16)
        k = by_point.get((qx, qy))
        if k is None:
            continue
        hits += 1
        valid, err = csp.Verify(k, bytes.fromhex(it["sig_der"]), bytes.fromhex(it["digest"]), None)
        if it["expect"] == "VALID":
            assert (valid, err) == (True, None), it["cert"]
        else:                                                            # the high-S certificates (msp/cert_test.go:70-97)
            assert not valid and "Invalid S. Must be smaller than half the order" in err, it["cert"]
    assert hits >= 20


def _requests(n, keys, seed):
    w = workload.Workload(n, keys, seed=seed)
    w.tamper_r(0.2)
    b = pkg().bccsp
    ks = [b.key_import(bytes(w.keys_xy[i])) for i in range(keys)]
    reqs = [(ks[int(w.key_idx[i])], w.sig(i), bytes(w.digest[i])) for i in range(n)]
    # plus the adversarial vectors (different keys, odd digests, malformed DER: those are answered without queueing)
    for c in vectors.build():
        if c["qx"] >= p256.P or c["qy"] >= p256.P:
            continue
        reqs.append((b.key_import((c["qx"], c["qy"])), c["sig"], c["digest"]))
    return reqs


def _expected(reqs):
    out = []
    for k, sig, dg in reqs:
        st = o.status(o.P256PublicKey(k.x, k.y), sig, dg)
        out.append((st == o.VALID, st in (o.VALID, o.INVALID), st))
    return out


def _run_threads(csp, reqs, nthreads, handles=None):
    got = [None] * len(reqs)

    def worker(t):
        for i in range(t, len(reqs), nthreads):
            k, sig, dg = reqs[i]
            got[i] = csp.VerifyQueued(k, sig, dg, None, handle=(handles[i] if handles is not None else None))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return got


def test_aggregator_protocol_many_threads(csp):
    """N caller threads each make single blocking Verify calls; one aggregator thread fills pinned slots and calls
    fabgpu_verify_p256_keyed_async / fabgpu_wait.  Every caller must get the oracle's answer for ITS request."""
    reqs = _requests(3000, 6, seed=77)
    exp = _expected(reqs)
    batches0 = csp.batches
    got = _run_threads(csp, reqs, 48)
    for i, (g, e) in enumerate(zip(got, exp)):
        if e[2] == o.ERR_OFF_CURVE:
            assert g[0] == e[0]                    # decided by the fallback provider
            continue
        assert g[0] == e[0] and (g[1] is None) == e[1], (i, g, e)
    nb = csp.batches - batches0
    assert 1 <= nb < len(reqs) / 4, "requests were not aggregated into batches (%d batches)" % nb
    # every key was verified hundreds of times: from its 32nd verification on it owns a small table (gpu.go: smallTableAfterUses)
    assert csp.ctx.key_table_stats()["small"] >= 5
    # same traffic with key-table handles: registered keys take the key-table kernel, the rest the generic one
    w_keys = np.stack([np.frombuffer(k.xy, np.uint8) for k, _, _ in reqs[:3000:500]])
    hs = csp.ctx.keys_register(w_keys)
    by_xy = {bytes(w_keys[i]): int(hs[i]) for i in range(len(hs))}
    handles = [by_xy.get(k.xy, -1) for k, _, _ in reqs]
    got2 = _run_threads(csp, reqs, 32, handles)
    assert [g[0] for g in got2] == [g[0] for g in got]


def test_aggregator_device_fault_falls_back_never_invalid(csp):
    reqs = _requests(400, 3, seed=78)
    exp = _expected(reqs)
    f0 = csp.fallbacks
    os.environ["FABGPU_FAULT_INJECT"] = "1"
    try:
        got = _run_threads(csp, reqs, 16)
    finally:
        del os.environ["FABGPU_FAULT_INJECT"]
    assert csp.fallbacks - f0 >= 300               # every queued request was answered by the embedded provider
    for g, e in zip(got, exp):
        assert g[0] == e[0]
    # and the device answers again afterwards
    got = _run_threads(csp, reqs[:64], 8)
    assert [g[0] for g in got] == [e[0] for e in exp[:64]]
