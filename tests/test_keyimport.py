"""Key objects of the provider mirror: KeyImport from X.509 certificates, PKIX bytes and
private keys, SKI, and the private-key verifier dispatch.  Host-only logic: runs on the GPU-less build box.

Reference: bccsp/sw/keyimport.go:62-134, bccsp/sw/ecdsakey.go:19-117, bccsp/sw/ecdsa.go:59-75; edge-case checklist A.5 items 9-10."""
import json
import os

import pytest

from util import pkg

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = json.load(open(os.path.join(HERE, "golden", "ski_fixtures.json")))


def bc():
    return pkg().bccsp


def test_x509_import_ski_matches_the_certificates_own_subject_key_identifier():
    # cryptogen stores bccsp's SKI formula in the certificate itself (internal/cryptogen/ca/ca.go:84,171-178)
    assert len(FIX["certificates"]) >= 10
    for c in FIX["certificates"]:
        k = bc().key_import(c["cert_pem"].encode(), bc().X509PublicKeyImportOpts)
        assert isinstance(k, bc().ECDSAP256PublicKey) and not k.Private() and not k.Symmetric()
        assert k.SKI().hex() == c["ski_hex"], c["path"]
        assert k.PublicKey() is k


def test_keystore_file_names_are_the_ski_of_the_key():
    # bccsp/sw/fileks.go names a stored private key hex(SKI) + "_sk"
    assert len(FIX["keystore"]) >= 5
    for e in FIX["keystore"]:
        k = bc().key_import((int(e["x_hex"], 16), int(e["y_hex"], 16)))
        assert k.SKI().hex() == e["ski_hex"], e["path"]
        k2 = bc().key_import(bytes.fromhex(e["pkix_der_hex"]), bc().ECDSAPKIXPublicKeyImportOpts)
        assert k2.xy == k.xy and k2.SKI() == k.SKI()
        k3 = bc().key_import(b"\x04" + k.xy)
        assert k3.xy == k.xy


def test_private_key_import_and_dispatch_types():
    from cryptography.hazmat.primitives import serialization
    from cryptography.hazmat.primitives.asymmetric import ec
    sk = ec.generate_private_key(ec.SECP256R1())
    der = sk.private_bytes(serialization.Encoding.DER, serialization.PrivateFormat.PKCS8, serialization.NoEncryption())
    pem = sk.private_bytes(serialization.Encoding.PEM, serialization.PrivateFormat.TraditionalOpenSSL, serialization.NoEncryption())
    nums = sk.public_key().public_numbers()
    for raw in (der, pem):
        k = bc().key_import(raw, bc().ECDSAPrivateKeyImportOpts)
        assert isinstance(k, bc().ECDSAP256PrivateKey) and k.Private() and not k.Symmetric()
        pub = k.PublicKey()
        assert isinstance(pub, bc().ECDSAP256PublicKey) and (pub.x, pub.y) == (nums.x, nums.y)
        assert k.SKI() == pub.SKI()                                       # ecdsakey.go:29-43 hashes the PUBLIC point
        with pytest.raises(ValueError, match="Not supported"):
            k.Bytes()


def test_import_error_paths():
    b = bc()
    with pytest.raises(ValueError, match="must not be nil"):
        b.key_import(None)
    with pytest.raises(ValueError, match="Expected \\*x509.Certificate"):
        b.key_import(12345, b.X509PublicKeyImportOpts)
    with pytest.raises(ValueError, match="Failed converting PKIX"):
        b.key_import(b"\x30\x03\x02\x01\x01", b.ECDSAPKIXPublicKeyImportOpts)
    with pytest.raises(ValueError, match="must not be nil"):
        b.key_import(b"", b.ECDSAPrivateKeyImportOpts)
    # non-P-256 keys are not this provider's (P-384 at security level 384, bccsp/sw/conf.go:37-50): refused here, the caller
    # keeps them on the embedded software provider (bccsp/pkcs11/pkcs11.go:259-261 rule)
    from cryptography.hazmat.primitives import serialization
    from cryptography.hazmat.primitives.asymmetric import ec, rsa
    p384 = ec.generate_private_key(ec.SECP384R1())
    pk = p384.public_key().public_bytes(serialization.Encoding.DER, serialization.PublicFormat.SubjectPublicKeyInfo)
    with pytest.raises(ValueError, match="Failed casting to ECDSA public key"):
        b.key_import(pk, b.ECDSAPKIXPublicKeyImportOpts)
    der = p384.private_bytes(serialization.Encoding.DER, serialization.PrivateFormat.PKCS8, serialization.NoEncryption())
    with pytest.raises(ValueError, match="Failed casting to ECDSA private key"):
        b.key_import(der, b.ECDSAPrivateKeyImportOpts)
    # an RSA certificate: x509PublicKeyImportOptsKeyImporter's default branch (keyimport.go:131-133)
    import datetime
    from cryptography import x509
    from cryptography.hazmat.primitives import hashes
    from cryptography.x509.oid import NameOID
    rk = rsa.generate_private_key(65537, 2048)
    name = x509.Name([x509.NameAttribute(NameOID.COMMON_NAME, "rsa.example.com")])
    now = datetime.datetime(2026, 1, 1)
    cert = (x509.CertificateBuilder().subject_name(name).issuer_name(name).public_key(rk.public_key()).serial_number(7)
            .not_valid_before(now).not_valid_after(now + datetime.timedelta(days=30)).sign(rk, hashes.SHA256()))
    with pytest.raises(ValueError, match=r"Supported keys: \[ECDSA\]"):
        b.key_import(cert, b.X509PublicKeyImportOpts)
