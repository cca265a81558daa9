"""Generates tests/golden/ski_fixtures.json from the reference tree (run in the build container, where /root/reference exists).

Two kinds of reference-held values pin bccsp's key identifier SKI = SHA-256(04 || X || Y) (bccsp/sw/ecdsakey.go:29-43,87-99):
  * cryptogen writes that same value into the SubjectKeyIdentifier extension of the CA certificates it generates
    (internal/cryptogen/ca/ca.go:84,171-178) -- every certificate below with a 32-byte SKI extension;
  * the file key store names a private key file hex(SKI) + "_sk" (bccsp/sw/fileks.go) -- the keystore entries below (only the
    PUBLIC point of those keys is recorded here).
"""
import glob
import json
import os

from cryptography import x509
from cryptography.hazmat.primitives import serialization
from cryptography.hazmat.primitives.asymmetric import ec

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ski_fixtures.json")


def main():
    certs, seen = [], set()
    for f in sorted(glob.glob(REF + "/sampleconfig/msp/*/*.pem") + glob.glob(REF + "/msp/testdata/**/*.pem", recursive=True)):
        try:
            c = x509.load_pem_x509_certificate(open(f, "rb").read())
            ski = c.extensions.get_extension_for_class(x509.SubjectKeyIdentifier).value.digest
        except Exception:
            continue
        pub = c.public_key()
        if len(ski) != 32 or not isinstance(pub, ec.EllipticCurvePublicKey) or not isinstance(pub.curve, ec.SECP256R1):
            continue
        pem = c.public_bytes(serialization.Encoding.PEM).decode()
        if pem in seen:
            continue
        seen.add(pem)
        certs.append({"path": os.path.relpath(f, REF), "cert_pem": pem, "ski_hex": ski.hex()})
    keys, seenk = [], set()
    for f in sorted(glob.glob(REF + "/**/*_sk", recursive=True)):
        name = os.path.basename(f)[:-3]
        if len(name) != 64:
            continue
        try:
            sk = serialization.load_pem_private_key(open(f, "rb").read(), None)
        except Exception:
            continue
        if not isinstance(sk, ec.EllipticCurvePrivateKey) or not isinstance(sk.curve, ec.SECP256R1) or name in seenk:
            continue
        seenk.add(name)
        nums = sk.public_key().public_numbers()
        pkix = sk.public_key().public_bytes(serialization.Encoding.DER, serialization.PublicFormat.SubjectPublicKeyInfo)
        keys.append({"path": os.path.relpath(f, REF), "x_hex": "%064x" % nums.x, "y_hex": "%064x" % nums.y, "pkix_der_hex": pkix.hex(), "ski_hex": name})
    json.dump({"certificates": certs[:16], "keystore": keys}, open(OUT, "w"), indent=1)
    print("wrote %d certificates, %d keystore entries" % (len(certs[:16]), len(keys)))


if __name__ == "__main__":
    main()
