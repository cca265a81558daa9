"""Generates tests/golden/x509_fixtures.json from the reference's frozen X.509 fixtures.

Run in the build container (needs /root/reference, python `cryptography`):
    python tests/golden/make_x509_fixtures.py

Every PEM certificate in the reference tree (msp/testdata/**, sampleconfig/msp/** first, then every
other *.pem / *.crt / *.cert test fixture) carries a frozen ECDSA-P256/SHA-256 issuer signature over frozen TBSCertificate bytes.  For each
unique certificate whose issuer's public key is also among the fixtures we emit the tuple the hot
path consumes:  (issuer Qx, Qy, SHA-256(TBS), DER signature)  plus the outcome bccsp/sw must give:
  low-S  and OpenSSL-verifies -> VALID       ((true, nil))
  high-S                      -> ERR_HIGH_S  ((false, "Invalid S. ..."))   [bccsp/sw/ecdsa.go:47-54]
OpenSSL (via `cryptography`) is used only as the independent judge of the curve arithmetic.
SURVEY.md section 8(c) item (1) describes this fixture set (96 issuer-resolvable certs: 65 + 31).
"""
import glob, hashlib, json, os, sys
from cryptography import x509
from cryptography.hazmat.primitives.asymmetric import ec, utils as asym_utils
from cryptography.hazmat.primitives import hashes
from cryptography.exceptions import InvalidSignature

REF = "/root/reference"
N = 0xFFFFFFFF00000000FFFFFFFFFFFFFFFFBCE6FAADA7179E84F3B9CAC2FC632551

def load_all():
    certs = {}
    for root in ("msp/testdata", "sampleconfig/msp", "."):
        paths = []
        for ext in ("*.pem", "*.crt", "*.cert"):
            paths += glob.glob(os.path.join(REF, root, "**", ext), recursive=True)
        for path in sorted(paths):
            data = open(path, "rb").read()
            if b"BEGIN CERTIFICATE" not in data:
                continue
            for blk in data.split(b"-----END CERTIFICATE-----"):
                if b"BEGIN CERTIFICATE" not in blk:
                    continue
                pem = blk[blk.index(b"-----BEGIN CERTIFICATE"):] + b"-----END CERTIFICATE-----\n"
                try:
                    c = x509.load_pem_x509_certificate(pem)
                except Exception:
                    continue
                fp = c.fingerprint(hashes.SHA256()).hex()
                certs.setdefault(fp, (c, os.path.relpath(path, REF)))
    return certs

def main():
    certs = load_all()
    by_subject = {}
    for fp, (c, path) in certs.items():
        by_subject.setdefault(c.subject.public_bytes(), []).append(c)
    out = []
    for fp, (c, path) in sorted(certs.items(), key=lambda kv: kv[1][1]):
        pub = None
        for cand in by_subject.get(c.issuer.public_bytes(), []):
            k = cand.public_key()
            if not isinstance(k, ec.EllipticCurvePublicKey) or k.curve.name != "secp256r1":
                continue
            try:
                k.verify(c.signature, c.tbs_certificate_bytes, ec.ECDSA(hashes.SHA256()))
                pub = k
                break
            except InvalidSignature:
                continue
        if pub is None:
            continue
        if c.signature_hash_algorithm.name != "sha256":
            continue
        r, s = asym_utils.decode_dss_signature(c.signature)
        nums = pub.public_numbers()
        out.append({
            "cert": path,
            "qx": "%064x" % nums.x, "qy": "%064x" % nums.y,
            "digest": hashlib.sha256(c.tbs_certificate_bytes).hexdigest(),
            "sig_der": c.signature.hex(),
            "r": "%x" % r, "s": "%x" % s,
            "expect": "VALID" if s <= (N >> 1) else "ERR_HIGH_S",
        })
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "x509_fixtures.json")
    json.dump(out, open(dst, "w"), indent=0)
    nv = sum(1 for o in out if o["expect"] == "VALID")
    print("wrote %d fixtures (%d VALID, %d ERR_HIGH_S) from %d unique certs" % (len(out), nv, len(out) - nv, len(certs)))

if __name__ == "__main__":
    main()
