"""ORACLE (test infrastructure, NOT product code).

CPU restatement of the reference's software verifier for ECDSA P-256:

  sw.CSP.Verify                  bccsp/sw/impl.go:247-270   (argument gates, error wrap)
  -> verifyECDSA                 bccsp/sw/ecdsa.go:41-57    (DER, low-S, ecdsa.Verify)
     -> UnmarshalECDSASignature  bccsp/utils/ecdsa.go:43-67
     -> IsLowS                   bccsp/utils/ecdsa.go:84-92
     -> [Go 1.14] ecdsa.Verify   restated in oracle/p256.py
  msp identity.Verify            msp/identities.go:169-196  (hash then bccsp.Verify)

Every call returns ``(valid: bool, err: str | None)`` exactly like the Go
``(bool, error)`` pair; ``status()`` folds that into VALID / INVALID / ERR_*.

Parity status: the curve arithmetic is pinned by the 96 X.509 fixture signatures
under tests/golden/ (reference msp/testdata/**, sampleconfig/msp/**) and by
three-way agreement with OpenSSL; the DER rules by the reference's five malformed
vectors (bccsp/sw/impl_test.go:931-959) and bccsp/utils/ecdsa_test.go.  Edge cases
the reference has no vector for (u1 = 0, u1*G = +-u2*Q, r + N < P) are "parity
unpinned by reference tests" and rest on the restated Go semantics.
"""
import hashlib

from . import goasn1, p256

# status codes shared with the product header include/fabgpu_ecdsa.h
VALID = 0          # (true, nil)
INVALID = 1        # (false, nil)
ERR_NIL_KEY = 2    # "Invalid Key. It must not be nil."
ERR_EMPTY_SIG = 3  # "Invalid signature. Cannot be empty."
ERR_EMPTY_DIGEST = 4
ERR_UNMARSHAL = 5  # "Failed unmashalling signature [failed unmashalling signature [...]]"
ERR_R_NOT_POSITIVE = 6
ERR_S_NOT_POSITIVE = 7
ERR_HIGH_S = 8
ERR_UNSUPPORTED_KEY = 9
ERR_OFF_CURVE = 10  # not a reference outcome: Q off-curve is outside the restated domain


class P256PublicKey:
    """Stands in for *ecdsa.PublicKey on elliptic.P256() wrapped by sw.ecdsaPublicKey
    (bccsp/sw/ecdsakey.go:72-117)."""

    def __init__(self, x: int, y: int):
        self.x = x
        self.y = y

    def ski(self) -> bytes:
        # bccsp/sw/ecdsakey.go:87-99: SHA-256 of elliptic.Marshal (uncompressed point)
        return hashlib.sha256(b"\x04" + self.x.to_bytes(32, "big") + self.y.to_bytes(32, "big")).digest()


def unmarshal_ecdsa_signature(raw):
    """bccsp/utils/ecdsa.go:43-67. Returns (r, s, None) or (None, None, errstr)."""
    try:
        r, s = goasn1.unmarshal_ecdsa_signature(raw)
    except goasn1.Asn1Error as exc:
        return None, None, "failed unmashalling signature [asn1: %s]" % exc
    if r <= 0:
        return None, None, "invalid signature, R must be larger than zero"
    if s <= 0:
        return None, None, "invalid signature, S must be larger than zero"
    return r, s, None


def is_low_s(s: int) -> bool:
    # bccsp/utils/ecdsa.go:84-92: s.Cmp(halfOrder) != 1
    return s <= p256.HALF_N


def verify_ecdsa(key: P256PublicKey, signature: bytes, digest: bytes):
    """bccsp/sw/ecdsa.go:41-57."""
    r, s, err = unmarshal_ecdsa_signature(signature)
    if err is not None:
        return False, "Failed unmashalling signature [%s]" % err
    if not is_low_s(s):
        return False, "Invalid S. Must be smaller than half the order [%d][%d]." % (s, p256.HALF_N)
    return p256.ecdsa_verify_go114(key.x, key.y, digest, r, s), None


def csp_verify(key, signature, digest):
    """sw.CSP.Verify, bccsp/sw/impl.go:247-270 (opts is always nil on this path)."""
    if key is None:
        return False, "Invalid Key. It must not be nil."
    if signature is None or len(signature) == 0:
        return False, "Invalid signature. Cannot be empty."
    if digest is None or len(digest) == 0:
        return False, "Invalid digest. Cannot be empty."
    if not isinstance(key, P256PublicKey):
        return False, "Unsupported 'VerifyKey' provided [%s]" % (key,)
    valid, err = verify_ecdsa(key, signature, digest)
    if err is not None:
        return False, "Failed verifing with opts [<nil>]: %s" % err
    return valid, None


def status(key, signature, digest) -> int:
    """Three-valued result plus error kind, as one small integer."""
    if key is None:
        return ERR_NIL_KEY
    if signature is None or len(signature) == 0:
        return ERR_EMPTY_SIG
    if digest is None or len(digest) == 0:
        return ERR_EMPTY_DIGEST
    if not isinstance(key, P256PublicKey):
        return ERR_UNSUPPORTED_KEY
    try:
        r, s = goasn1.unmarshal_ecdsa_signature(signature)
    except goasn1.Asn1Error:
        return ERR_UNMARSHAL
    if r <= 0:
        return ERR_R_NOT_POSITIVE
    if s <= 0:
        return ERR_S_NOT_POSITIVE
    if not is_low_s(s):
        return ERR_HIGH_S
    if not p256.is_on_curve(key.x, key.y):
        return ERR_OFF_CURVE
    return VALID if p256.ecdsa_verify_go114(key.x, key.y, digest, r, s) else INVALID


def identity_verify(key, msg: bytes, sig: bytes, hash_family: str = "SHA2"):
    """msp identity.Verify, msp/identities.go:169-196: digest = Hash(msg) with
    SHA2 -> SHA-256, SHA3 -> SHA3-256 (getHashOpt :216-224); any non-valid outcome
    becomes an error.  Returns None on success, else the error string."""
    if hash_family == "SHA2":
        digest = hashlib.sha256(msg).digest()
    elif hash_family == "SHA3":
        digest = hashlib.sha3_256(msg).digest()
    else:
        return "failed getting hash function options: hash familiy not recognized [%s]" % hash_family
    valid, err = csp_verify(key, sig, digest)
    if err is not None:
        return "could not determine the validity of the signature: %s" % err
    if not valid:
        return "The signature is invalid"
    return None
