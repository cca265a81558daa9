"""Host-side mirror of the reference's provider interface for the verify path, on top of the C ABI.

Names, argument meaning and error behaviour follow the reference so the parity tests read like its own:

  GPUCSP.KeyImport(raw, opts)           bccsp.BCCSP.KeyImport     bccsp/bccsp.go:101, sw import bccsp/sw/keyimport.go:114-134
  GPUCSP.Verify(k, signature, digest)   bccsp.BCCSP.Verify        bccsp/bccsp.go:123-125, bccsp/sw/impl.go:247-270
  GPUCSP.Hash(msg, opts)                bccsp.BCCSP.Hash          bccsp/sw/impl.go:177-194, bccsp/sw/hash.go:29-33
  Identity.Verify(msg, sig)             msp identity.Verify       msp/identities.go:169-196

``Verify`` returns the Go pair ``(valid, err)`` where ``err`` is ``None`` or the reference's error string.  Like the
pkcs11 provider (reference bccsp/pkcs11/pkcs11.go:241-262) only ECDSA keys are handled here; anything else is
reported as unsupported so the caller can delegate to the embedded software provider.
"""
import hashlib

import numpy as np

from . import binding

SHA2 = "SHA2"
SHA3 = "SHA3"
SHA256 = "SHA256"
SHA3_256 = "SHA3_256"


class ECDSAP256PublicKey:
    """bccsp.Key for an ECDSA P-256 public key (reference bccsp/sw/ecdsakey.go:72-117)."""

    def __init__(self, x: int, y: int):
        self.x, self.y = x, y
        self.xy = x.to_bytes(32, "big") + y.to_bytes(32, "big")

    def SKI(self) -> bytes:
        # bccsp/sw/ecdsakey.go:87-99
        return hashlib.sha256(b"\x04" + self.xy).digest()

    def Symmetric(self):
        return False

    def Private(self):
        return False

    def PublicKey(self):
        return self


class GPUCSP:
    """The GPU provider: wraps one fabgpu context (one or more B200s)."""

    def __init__(self, max_batch=65536, device_ids=None):
        self.ctx = binding.Context(max_batch=max_batch, device_ids=device_ids)

    def close(self):
        self.ctx.close()

    # -- bccsp.BCCSP ---------------------------------------------------------------------------------------
    def KeyImport(self, raw, opts=None):
        """raw: (x, y) ints, 64/65-byte uncompressed point, or an object with .x/.y (an *ecdsa.PublicKey stand-in)."""
        if raw is None:
            raise ValueError("Invalid raw. It must not be nil.")           # bccsp/sw/impl.go:120-122
        if isinstance(raw, (bytes, bytearray)):
            b = bytes(raw)
            if len(b) == 65 and b[0] == 4:
                b = b[1:]
            if len(b) != 64:
                raise ValueError("Invalid raw material. Expected 64 or 65 byte uncompressed P-256 point")
            return ECDSAP256PublicKey(int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big"))
        if isinstance(raw, tuple):
            return ECDSAP256PublicKey(int(raw[0]), int(raw[1]))
        return ECDSAP256PublicKey(int(raw.x), int(raw.y))

    def Hash(self, msg: bytes, opts=SHA256) -> bytes:
        if opts in (SHA256, SHA2):
            return hashlib.sha256(msg).digest()
        if opts in (SHA3_256, SHA3):
            return hashlib.sha3_256(msg).digest()
        raise ValueError("Unsupported 'HashOpt' provided [%s]" % (opts,))

    def Verify(self, k, signature, digest, opts=None):
        if k is None:
            return False, "Invalid Key. It must not be nil."
        if not isinstance(k, ECDSAP256PublicKey):
            return False, "Unsupported 'VerifyKey' provided [%s]" % (k,)
        return self.ctx.bccsp_verify(k.xy, signature or b"", digest or b"")

    # -- batch form (what a block-level pre-pass calls) -------------------------------------------------------
    def VerifyBatch(self, keys, key_idx, digests, signatures):
        """keys: list of ECDSAP256PublicKey (or None); key_idx[i] indexes keys (or -1); digests/signatures: lists of bytes.
        Returns uint8 status codes (binding.ST_*)."""
        keys_xy = np.zeros((max(1, len(keys)), 64), np.uint8)
        remap = np.array(key_idx, np.int32)
        for i, k in enumerate(keys):
            if k is None:
                remap[remap == i] = -1
            else:
                keys_xy[i] = np.frombuffer(k.xy, np.uint8)
        doff = np.zeros(len(digests) + 1, np.uint32)
        soff = np.zeros(len(signatures) + 1, np.uint32)
        doff[1:] = np.cumsum([len(d) for d in digests])
        soff[1:] = np.cumsum([len(s) for s in signatures])
        dig = np.frombuffer(b"".join(digests), np.uint8)
        sig = np.frombuffer(b"".join(signatures), np.uint8)
        return self.ctx.bccsp_verify_batch(keys_xy, remap, dig, doff, sig, soff)


class Identity:
    """msp identity for this path: a public key, a provider and the MSP's SignatureHashFamily (msp/identities.go:29-53)."""

    def __init__(self, csp: GPUCSP, pk: ECDSAP256PublicKey, hash_family: str = SHA2):
        self.csp, self.pk, self.hash_family = csp, pk, hash_family

    def Verify(self, msg: bytes, sig: bytes):
        """Returns None when the signature is valid, else the error string (msp/identities.go:169-196)."""
        if self.hash_family == SHA2:
            opt = SHA256
        elif self.hash_family == SHA3:
            opt = SHA3_256
        else:
            return "failed getting hash function options: hash familiy not recognized [%s]" % self.hash_family
        digest = self.csp.Hash(msg, opt)
        valid, err = self.csp.Verify(self.pk, sig, digest, None)
        if err is not None:
            return "could not determine the validity of the signature: %s" % err
        if not valid:
            return "The signature is invalid"
        return None
