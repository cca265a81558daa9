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
import queue
import threading

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


class ECDSAP256PrivateKey:
    """bccsp.Key for an ECDSA P-256 private key (reference bccsp/sw/ecdsakey.go:19-70): SKI is the public point's, Bytes() is
    refused, PublicKey() hands out the public half.  Verify accepts it and behaves as with that public half
    (ecdsaPrivateKeyVerifier, bccsp/sw/ecdsa.go:65-69)."""

    def __init__(self, d: int, x: int, y: int):
        self.d = d
        self._pub = ECDSAP256PublicKey(x, y)

    def SKI(self) -> bytes:
        return self._pub.SKI()

    def Bytes(self):
        raise ValueError("Not supported.")                                 # bccsp/sw/ecdsakey.go:25-27

    def Symmetric(self):
        return False

    def Private(self):
        return True

    def PublicKey(self):
        return self._pub


# bccsp key-import options on this path (bccsp/opts.go): the sw importers are at bccsp/sw/keyimport.go:62-134
ECDSAGoPublicKeyImportOpts = "ECDSAGoPublicKeyImportOpts"
ECDSAPKIXPublicKeyImportOpts = "ECDSAPKIXPublicKeyImportOpts"
ECDSAPrivateKeyImportOpts = "ECDSAPrivateKeyImportOpts"
X509PublicKeyImportOpts = "X509PublicKeyImportOpts"

_P256_OID = "1.2.840.10045.3.1.7"


def _point_of(pub):
    """(x, y) of a `cryptography` EC public key on P-256; any other key -> None (the provider keeps P-256 only and leaves
    everything else to the embedded software provider, like bccsp/pkcs11/pkcs11.go:259-261)."""
    try:
        from cryptography.hazmat.primitives.asymmetric import ec
    except ImportError:                                                    # pragma: no cover
        return None
    if not isinstance(pub, ec.EllipticCurvePublicKey) or not isinstance(pub.curve, ec.SECP256R1):
        return None
    nums = pub.public_numbers()
    return nums.x, nums.y


def key_import(raw, opts=None):
    """opts None / ECDSAGoPublicKeyImportOpts: raw = (x, y) ints, a 64/65-byte uncompressed point, or an object with
    .x/.y (an *ecdsa.PublicKey stand-in).  X509PublicKeyImportOpts: raw = a `cryptography` x509.Certificate, PEM or DER
    bytes (bccsp/sw/keyimport.go:114-134: ECDSA certificates only).  ECDSAPKIXPublicKeyImportOpts: raw = PKIX DER bytes
    (keyimport.go:62-81).  ECDSAPrivateKeyImportOpts: raw = DER / PEM private key bytes (keyimport.go:83-101)."""
    if raw is None:
        raise ValueError("Invalid raw. It must not be nil.")           # bccsp/sw/impl.go:120-122
    if opts == X509PublicKeyImportOpts:
        from cryptography import x509
        cert = raw
        if isinstance(raw, (bytes, bytearray)):
            b = bytes(raw)
            cert = x509.load_pem_x509_certificate(b) if b.lstrip().startswith(b"-----BEGIN") else x509.load_der_x509_certificate(b)
        if not hasattr(cert, "public_key"):
            raise ValueError("Invalid raw material. Expected *x509.Certificate.")
        from cryptography.hazmat.primitives.asymmetric import ec
        pub = cert.public_key()
        if not isinstance(pub, ec.EllipticCurvePublicKey):
            raise ValueError("Certificate's public key type not recognized. Supported keys: [ECDSA]")
        pt = _point_of(pub)
        if pt is None:
            raise ValueError("Unsupported curve for the GPU provider (P-256 only): delegate to the software provider")
        return ECDSAP256PublicKey(*pt)
    if opts == ECDSAPKIXPublicKeyImportOpts:
        from cryptography.hazmat.primitives import serialization
        if not isinstance(raw, (bytes, bytearray)):
            raise ValueError("Invalid raw material. Expected byte array.")
        if len(raw) == 0:
            raise ValueError("Invalid raw. It must not be nil.")
        try:
            pub = serialization.load_der_public_key(bytes(raw))
        except Exception as e:
            raise ValueError("Failed converting PKIX to ECDSA public key [%s]" % e)
        pt = _point_of(pub)
        if pt is None:
            raise ValueError("Failed casting to ECDSA public key. Invalid raw material.")
        return ECDSAP256PublicKey(*pt)
    if opts == ECDSAPrivateKeyImportOpts:
        from cryptography.hazmat.primitives import serialization
        from cryptography.hazmat.primitives.asymmetric import ec
        if not isinstance(raw, (bytes, bytearray)):
            raise ValueError("[ECDSADERPrivateKeyImportOpts] Invalid raw material. Expected byte array.")
        if len(raw) == 0:
            raise ValueError("[ECDSADERPrivateKeyImportOpts] Invalid raw. It must not be nil.")
        b = bytes(raw)
        try:
            sk = serialization.load_pem_private_key(b, None) if b.lstrip().startswith(b"-----BEGIN") else serialization.load_der_private_key(b, None)
        except Exception as e:
            raise ValueError("Failed converting PKIX to ECDSA public key [%s]" % e)
        if not isinstance(sk, ec.EllipticCurvePrivateKey) or not isinstance(sk.curve, ec.SECP256R1):
            raise ValueError("Failed casting to ECDSA private key. Invalid raw material.")
        nums = sk.private_numbers()
        return ECDSAP256PrivateKey(nums.private_value, nums.public_numbers.x, nums.public_numbers.y)
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



class GPUCSP:
    """The GPU provider: wraps one fabgpu context (one or more B200s).

    fallback: the embedded software provider's Verify, ``(key, signature, digest) -> (valid, err)`` -- what the Go provider
    holds as its embedded sw CSP.  It is used only when the DEVICE fails (a fault must never read as "invalid signature",
    SURVEY.md section 5); without one such a failure raises."""

    def __init__(self, max_batch=65536, device_ids=None, fallback=None, flush_seconds=200e-6):
        self.ctx = binding.Context(max_batch=max_batch, device_ids=device_ids)
        self.max_batch = max_batch
        self.fallback = fallback
        self.flush_seconds = flush_seconds
        self.fallbacks = 0
        self.batches = 0
        # per-key use counts and table handles, as the Go provider keeps them on its key objects (gpu.go: smallTableAfterUses / tableAfterUses)
        self._key_state = {}
        self._key_mu = threading.Lock()
        self.small_table_after_uses = 32
        self.table_after_uses = 512
        self._reqs = None
        self._agg = None

    def close(self):
        if self._agg is not None:
            self._reqs.put(None)
            self._agg.join()
            self._agg = None
        self.ctx.close()

    # -- bccsp.BCCSP ---------------------------------------------------------------------------------------
    def KeyImport(self, raw, opts=None):
        """bccsp.BCCSP.KeyImport; see key_import (host-only: no device is involved in importing a key)."""
        return key_import(raw, opts)

    def Hash(self, msg: bytes, opts=SHA256) -> bytes:
        if opts in (SHA256, SHA2):
            return hashlib.sha256(msg).digest()
        if opts in (SHA3_256, SHA3):
            return hashlib.sha3_256(msg).digest()
        raise ValueError("Unsupported 'HashOpt' provided [%s]" % (opts,))

    def Verify(self, k, signature, digest, opts=None):
        if k is None:
            return False, "Invalid Key. It must not be nil."
        if isinstance(k, ECDSAP256PrivateKey):                             # ecdsaPrivateKeyVerifier: the public half decides
            k = k.PublicKey()
        if not isinstance(k, ECDSAP256PublicKey):
            return False, "Unsupported 'VerifyKey' provided [%s]" % (k,)
        return self.ctx.bccsp_verify(k.xy, signature or b"", digest or b"")

    # -- the Go provider's aggregator (fabric-mod_b200/go/bccsp/gpu/gpu.go), replayed over the same C-ABI calls ------------
    # bccsp.BCCSP.Verify is synchronous and per signature; up to validatorPoolSize goroutines call it at once
    # (core/committer/txvalidator/v20/validator.go:195-210).  VerifyQueued is that call: it runs the reference's gates on the
    # calling thread, parks the request on a queue and waits; ONE aggregator thread drains the queue into a pinned SoA slot
    # (fabgpu_host_buffers / fabgpu_host_key_slots), issues fabgpu_verify_p256_keyed_async and later fabgpu_wait, and answers
    # every waiter from the mask.  A device error -- or an off-curve flag -- sends the request to `fallback`.
    def _start_aggregator(self):
        if self._agg is None:
            self._reqs = queue.Queue()
            self._slots = [(self.ctx.host_buffers(i), self.ctx.host_key_slots(i)) for i in range(binding.SLOTS)]
            self._agg = threading.Thread(target=self._aggregate, daemon=True)
            self._agg.start()

    def _handle_of(self, k):
        """The table handle VerifyQueued passes for key k: none at first, a small table from the 32nd verification on, the window table
        from the 512th (gpu.go: registerSmallTable / registerTable)."""
        with self._key_mu:
            st = self._key_state.setdefault(k.xy, [0, -1])
            st[0] += 1
            uses = st[0]
        if uses == self.small_table_after_uses:
            h = int(self.ctx.keys_register_small(np.frombuffer(k.xy, np.uint8))[0])
            with self._key_mu:
                if h <= -2 and st[1] == -1:
                    st[1] = h
        elif uses == self.table_after_uses:
            h = int(self.ctx.keys_register(np.frombuffer(k.xy, np.uint8))[0])
            with self._key_mu:
                if h >= 0:
                    st[1] = h
        return st[1]

    def VerifyQueued(self, k, signature, digest, opts=None, handle=None):
        if k is None:
            return False, "Invalid Key. It must not be nil."
        if isinstance(k, ECDSAP256PrivateKey):
            k = k.PublicKey()
        if not isinstance(k, ECDSAP256PublicKey):
            return False, "Unsupported 'VerifyKey' provided [%s]" % (k,)
        if not signature:
            return False, "Invalid signature. Cannot be empty."
        if not digest:
            return False, "Invalid digest. Cannot be empty."
        st, r, s = binding.gate_signature(signature)                       # DER, positivity, low-S: the host gates of the library
        if st != binding.ST_VALID:
            return self.ctx.bccsp_verify(k.xy, signature, digest)          # rare: let the one-signature entry point word the error
        self._start_aggregator()
        if handle is None:
            handle = self._handle_of(k)
        d = bytes(digest[:32])
        req = {"k": k, "sig": signature, "dig": digest, "e": b"\x00" * (32 - len(d)) + d, "r": r, "s": s, "h": handle,
               "ev": threading.Event(), "res": None}
        self._reqs.put(req)
        req["ev"].wait()
        if req["res"] is None:                                             # the device could not decide
            self.fallbacks += 1
            if self.fallback is None:
                raise binding.FabGpuError(binding.E_CUDA, "device failure and no fallback provider")
            return self.fallback(k, signature, digest)
        return req["res"], None

    def _aggregate(self):
        inflight = []                                                      # (slot, requests, enqueue error)
        free = list(range(binding.SLOTS))

        def retire():
            slot, reqs, err = inflight.pop(0)
            if err is None:
                try:
                    self.ctx.wait(slot)
                except binding.FabGpuError as e:
                    err = e
            hb = self._slots[slot][0]
            for i, rq in enumerate(reqs):
                off = (int(hb["offcurve"][i >> 5]) >> (i & 31)) & 1
                rq["res"] = None if (err is not None or off) else bool((int(hb["mask"][i >> 5]) >> (i & 31)) & 1)
                rq["ev"].set()
            free.append(slot)

        stop = False
        while not stop:
            try:
                first = self._reqs.get(timeout=0.001 if inflight else None)
            except queue.Empty:
                retire()
                continue
            if first is None:
                break
            if not free:
                retire()
            slot = free.pop(0)
            pending = [first]
            import time
            deadline = time.monotonic() + self.flush_seconds
            while len(pending) < self.max_batch:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                try:
                    rq = self._reqs.get(timeout=left)
                except queue.Empty:
                    break
                if rq is None:
                    stop = True
                    break
                pending.append(rq)
            hb, ks = self._slots[slot]
            for i, rq in enumerate(pending):
                hb["qx"][i] = np.frombuffer(rq["k"].xy[:32], np.uint8)
                hb["qy"][i] = np.frombuffer(rq["k"].xy[32:], np.uint8)
                hb["e"][i] = np.frombuffer(rq["e"], np.uint8)
                hb["r"][i] = np.frombuffer(rq["r"], np.uint8)
                hb["s"][i] = np.frombuffer(rq["s"], np.uint8)
                ks[i] = rq["h"]                                            # table handle or -1; stale handles degrade to the generic kernel
            err = None
            try:
                self.ctx.verify_p256_keyed_async(slot, len(pending))
            except binding.FabGpuError as e:
                err = e
            self.batches += 1
            inflight.append((slot, pending, err))
        while inflight:
            retire()

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
