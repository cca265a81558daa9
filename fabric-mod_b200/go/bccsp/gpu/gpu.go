// Copyright the fabgpu authors. SPDX-License-Identifier: Apache-2.0
// +build gpu
//
// Package gpu is a bccsp.BCCSP provider that verifies ECDSA P-256 signatures on NVIDIA B200 GPUs.
//
// It follows the precedent of bccsp/pkcs11 (pkcs11.go:36-87,241-262): embed the software provider, override
// KeyImport and Verify, delegate everything else.  The gates that precede the curve arithmetic are the
// reference's own functions (utils.UnmarshalECDSASignature, utils.IsLowS), so error values are identical to
// bccsp/sw by construction; only the call at bccsp/sw/ecdsa.go:56 (ecdsa.Verify) is replaced by the GPU.
//
// bccsp.BCCSP.Verify is synchronous and per-signature, while a GPU wants thousands of signatures per launch.
// Verify therefore parks the calling goroutine on a request queue; an aggregator goroutine drains the queue into
// the library's pinned SoA slot and issues one cgo call per batch (flush on MaxBatch or FlushMicros).  Raise
// peer.validatorPoolSize (core/peer/config.go:255-258) so that enough transactions are in flight to fill batches.
//
// Safety rule (SURVEY.md section 5): a device error must never look like "invalid signature".  On any error the
// affected requests are re-run on the embedded software provider.
//
// NOT COMPILED in this repository's build image (no Go toolchain); see INTEGRATION.md.
package gpu

import (
	"crypto/ecdsa"
	"crypto/elliptic"
	"crypto/x509"
	"fmt"
	"math/big"
	"sync/atomic"
	"time"

	"github.com/hyperledger/fabric/bccsp"
	"github.com/hyperledger/fabric/bccsp/sw"
	"github.com/hyperledger/fabric/bccsp/utils"
	"github.com/pkg/errors"
)

// ecdsaP256Key wraps the software provider's key and caches the affine coordinates next to it.
type ecdsaP256Key struct {
	bccsp.Key                  // the sw key: SKI, Bytes, ... are unchanged
	pub       *ecdsa.PublicKey // parsed once at import
	x, y      [32]byte
	slot      int32  // handle of the key's table on the device(s): >= 0 window table (fabgpu_keys_register), <= -2 small table (fabgpu_keys_register_small), -1 none; atomic
	uses      uint32 // verifications requested with this key; atomic
}

// A key gets its 64 MiB window table once it has been used this often: identities the MSP imports but that sign rarely
// (or once) do not evict the tables of the busy ones.  Before that, from its smallTableAfterUses-th verification on, it owns a
// SMALL table (264 KiB, no doublings: more than five times the generic kernel's rate) -- the tier client / creator certificates live in.
const tableAfterUses = 512
const smallTableAfterUses = 32 // a small table costs about as much GPU time as 37 generic verifications: rent until then, buy from here on

// fill32 writes v as 32 big-endian bytes (big.Int.FillBytes needs Go 1.15; the reference builds with 1.14).
func fill32(dst *[32]byte, v *big.Int) {
	b := v.Bytes()
	for i := range dst {
		dst[i] = 0
	}
	copy(dst[32-len(b):], b)
}

type request struct {
	key     *ecdsaP256Key
	r, s, e [32]byte
	done    chan result
}

type result struct {
	valid bool
	err   error // non-nil only for "could not decide" -> caller falls back to sw
}

type impl struct {
	bccsp.BCCSP // embedded software provider (fallback and every non-verify method)

	dev       *device
	reqs      chan *request
	flush     time.Duration
	results   *resultCache // verdicts of the block pre-pass (prepass.go), consulted before anything is queued
	Fallbacks uint64       // exported counters for the operations endpoint
	Batches   uint64
	CacheHits uint64
}

// New returns the GPU provider.  keyStore is handed to the embedded software provider.
func New(opts GPUOpts, keyStore bccsp.KeyStore) (bccsp.BCCSP, error) {
	swCSP, err := sw.NewWithParams(opts.SecLevel, opts.HashFamily, keyStore)
	if err != nil {
		return nil, errors.Wrapf(err, "Failed initializing fallback SW BCCSP")
	}
	if opts.MaxBatch <= 0 {
		opts.MaxBatch = 65536
	}
	if opts.FlushMicros <= 0 {
		opts.FlushMicros = 200
	}
	dev, err := openDevice(opts.Devices, opts.MaxBatch)
	if err != nil {
		return nil, errors.Wrapf(err, "Failed initializing GPU BCCSP")
	}
	csp := &impl{BCCSP: swCSP, dev: dev, reqs: make(chan *request, 4*opts.MaxBatch), flush: time.Duration(opts.FlushMicros) * time.Microsecond,
		results: newResultCache(opts.ResultCacheEntries)}
	go csp.aggregate()
	return csp, nil
}

// KeyImport delegates to sw and, for ECDSA P-256 public keys, caches X||Y beside the imported key (the window table follows
// after tableAfterUses verifications).
func (csp *impl) KeyImport(raw interface{}, opts bccsp.KeyImportOpts) (bccsp.Key, error) {
	k, err := csp.BCCSP.KeyImport(raw, opts)
	if err != nil {
		return nil, err
	}
	var pub *ecdsa.PublicKey
	switch v := raw.(type) {
	case *ecdsa.PublicKey: // ECDSAGoPublicKeyImportOpts (bccsp/sw/keyimport.go:94-112)
		pub = v
	case *x509.Certificate: // X509PublicKeyImportOpts (bccsp/sw/keyimport.go:114-134): ECDSA certificates only reach this point
		pub, _ = v.PublicKey.(*ecdsa.PublicKey)
	}
	if pub == nil {
		// ECDSAPKIXPublicKeyImportOpts (DER bytes) and ECDSAPrivateKeyImportOpts: sw parsed the bytes; recover the point from
		// the key it built.  A private key verifies with its public half (ecdsaPrivateKeyVerifier, bccsp/sw/ecdsa.go:65-69).
		pub = publicHalf(k)
	}
	if pub == nil || pub.Curve != elliptic.P256() || !pub.Curve.IsOnCurve(pub.X, pub.Y) {
		return k, nil // P-384, RSA, AES, ...: stay on the software path (pkcs11.go:259-261 pattern)
	}
	gk := &ecdsaP256Key{Key: k, pub: pub, slot: -1}
	fill32(&gk.x, pub.X)
	fill32(&gk.y, pub.Y)
	return gk, nil
}

// publicHalf returns the ECDSA public key behind an sw key object (public or private), nil for anything else.  The sw key
// types are unexported; PublicKey() + Bytes() (PKIX DER, bccsp/sw/ecdsakey.go:60-66,101-110) is their public surface.
func publicHalf(k bccsp.Key) *ecdsa.PublicKey {
	pk, err := k.PublicKey()
	if err != nil {
		return nil
	}
	der, err := pk.Bytes()
	if err != nil {
		return nil
	}
	parsed, err := x509.ParsePKIXPublicKey(der)
	if err != nil {
		return nil
	}
	pub, _ := parsed.(*ecdsa.PublicKey)
	return pub
}

// The wrapper type is this provider's own: every method that hands a key to the embedded software provider unwraps it first
// (sw dispatches on reflect.TypeOf(key), bccsp/sw/impl.go:205-270).
func unwrap(k bccsp.Key) bccsp.Key {
	if gk, ok := k.(*ecdsaP256Key); ok {
		return gk.Key
	}
	return k
}

func (csp *impl) Sign(k bccsp.Key, digest []byte, opts bccsp.SignerOpts) ([]byte, error) {
	return csp.BCCSP.Sign(unwrap(k), digest, opts)
}

func (csp *impl) KeyDeriv(k bccsp.Key, opts bccsp.KeyDerivOpts) (bccsp.Key, error) {
	return csp.BCCSP.KeyDeriv(unwrap(k), opts)
}

func (csp *impl) Encrypt(k bccsp.Key, plaintext []byte, opts bccsp.EncrypterOpts) ([]byte, error) {
	return csp.BCCSP.Encrypt(unwrap(k), plaintext, opts)
}

func (csp *impl) Decrypt(k bccsp.Key, ciphertext []byte, opts bccsp.DecrypterOpts) ([]byte, error) {
	return csp.BCCSP.Decrypt(unwrap(k), ciphertext, opts)
}

// registerTable builds the key's window table on the device(s) (about 0.3-2 ms of GPU time, once).
func (csp *impl) registerTable(gk *ecdsaP256Key) {
	var xy [64]byte
	copy(xy[:32], gk.x[:])
	copy(xy[32:], gk.y[:])
	if h := csp.dev.registerKey(&xy); h >= 0 {
		atomic.StoreInt32(&gk.slot, h) // a failed registration keeps whatever table the key already has
	}
}

// registerSmallTable gives the key a small table (enqueued on the device's build stream; batches order themselves behind it).
func (csp *impl) registerSmallTable(gk *ecdsaP256Key) {
	var xy [64]byte
	copy(xy[:32], gk.x[:])
	copy(xy[32:], gk.y[:])
	if h := csp.dev.registerSmallKey(&xy); h <= -2 {
		atomic.CompareAndSwapInt32(&gk.slot, -1, h) // never replaces a window table
	}
}

// Verify has exactly sw.CSP.Verify's contract (bccsp/sw/impl.go:247-270).
func (csp *impl) Verify(k bccsp.Key, signature, digest []byte, opts bccsp.SignerOpts) (bool, error) {
	gk, ok := k.(*ecdsaP256Key)
	if !ok {
		return csp.BCCSP.Verify(k, signature, digest, opts)
	}
	if len(signature) == 0 {
		return false, errors.New("Invalid signature. Cannot be empty.")
	}
	if len(digest) == 0 {
		return false, errors.New("Invalid digest. Cannot be empty.")
	}
	// the reference's own gates, in the reference's order (bccsp/sw/ecdsa.go:41-54)
	r, s, err := utils.UnmarshalECDSASignature(signature)
	if err != nil {
		return false, errors.Wrapf(fmt.Errorf("Failed unmashalling signature [%s]", err), "Failed verifing with opts [%v]", opts)
	}
	lowS, err := utils.IsLowS(gk.pub, s)
	if err != nil {
		return false, errors.Wrapf(err, "Failed verifing with opts [%v]", opts)
	}
	if !lowS {
		return false, errors.Wrapf(fmt.Errorf("Invalid S. Must be smaller than half the order [%s][%s].", s, utils.GetCurveHalfOrdersAt(gk.pub.Curve)),
			"Failed verifing with opts [%v]", opts)
	}
	if r.BitLen() > 256 {
		return false, nil // r >= 2^256 > N: ecdsa.Verify returns false
	}
	// a verdict the block pre-pass already obtained for exactly this (key, digest, signature): no queueing, no GPU round trip
	if valid, hit := csp.results.lookup(&gk.x, &gk.y, digest, signature); hit {
		atomic.AddUint64(&csp.CacheHits, 1)
		return valid, nil
	}
	switch atomic.AddUint32(&gk.uses, 1) {
	case smallTableAfterUses:
		go csp.registerSmallTable(gk) // seen again: the cheap table
	case tableAfterUses:
		go csp.registerTable(gk) // identities are verified many times (msp/cache keeps them): worth the window table from here on
	}
	req := &request{key: gk, done: make(chan result, 1)}
	fill32(&req.r, r)
	fill32(&req.s, s)
	d := digest
	if len(d) > 32 {
		d = d[:32] // hashToInt keeps the leftmost 32 bytes for a 256-bit order
	}
	copy(req.e[32-len(d):], d)
	csp.reqs <- req
	res := <-req.done
	if res.err != nil {
		atomic.AddUint64(&csp.Fallbacks, 1)
		return csp.BCCSP.Verify(gk.Key, signature, digest, opts) // never report a device fault as "invalid"
	}
	return res.valid, nil
}

// aggregate drains requests into the pinned buffers of a free slot and enqueues one batch per flush; complete (below)
// waits for batches in launch order and answers the callers.  With the slots of a context (FABGPU_SLOTS = 3), the host fills and copies
// batch k+1 while the GPU verifies batch k.
type batch struct {
	slot    int
	pending []*request
	err     error
}

func (csp *impl) aggregate() {
	free := make(chan int, len(csp.dev.slots))
	for i := range csp.dev.slots {
		free <- i
	}
	inflight := make(chan *batch, len(csp.dev.slots))
	go csp.complete(inflight, free)
	defer close(inflight)
	timer := time.NewTimer(time.Hour)
	for {
		first, ok := <-csp.reqs
		if !ok {
			return
		}
		slotIdx := <-free // blocks only when every slot is still on the device
		pending := make([]*request, 0, csp.dev.maxBatch)
		pending = append(pending, first)
		// Reset is only defined on a stopped, drained timer: a tick left over from a batch that filled up before its deadline
		// would otherwise flush this batch at once with one or two requests in it.
		if !timer.Stop() {
			select {
			case <-timer.C:
			default:
			}
		}
		timer.Reset(csp.flush)
	fill:
		for len(pending) < csp.dev.maxBatch {
			select {
			case rq := <-csp.reqs:
				pending = append(pending, rq)
			case <-timer.C:
				break fill
			}
		}
		sl := &csp.dev.slots[slotIdx]
		for i, rq := range pending {
			o := 32 * i
			copy(sl.qx[o:o+32], rq.key.x[:])
			copy(sl.qy[o:o+32], rq.key.y[:])
			copy(sl.e[o:o+32], rq.e[:])
			copy(sl.r[o:o+32], rq.r[:])
			copy(sl.s[o:o+32], rq.s[:])
			sl.keySlot[i] = atomic.LoadInt32(&rq.key.slot) // every entry is rewritten per batch; stale handles degrade to the generic kernel
		}
		b := &batch{slot: slotIdx, pending: pending}
		b.err = csp.dev.enqueue(slotIdx, len(pending)) // H2D + kernels + D2H on the slot's stream; returns at once
		atomic.AddUint64(&csp.Batches, 1)
		inflight <- b
	}
}

func (csp *impl) complete(inflight <-chan *batch, free chan<- int) {
	for b := range inflight {
		err := b.err
		if err == nil {
			err = csp.dev.wait(b.slot)
		}
		sl := &csp.dev.slots[b.slot]
		for i, rq := range b.pending {
			if err != nil || (sl.offcurve[i>>5]>>(uint(i)&31))&1 == 1 {
				rq.done <- result{err: errors.New("gpu could not decide")}
				continue
			}
			rq.done <- result{valid: (sl.mask[i>>5]>>(uint(i)&31))&1 == 1}
		}
		free <- b.slot
	}
}
