// Copyright the fabgpu authors. SPDX-License-Identifier: Apache-2.0
// +build gpu
//
// Block pre-pass: every signature the stock validator is about to check for a block -- the creator signature of each
// envelope (core/common/validation/msgvalidation.go:26-64) and each endorsement over prp || endorser
// (core/common/validation/statebased/validator_keylevel.go:243-259) -- is extracted here, verified in ONE
// fabgpu_bccsp_verify_batch call, and the verdicts are left in the result cache for bccsp.Verify to find.  The validator itself
// (core/committer/txvalidator/v20/validator.go:182-267, v14/validator.go:135-269) runs unchanged afterwards: its decision order,
// de-duplication (common/policies/policy.go:380-386) and per-namespace policies (plugindispatcher/dispatcher.go:102-221) stay
// the reference's own code.  Verifying a superset of what it will ask for cannot change any outcome (SURVEY.md A.4); an
// envelope this walk cannot parse is skipped and simply misses the cache.

package gpu

import (
	"crypto/ecdsa"
	"crypto/elliptic"
	"crypto/sha256"
	"crypto/x509"
	"encoding/pem"
	"runtime"
	"sync"

	"github.com/hyperledger/fabric-protos-go/common"
	"github.com/hyperledger/fabric/protoutil"
)

// BlockPrePasser is what extensions/validation looks for on the channel's bccsp.BCCSP.
type BlockPrePasser interface {
	PrePass(block *common.Block)
}

type sigJob struct {
	identity []byte // serialized msp.SerializedIdentity
	msg      [][]byte
	sig      []byte
}

// identityKeys caches SerializedIdentity bytes -> P-256 point (the pre-pass must not parse a PEM certificate per signature).
var identityKeys sync.Map // string(identity) -> *[64]byte, or nil entry for "not a P-256 ECDSA certificate"

func pointOf(identity []byte) *[64]byte {
	if v, ok := identityKeys.Load(string(identity)); ok {
		xy, _ := v.(*[64]byte)
		return xy
	}
	var xy *[64]byte
	if sid, err := protoutil.UnmarshalSerializedIdentity(identity); err == nil {
		if blk, _ := pem.Decode(sid.IdBytes); blk != nil {
			if cert, err := x509.ParseCertificate(blk.Bytes); err == nil {
				if pub, ok := cert.PublicKey.(*ecdsa.PublicKey); ok && pub.Curve == elliptic.P256() {
					xy = new([64]byte)
					var t [32]byte
					fill32(&t, pub.X)
					copy(xy[:32], t[:])
					fill32(&t, pub.Y)
					copy(xy[32:], t[:])
				}
			}
		}
	}
	identityKeys.Store(string(identity), xy)
	return xy
}

// extract walks one envelope the way ValidateTransaction + KeyLevelValidator.Validate do and appends its signature jobs.
func extract(envBytes []byte, jobs []sigJob) []sigJob {
	env, err := protoutil.GetEnvelopeFromBlock(envBytes)
	if err != nil || env == nil || len(env.Signature) == 0 {
		return jobs
	}
	payload, err := protoutil.UnmarshalPayload(env.Payload)
	if err != nil || payload.Header == nil {
		return jobs
	}
	shdr, err := protoutil.UnmarshalSignatureHeader(payload.Header.SignatureHeader)
	if err != nil || len(shdr.Creator) == 0 {
		return jobs
	}
	jobs = append(jobs, sigJob{identity: shdr.Creator, msg: [][]byte{env.Payload}, sig: env.Signature})
	chdr, err := protoutil.UnmarshalChannelHeader(payload.Header.ChannelHeader)
	if err != nil || common.HeaderType(chdr.Type) != common.HeaderType_ENDORSER_TRANSACTION {
		return jobs
	}
	tx, err := protoutil.UnmarshalTransaction(payload.Data)
	if err != nil {
		return jobs
	}
	for _, act := range tx.Actions {
		cap, err := protoutil.UnmarshalChaincodeActionPayload(act.Payload)
		if err != nil || cap.Action == nil {
			continue
		}
		prp := cap.Action.ProposalResponsePayload
		for _, e := range cap.Action.Endorsements {
			if e == nil || len(e.Endorser) == 0 || len(e.Signature) == 0 {
				continue
			}
			jobs = append(jobs, sigJob{identity: e.Endorser, msg: [][]byte{prp, e.Endorser}, sig: e.Signature})
		}
	}
	return jobs
}

// PrePass verifies all signatures of the block on the device and fills the result cache.  It never fails: on any device error
// the cache simply stays cold and the validator's own Verify calls take the aggregator or the software path.
func (csp *impl) PrePass(block *common.Block) {
	if block == nil || block.Data == nil || len(block.Data.Data) == 0 {
		return
	}
	jobs := make([]sigJob, 0, 4*len(block.Data.Data))
	for _, d := range block.Data.Data {
		jobs = extract(d, jobs)
	}
	n := len(jobs)
	if n == 0 {
		return
	}
	// keys: distinct identities of the block
	keyOf := map[string]int32{}
	keysXY := make([]byte, 0, 64*64)
	keyIdx := make([]int32, 0, n)
	kept := make([]int, 0, n)
	for i := range jobs {
		id := string(jobs[i].identity)
		ki, ok := keyOf[id]
		if !ok {
			xy := pointOf(jobs[i].identity)
			if xy == nil {
				keyOf[id] = -1
				continue
			}
			ki = int32(len(keysXY) / 64)
			keysXY = append(keysXY, xy[:]...)
			keyOf[id] = ki
		}
		if ki < 0 {
			continue // not P-256: the software path handles it
		}
		keyIdx = append(keyIdx, ki)
		kept = append(kept, i)
	}
	m := len(kept)
	if m == 0 {
		return
	}
	// digests: SHA-256 of the signed bytes (msp/identities.go:173-178; SHA-2 family), hashed on all cores
	digests := make([]byte, 32*m)
	workers := runtime.NumCPU()
	var wg sync.WaitGroup
	for w := 0; w < workers; w++ {
		wg.Add(1)
		go func(w int) {
			defer wg.Done()
			for j := w; j < m; j += workers {
				h := sha256.New()
				for _, part := range jobs[kept[j]].msg {
					h.Write(part)
				}
				h.Sum(digests[32*j : 32*j : 32*j+32])
			}
		}(w)
	}
	wg.Wait()
	digOff := make([]uint32, m+1)
	sigOff := make([]uint32, m+1)
	total := 0
	for j := 0; j < m; j++ {
		digOff[j+1] = uint32(32 * (j + 1))
		total += len(jobs[kept[j]].sig)
		sigOff[j+1] = uint32(total)
	}
	sigs := make([]byte, 0, total)
	for j := 0; j < m; j++ {
		sigs = append(sigs, jobs[kept[j]].sig...)
	}
	status := make([]byte, m)
	if err := csp.dev.verifyBatch(keysXY, keyIdx, digests, digOff, sigs, sigOff, status); err != nil {
		return // cold cache; never a verdict
	}
	for j := 0; j < m; j++ {
		st := status[j]
		if st > 1 {
			continue // FABGPU_ST_ERR_*: let the regular path produce the reference's error value
		}
		var x, y [32]byte
		off := 64 * int(keyIdx[j])
		copy(x[:], keysXY[off:off+32])
		copy(y[:], keysXY[off+32:off+64])
		sg := jobs[kept[j]].sig
		csp.results.store(resultKey(&x, &y, digests[32*j:32*j+32], sg), st == 0) // FABGPU_ST_VALID = 0, FABGPU_ST_INVALID = 1
	}
}
