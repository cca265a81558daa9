// Copyright the fabgpu authors. SPDX-License-Identifier: Apache-2.0
//
// Wrapper that NewTxValidator (extensions/validation/validation.go:48-64; reached from core/peer/peer.go:339-355) puts around
// the stock ValidationRouter: before a block goes to the v2.0 or the v1.4 validator (core/committer/txvalidator/v20/
// validator.go:182-267, v14/validator.go:135-269), the channel's crypto provider -- if it is the GPU provider -- verifies all
// of the block's signatures in one batch and remembers the verdicts; the validators then run UNCHANGED and their
// Identity.Verify calls are answered from that cache.  No build tag and no cgo here: the provider is found through an
// interface, so a peer built without `-tags gpu` takes the `ok == false` branch and behaves exactly like the reference.

package validation

import (
	"github.com/hyperledger/fabric-protos-go/common"
	"github.com/hyperledger/fabric/bccsp"
	"github.com/hyperledger/fabric/core/committer/txvalidator"
)

// blockPrePasser is implemented by bccsp/gpu's provider (prepass.go).
type blockPrePasser interface {
	PrePass(block *common.Block)
}

type prePassValidator struct {
	next txvalidator.Validator
	pre  blockPrePasser
}

// Validate implements txvalidator.Validator.
func (v *prePassValidator) Validate(block *common.Block) error {
	v.pre.PrePass(block) // never fails and never decides anything: it only warms the provider's verify-result cache
	return v.next.Validate(block)
}

// withPrePass returns next itself unless the crypto provider offers a block pre-pass.
func withPrePass(next txvalidator.Validator, cryptoProvider bccsp.BCCSP) txvalidator.Validator {
	if pre, ok := cryptoProvider.(blockPrePasser); ok {
		return &prePassValidator{next: next, pre: pre}
	}
	return next
}
