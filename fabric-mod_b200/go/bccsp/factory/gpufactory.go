// Copyright the fabgpu authors. SPDX-License-Identifier: Apache-2.0
// +build gpu
//
// Factory of the GPU provider, the twin of bccsp/factory/pkcs11factory.go.  Selected with
//     peer.BCCSP.Default: GPU          (sampleconfig/core.yaml:297-319)
// NOT COMPILED in this repository's build image (no Go toolchain); see INTEGRATION.md for the three-line
// change to initFactories / GetBCCSPFromOpts (bccsp/factory/pkcs11.go:38-96) that registers it.

package factory

import (
	"github.com/hyperledger/fabric/bccsp"
	"github.com/hyperledger/fabric/bccsp/gpu"
	"github.com/hyperledger/fabric/bccsp/sw"
	"github.com/pkg/errors"
)

const GPUBasedFactoryName = "GPU"

type GPUFactory struct{}

func (f *GPUFactory) Name() string { return GPUBasedFactoryName }

func (f *GPUFactory) Get(config *FactoryOpts) (bccsp.BCCSP, error) {
	if config == nil || config.GpuOpts == nil {
		return nil, errors.New("Invalid config. It must not be nil.")
	}
	return gpu.New(*config.GpuOpts, sw.NewDummyKeyStore())
}
