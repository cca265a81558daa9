// Copyright the fabgpu authors. SPDX-License-Identifier: Apache-2.0
// +build gpu
//
// Factory of the GPU provider, the twin of bccsp/factory/pkcs11factory.go.  Selected with
//     peer.BCCSP.Default: GPU          (sampleconfig/core.yaml:297-319)
// Compiled only with `-tags gpu` (it pulls in cgo and libfabgpu_ecdsa.so); registers itself with the hook of gpuhook.go,
// which initFactories / GetBCCSPFromOpts consult in both the pkcs11 and the nopkcs11 build (patches/bccsp-factory-gpu.patch).

package factory

import (
	"github.com/hyperledger/fabric/bccsp"
	"github.com/hyperledger/fabric/bccsp/gpu"
	"github.com/hyperledger/fabric/bccsp/sw"
	"github.com/pkg/errors"
)

type GPUFactory struct{}

func init() { gpuFactory = &GPUFactory{} }

func (f *GPUFactory) Name() string { return GPUBasedFactoryName }

func (f *GPUFactory) Get(config *FactoryOpts) (bccsp.BCCSP, error) {
	if config == nil || config.GpuOpts == nil {
		return nil, errors.New("Invalid config. It must not be nil.")
	}
	return gpu.New(*config.GpuOpts, sw.NewDummyKeyStore())
}
