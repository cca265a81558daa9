// Copyright the fabgpu authors. SPDX-License-Identifier: Apache-2.0
//
// Hook through which the GPU provider's factory (gpufactory.go, `-tags gpu`) reaches initFactories / GetBCCSPFromOpts of
// BOTH reference builds (bccsp/factory/nopkcs11.go:36-82, pkcs11.go:38-96).  Without the tag gpuFactory stays nil and asking
// for the "GPU" provider fails with the same "Could not find ..." errors an unknown provider name gets.

package factory

import (
	"github.com/hyperledger/fabric/bccsp"
	"github.com/pkg/errors"
)

// GPUBasedFactoryName is the value of `BCCSP.Default` that selects the provider.
const GPUBasedFactoryName = "GPU"

var gpuFactory BCCSPFactory // set by gpufactory.go's init() when built with -tags gpu

// initGPU is the "GPU-Based BCCSP" block of initFactories: (nil, nil) when the provider is not asked for.
func initGPU(config *FactoryOpts) (bccsp.BCCSP, error) {
	if config.ProviderName != GPUBasedFactoryName || config.GpuOpts == nil {
		return nil, nil
	}
	if gpuFactory == nil {
		return nil, errors.New("this binary was built without the `gpu` build tag")
	}
	csp, err := initBCCSP(gpuFactory, config)
	if err != nil {
		return nil, errors.Wrapf(err, "Failed initializing GPU.BCCSP")
	}
	return csp, nil
}
