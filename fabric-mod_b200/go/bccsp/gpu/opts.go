// Copyright the fabgpu authors. SPDX-License-Identifier: Apache-2.0
//
// GPUOpts lives in a file WITHOUT the `gpu` build tag (and without cgo): bccsp/factory's FactoryOpts refers to it in every
// build, while the provider itself (gpu.go, cgo.go, prepass.go) is only compiled with `-tags gpu`.

package gpu

// GPUOpts is the `GPU:` block of the BCCSP section in core.yaml (beside SW: and PKCS11:, sampleconfig/core.yaml:297-319).
type GPUOpts struct {
	SecLevel           int    `mapstructure:"security" json:"security" yaml:"Security"`
	HashFamily         string `mapstructure:"hash" json:"hash" yaml:"Hash"`
	Devices            []int  `mapstructure:"devices" json:"devices" yaml:"Devices"`
	MaxBatch           int    `mapstructure:"maxbatch" json:"maxbatch" yaml:"MaxBatch"`
	FlushMicros        int    `mapstructure:"flushmicros" json:"flushmicros" yaml:"FlushMicros"`
	ResultCacheEntries int    `mapstructure:"resultcacheentries" json:"resultcacheentries" yaml:"ResultCacheEntries"`
}
