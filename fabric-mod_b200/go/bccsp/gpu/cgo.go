// Copyright the fabgpu authors. SPDX-License-Identifier: Apache-2.0
// +build gpu
//
// cgo binding of include/fabgpu_ecdsa.h (libfabgpu_ecdsa.so).  NOT COMPILED in the build image of this
// repository (it has no Go toolchain); shipped as the reference-side binding a maintainer adds under
// github.com/hyperledger/fabric/bccsp/gpu.  See INTEGRATION.md.

package gpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -lfabgpu_ecdsa
#include <stdlib.h>
#include "fabgpu_ecdsa.h"
*/
import "C"

import (
	"fmt"
	"unsafe"
)

// device is one fabgpu context plus Go views of its pinned SoA slots.
type device struct {
	ctx      *C.fabgpu_ctx
	maxBatch int
	slots    [C.FABGPU_SLOTS]slot
}

// slot exposes the library-owned pinned buffers as Go slices (no Go pointer is ever retained by C).
type slot struct {
	qx, qy, e, r, s []byte
	mask, offcurve  []uint32
	keySlot         []int32 // per signature: handle of its key's table (fabgpu_keys_register / fabgpu_keys_register_small) or -1
}

func openDevice(deviceIDs []int, maxBatch int) (*device, error) {
	d := &device{maxBatch: maxBatch}
	var ids *C.int
	cids := make([]C.int, len(deviceIDs))
	for i, v := range deviceIDs {
		cids[i] = C.int(v)
	}
	if len(cids) > 0 {
		ids = &cids[0]
	}
	if rc := C.fabgpu_init(ids, C.int(len(cids)), C.size_t(maxBatch), &d.ctx); rc != C.FABGPU_OK {
		return nil, fmt.Errorf("fabgpu_init failed [%d]: %s", int(rc), C.GoString(C.fabgpu_last_error(nil)))
	}
	words := (maxBatch + 31) / 32
	for i := range d.slots {
		var qx, qy, e, r, s *C.uint8_t
		var mask, off *C.uint32_t
		if rc := C.fabgpu_host_buffers(d.ctx, C.int(i), &qx, &qy, &e, &r, &s, &mask, &off); rc != C.FABGPU_OK {
			d.close()
			return nil, fmt.Errorf("fabgpu_host_buffers failed [%d]", int(rc))
		}
		n := maxBatch * 32
		d.slots[i] = slot{
			qx: cBytes(unsafe.Pointer(qx), n), qy: cBytes(unsafe.Pointer(qy), n),
			e: cBytes(unsafe.Pointer(e), n), r: cBytes(unsafe.Pointer(r), n), s: cBytes(unsafe.Pointer(s), n),
			mask: cUint32s(unsafe.Pointer(mask), words), offcurve: cUint32s(unsafe.Pointer(off), words),
		}
		var ks *C.int32_t
		if rc := C.fabgpu_host_key_slots(d.ctx, C.int(i), &ks); rc != C.FABGPU_OK {
			d.close()
			return nil, fmt.Errorf("fabgpu_host_key_slots failed [%d]", int(rc))
		}
		d.slots[i].keySlot = cInt32s(unsafe.Pointer(ks), maxBatch)
	}
	return d, nil
}

// Go views of C-owned (pinned) memory.  The reference builds with Go 1.14.4 (Makefile:79): unsafe.Slice (Go 1.17) is not
// available, so the views are made the pre-1.17 way -- cast to a pointer to a huge array type, then slice with a capacity.
// The array types only bound the index arithmetic; nothing of that size is allocated.
func cBytes(p unsafe.Pointer, n int) []byte     { return (*[1 << 40]byte)(p)[:n:n] }
func cUint32s(p unsafe.Pointer, n int) []uint32 { return (*[1 << 38]uint32)(p)[:n:n] }
func cInt32s(p unsafe.Pointer, n int) []int32   { return (*[1 << 38]int32)(p)[:n:n] }

func (d *device) close() {
	if d.ctx != nil {
		C.fabgpu_destroy(d.ctx)
		d.ctx = nil
	}
}

// enqueue starts H2D + kernel(s) + D2H for the first n tuples of slot i (key slots included) on the slot's stream and
// returns; wait blocks until that batch is complete.  One cgo call per batch and phase, never per signature.
func (d *device) enqueue(i, n int) error {
	if rc := C.fabgpu_verify_p256_keyed_async(d.ctx, C.int(i), C.size_t(n)); rc != C.FABGPU_OK {
		return fmt.Errorf("fabgpu_verify_p256_keyed_async failed [%d]: %s", int(rc), C.GoString(C.fabgpu_last_error(d.ctx)))
	}
	return nil
}

func (d *device) wait(i int) error {
	if rc := C.fabgpu_wait(d.ctx, C.int(i)); rc != C.FABGPU_OK {
		return fmt.Errorf("fabgpu_wait failed [%d]: %s", int(rc), C.GoString(C.fabgpu_last_error(d.ctx)))
	}
	return nil
}

// registerKey builds (or finds) the fixed-base table of one public key; -1 means "no table" and is always usable.
func (d *device) registerKey(xy *[64]byte) int32 {
	var s C.int32_t = -1
	if rc := C.fabgpu_keys_register(d.ctx, (*C.uint8_t)(unsafe.Pointer(&xy[0])), 1, &s); rc != C.FABGPU_OK {
		return -1
	}
	return int32(s)
}

// registerSmallKey builds (or finds) the small table of one public key (handle <= -2); -1 means "no table".
func (d *device) registerSmallKey(xy *[64]byte) int32 {
	var s C.int32_t = -1
	if rc := C.fabgpu_keys_register_small(d.ctx, (*C.uint8_t)(unsafe.Pointer(&xy[0])), 1, &s); rc != C.FABGPU_OK {
		return -1
	}
	return int32(s)
}

// verifyBatch is the bccsp-level batch entry point (fabgpu_bccsp_verify_batch): raw DER signatures, digests and keys in Go
// memory (borrowed for the call; the library stages them into its own pinned buffers), one status byte per signature out.
// Used by the block pre-pass (prepass.go).  keysXY: K x 64 bytes; keyIdx[i] names signature i's key.
func (d *device) verifyBatch(keysXY []byte, keyIdx []int32, digests []byte, digOff []uint32, sigs []byte, sigOff []uint32, status []byte) error {
	n := len(keyIdx)
	if n == 0 {
		return nil
	}
	if len(sigs) == 0 {
		sigs = []byte{0}
	}
	if len(digests) == 0 {
		digests = []byte{0}
	}
	rc := C.fabgpu_bccsp_verify_batch(d.ctx, (*C.uint8_t)(unsafe.Pointer(&keysXY[0])), C.int(len(keysXY)/64), (*C.int32_t)(unsafe.Pointer(&keyIdx[0])),
		(*C.uint8_t)(unsafe.Pointer(&digests[0])), (*C.uint32_t)(unsafe.Pointer(&digOff[0])),
		(*C.uint8_t)(unsafe.Pointer(&sigs[0])), (*C.uint32_t)(unsafe.Pointer(&sigOff[0])), C.size_t(n), (*C.uint8_t)(unsafe.Pointer(&status[0])))
	if rc != C.FABGPU_OK {
		return fmt.Errorf("fabgpu_bccsp_verify_batch failed [%d]: %s", int(rc), C.GoString(C.fabgpu_last_error(d.ctx)))
	}
	return nil
}
