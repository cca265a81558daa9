// Copyright the fabgpu authors. SPDX-License-Identifier: Apache-2.0
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
	keySlot         []int32 // per signature: table slot of its key (fabgpu_keys_register) or -1
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
			qx: unsafe.Slice((*byte)(unsafe.Pointer(qx)), n), qy: unsafe.Slice((*byte)(unsafe.Pointer(qy)), n),
			e: unsafe.Slice((*byte)(unsafe.Pointer(e)), n), r: unsafe.Slice((*byte)(unsafe.Pointer(r)), n),
			s:    unsafe.Slice((*byte)(unsafe.Pointer(s)), n),
			mask: unsafe.Slice((*uint32)(unsafe.Pointer(mask)), words), offcurve: unsafe.Slice((*uint32)(unsafe.Pointer(off)), words),
		}
		var ks *C.int32_t
		if rc := C.fabgpu_host_key_slots(d.ctx, C.int(i), &ks); rc != C.FABGPU_OK {
			d.close()
			return nil, fmt.Errorf("fabgpu_host_key_slots failed [%d]", int(rc))
		}
		d.slots[i].keySlot = unsafe.Slice((*int32)(unsafe.Pointer(ks)), maxBatch)
	}
	return d, nil
}

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
