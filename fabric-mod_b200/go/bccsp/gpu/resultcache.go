// Copyright the fabgpu authors. SPDX-License-Identifier: Apache-2.0
// +build gpu
//
// Verify-result cache (SURVEY.md section 8f rank 1): the block pre-pass verifies every signature of a block in one GPU batch
// and records the verdicts here; the stock validator's per-transaction Identity.Verify -> bccsp.Verify calls
// (core/common/validation/msgvalidation.go:26-64, common/policies/policy.go:365-402) then hit the cache instead of queueing.
// Key = SHA-256(X || Y || digest || signature): a verdict depends on nothing else (SURVEY.md A.4).  Only the two decided
// outcomes are stored -- (true, nil) and (false, nil); anything the device reported as an error status is NOT cached, so the
// regular path reproduces the reference's exact error value.  A miss is always safe: the caller simply verifies.

package gpu

import (
	"crypto/sha256"
	"sync"
)

type cacheKey [32]byte

// resultCache is a fixed-size, sharded, generation-evicted map: when a shard is full its older half is dropped.
type resultCache struct {
	shards [64]cacheShard
	perCap int
}

type cacheShard struct {
	mu       sync.RWMutex
	cur, old map[cacheKey]bool
}

func newResultCache(entries int) *resultCache {
	if entries <= 0 {
		entries = 1 << 20 // ~25 blocks of 10 000 transactions x 4 signatures
	}
	c := &resultCache{perCap: entries/64/2 + 1}
	for i := range c.shards {
		c.shards[i].cur = make(map[cacheKey]bool, c.perCap)
		c.shards[i].old = map[cacheKey]bool{}
	}
	return c
}

func resultKey(x, y *[32]byte, digest, signature []byte) cacheKey {
	h := sha256.New()
	h.Write(x[:])
	h.Write(y[:])
	h.Write(digest)
	h.Write(signature)
	var k cacheKey
	h.Sum(k[:0])
	return k
}

func (c *resultCache) store(k cacheKey, valid bool) {
	s := &c.shards[k[0]&63]
	s.mu.Lock()
	if len(s.cur) >= c.perCap {
		s.old, s.cur = s.cur, make(map[cacheKey]bool, c.perCap)
	}
	s.cur[k] = valid
	s.mu.Unlock()
}

func (c *resultCache) lookup(x, y *[32]byte, digest, signature []byte) (valid, hit bool) {
	k := resultKey(x, y, digest, signature)
	s := &c.shards[k[0]&63]
	s.mu.RLock()
	valid, hit = s.cur[k]
	if !hit {
		valid, hit = s.old[k]
	}
	s.mu.RUnlock()
	return
}
