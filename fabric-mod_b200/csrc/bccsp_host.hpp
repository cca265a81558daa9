// Host-side gates of the verify path, mirroring the reference's Go code that runs BEFORE crypto/ecdsa.Verify:
//   sw.CSP.Verify                 bccsp/sw/impl.go:247-270   (nil key / empty signature / empty digest, error wrap)
//   verifyECDSA                   bccsp/sw/ecdsa.go:41-57    (unmarshal, low-S)
//   utils.UnmarshalECDSASignature bccsp/utils/ecdsa.go:43-67 (Go encoding/asn1 DER rules; R > 0; S > 0)
//   utils.IsLowS                  bccsp/utils/ecdsa.go:84-92 (s <= N >> 1)
// Product code (no CUDA in this file).  A real Fabric build calls the reference's own Go functions for these
// gates (go/bccsp/gpu) and enters the library below them; this C++ mirror serves non-Go hosts and the tests.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace fabgpu { namespace host {

struct BigBytes { const uint8_t* p; size_t len; bool neg; };   // a DER INTEGER's content octets (two's complement)

// Result of gating one signature.
struct Gate {
    int status;            // FABGPU_ST_VALID = passed all gates (ask the GPU); otherwise the final status
    uint8_t r[32], s[32];  // big-endian, valid when status == FABGPU_ST_VALID
    std::string err;       // Go-style error text of verifyECDSA (empty when none); filled only if want_text
};

// asn1.Unmarshal(raw, &ECDSASignature{}) restated.  Returns true on success; on failure `why` gets Go's message.
bool unmarshal_ecdsa_signature(const uint8_t* raw, size_t n, BigBytes& r, BigBytes& s, std::string* why);

// verifyECDSA's gates on one DER signature (everything except the curve arithmetic).
void gate_signature(const uint8_t* sig, size_t sig_len, Gate& out, bool want_text);

// e = hashToInt(digest) for a 256-bit order, as 32 big-endian bytes (digest_len >= 1).
void hash_to_e(const uint8_t* digest, size_t digest_len, uint8_t e[32]);

// decimal rendering of an unsigned big-endian integer (for the "Invalid S..." message)
std::string to_decimal(const uint8_t* be, size_t n);

extern const uint8_t kHalfOrderBE[32];   // N >> 1, bccsp/utils/ecdsa.go:27-32

} }  // namespace fabgpu::host
