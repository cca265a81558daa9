// See bccsp_host.hpp.  Error strings are the reference's, byte for byte, where the reference's tests pin them
// (bccsp/sw/ecdsa_test.go:62-73, bccsp/utils/ecdsa_test.go:19-62); the inner "asn1: ..." texts follow Go 1.14's
// encoding/asn1 messages (not pinned by any reference test beyond the "failed unmashalling signature [" prefix).
#include "bccsp_host.hpp"

#include <cstring>
#include <vector>

#include "../../include/fabgpu_ecdsa.h"

namespace fabgpu { namespace host {

const uint8_t kHalfOrderBE[32] = {
    0x7F, 0xFF, 0xFF, 0xFF, 0x80, 0x00, 0x00, 0x00, 0x7F, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF, 0xFF,
    0xDE, 0x73, 0x7D, 0x56, 0xD3, 0x8B, 0xCF, 0x42, 0x79, 0xDC, 0xE5, 0x61, 0x7E, 0x31, 0x92, 0xA8};

namespace {

struct TL { int cls; bool compound; uint32_t tag; size_t length; };

// Go parseTagAndLength.  Returns nullptr on success, else the asn1 error text.
const char* tag_and_length(const uint8_t* b, size_t n, size_t& off, TL& t)
{
    if (off >= n) return "asn1: syntax error: truncated tag or length";
    uint8_t c = b[off++];
    t.cls = c >> 6; t.compound = (c & 0x20) != 0; t.tag = c & 0x1f;
    if (t.tag == 0x1f) {
        uint64_t v = 0; int shifted = 0;
        for (;;) {
            if (off >= n) return "asn1: syntax error: truncated base 128 integer";
            if (shifted == 5) return "asn1: structure error: base 128 integer too large";
            c = b[off++];
            if (shifted == 0 && c == 0x80) return "asn1: syntax error: integer is not minimally encoded";
            v = (v << 7) | (c & 0x7f); shifted++;
            if (!(c & 0x80)) break;
        }
        if (v > 0x7fffffffull) return "asn1: structure error: base 128 integer too large";
        if (v < 0x1f) return "asn1: syntax error: non-minimal tag";
        t.tag = (uint32_t)v;
    }
    if (off >= n) return "asn1: syntax error: truncated tag or length";
    c = b[off++];
    if (!(c & 0x80)) { t.length = c & 0x7f; return nullptr; }
    const int nb = c & 0x7f;
    if (nb == 0) return "asn1: syntax error: indefinite length found (not DER)";
    size_t L = 0;
    for (int i = 0; i < nb; i++) {
        if (off >= n) return "asn1: syntax error: truncated tag or length";
        c = b[off++];
        if (L >= (1u << 23)) return "asn1: structure error: length too large";
        L = (L << 8) | c;
        if (L == 0) return "asn1: syntax error: superfluous leading zeros in length";
    }
    if (L < 0x80) return "asn1: syntax error: non-minimal length";
    t.length = L;
    return nullptr;
}

const char* parse_int_field(const uint8_t* b, size_t n, size_t& off, BigBytes& out)
{
    if (off == n) return "asn1: syntax error: sequence truncated";
    TL t;
    if (const char* e = tag_and_length(b, n, off, t)) return e;
    if (off + t.length > n) return "asn1: syntax error: data truncated";
    if (t.cls != 0 || t.tag != 2 || t.compound) return "asn1: structure error: tags don't match";
    const uint8_t* p = b + off;
    if (t.length == 0) return "asn1: structure error: empty integer";
    if (t.length > 1 && ((p[0] == 0 && !(p[1] & 0x80)) || (p[0] == 0xff && (p[1] & 0x80))))
        return "asn1: structure error: integer not minimally-encoded";
    out.p = p; out.len = t.length; out.neg = (p[0] & 0x80) != 0;
    off += t.length;
    return nullptr;
}

bool all_zero(const BigBytes& a)
{
    for (size_t i = 0; i < a.len; i++) if (a.p[i]) return false;
    return true;
}

// compare unsigned big-endian a (any length) with 32-byte b: -1, 0, 1
int cmp_be(const uint8_t* a, size_t n, const uint8_t b[32])
{
    while (n > 0 && *a == 0) { a++; n--; }
    if (n > 32) return 1;
    uint8_t t[32] = {0};
    memcpy(t + 32 - n, a, n);
    return memcmp(t, b, 32) < 0 ? -1 : (memcmp(t, b, 32) > 0 ? 1 : 0);
}

}  // namespace

bool unmarshal_ecdsa_signature(const uint8_t* raw, size_t n, BigBytes& r, BigBytes& s, std::string* why)
{
    const char* e = nullptr;
    do {
        if (n == 0 || raw == nullptr) { e = "asn1: syntax error: sequence truncated"; break; }
        size_t off = 0; TL t;
        if ((e = tag_and_length(raw, n, off, t))) break;
        if (off + t.length > n) { e = "asn1: syntax error: data truncated"; break; }
        if (t.cls != 0 || t.tag != 16 || !t.compound) { e = "asn1: structure error: tags don't match"; break; }
        const uint8_t* inner = raw + off; size_t ioff = 0;
        if ((e = parse_int_field(inner, t.length, ioff, r))) break;
        if ((e = parse_int_field(inner, t.length, ioff, s))) break;
        // bytes after S inside the SEQUENCE and bytes after the SEQUENCE are accepted (Go discards `rest`,
        // reference bccsp/utils/ecdsa.go:46)
    } while (false);
    if (e) { if (why) *why = e; return false; }
    return true;
}

std::string to_decimal(const uint8_t* be, size_t n)
{
    std::vector<uint32_t> w;                      // little-endian base 2^32
    for (size_t i = 0; i < n; i++) {
        uint64_t carry = be[i];
        for (size_t k = 0; k < w.size(); k++) { uint64_t cur = ((uint64_t)w[k] << 8) | carry; w[k] = (uint32_t)cur; carry = cur >> 32; }
        if (carry) w.push_back((uint32_t)carry);
    }
    if (w.empty()) return "0";
    std::string out;                              // least significant digit first
    while (!w.empty()) {
        uint64_t rem = 0;
        for (size_t k = w.size(); k-- > 0;) { uint64_t cur = (rem << 32) | w[k]; w[k] = (uint32_t)(cur / 1000000000u); rem = cur % 1000000000u; }
        while (!w.empty() && w.back() == 0) w.pop_back();
        if (w.empty()) { do { out.push_back((char)('0' + rem % 10)); rem /= 10; } while (rem); }
        else { for (int d = 0; d < 9; d++) { out.push_back((char)('0' + rem % 10)); rem /= 10; } }
    }
    return std::string(out.rbegin(), out.rend());
}

void gate_signature(const uint8_t* sig, size_t sig_len, Gate& out, bool want_text)
{
    BigBytes r{nullptr, 0, false}, s{nullptr, 0, false};
    std::string why;
    out.err.clear();
    if (!unmarshal_ecdsa_signature(sig, sig_len, r, s, want_text ? &why : nullptr)) {
        out.status = FABGPU_ST_ERR_UNMARSHAL;
        if (want_text) out.err = "Failed unmashalling signature [failed unmashalling signature [" + why + "]]";
        return;
    }
    if (r.neg || all_zero(r)) {
        out.status = FABGPU_ST_ERR_R_NOT_POSITIVE;
        if (want_text) out.err = "Failed unmashalling signature [invalid signature, R must be larger than zero]";
        return;
    }
    if (s.neg || all_zero(s)) {
        out.status = FABGPU_ST_ERR_S_NOT_POSITIVE;
        if (want_text) out.err = "Failed unmashalling signature [invalid signature, S must be larger than zero]";
        return;
    }
    if (cmp_be(s.p, s.len, kHalfOrderBE) > 0) {
        out.status = FABGPU_ST_ERR_HIGH_S;
        if (want_text)
            out.err = "Invalid S. Must be smaller than half the order [" + to_decimal(s.p, s.len) + "][" + to_decimal(kHalfOrderBE, 32) + "].";
        return;
    }
    // s <= N/2 fits 32 bytes.  r may be arbitrarily long: r >= 2^256 > N makes ecdsa.Verify return false.
    const uint8_t* rp = r.p; size_t rl = r.len;
    while (rl > 0 && *rp == 0) { rp++; rl--; }
    if (rl > 32) { out.status = FABGPU_ST_INVALID; return; }
    memset(out.r, 0, 32); memcpy(out.r + 32 - rl, rp, rl);
    const uint8_t* sp = s.p; size_t sl = s.len;
    while (sl > 0 && *sp == 0) { sp++; sl--; }
    memset(out.s, 0, 32); memcpy(out.s + 32 - sl, sp, sl);
    out.status = FABGPU_ST_VALID;
}

void hash_to_e(const uint8_t* digest, size_t digest_len, uint8_t e[32])
{
    // Go 1.14 crypto/ecdsa hashToInt, orderBits = 256: keep the leftmost 32 bytes; a shorter digest is a shorter integer
    if (digest_len > 32) digest_len = 32;
    memset(e, 0, 32);
    memcpy(e + 32 - digest_len, digest, digest_len);
}

} }  // namespace fabgpu::host
