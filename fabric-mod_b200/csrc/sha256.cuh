// Batched SHA-256 on the device: the digest msp identity.Verify computes before every
// bccsp.Verify (msp/identities.go:178 -> bccsp/sw/hash.go:29-33), plus the two digests the transaction checks need
// (tx id: protoutil/proputils.go:357-364; proposal hash: protoutil/txutils.go:431-448).
//
// A message is the concatenation of up to three byte ranges of one device buffer (the block), e.g. the signed bytes of
// an endorsement are  ProposalResponsePayload || endorser  (core/common/validation/statebased/validator_keylevel.go:
// 246-249) -- two ranges of the block, never copied together on the host.  One thread hashes one message; the digest is
// written as 32 big-endian bytes, which is exactly the `e` operand layout of the verify kernels (hashToInt of a 32-byte
// digest is the identity).  HBM-streaming work: ~2 KiB read per message, 32 B written.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fabgpu {

struct ShaJob { uint32_t off[3]; uint32_t len[3]; };

__device__ __constant__ uint32_t kSha256K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

__device__ __forceinline__ void sha256_compress(uint32_t* h, uint32_t* w)
{
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            const uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
            const uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        const uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        const uint32_t ch = (e & f) ^ (~e & g);
        const uint32_t t1 = hh + S1 + ch + kSha256K[i] + w[i & 15];
        const uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        const uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// Streams message bytes into 16-word blocks.  The block buffer lives in local memory so that it can be indexed with a
// run-time position (one STL per word, 16 LDL per compression); source bytes are fetched as aligned 32-bit words wherever
// the segment allows (a segment may start at any byte offset of the block).
struct ShaStream {
    uint32_t h[8];
    uint32_t wl[16];
    uint32_t fillw;      // words in wl
    uint64_t acc;        // pending bytes (nacc of them) in the low bits, most significant first
    uint32_t nacc;

    __device__ __forceinline__ void init()
    {
        h[0] = 0x6a09e667; h[1] = 0xbb67ae85; h[2] = 0x3c6ef372; h[3] = 0xa54ff53a; h[4] = 0x510e527f; h[5] = 0x9b05688c; h[6] = 0x1f83d9ab; h[7] = 0x5be0cd19;
        fillw = 0; acc = 0; nacc = 0;
    }
    __device__ __forceinline__ void emit(uint32_t be_word)
    {
        wl[fillw++] = be_word;
        if (fillw == 16) {
            uint32_t w[16];
#pragma unroll
            for (int k = 0; k < 16; k++) w[k] = wl[k];
            sha256_compress(h, w);
            fillw = 0;
        }
    }
    __device__ __forceinline__ void byte(uint32_t b)
    {
        acc = (acc << 8) | b;
        if (++nacc == 4) { emit((uint32_t)acc); acc = 0; nacc = 0; }
    }
    __device__ __forceinline__ void word_le(uint32_t le)   // four message bytes as loaded from memory
    {
        const uint32_t be = __byte_perm(le, 0, 0x0123);
        acc = (acc << 32) | be;
        emit((uint32_t)(acc >> (8 * nacc)));
        acc &= (1ull << (8 * nacc)) - 1ull;
    }
    __device__ __forceinline__ void feed(const uint8_t* __restrict__ p, uint32_t len)
    {
        uint32_t i = 0;
        while (i < len && ((uintptr_t)(p + i) & 3u)) byte(p[i++]);
        // 64 bytes at a time: all sixteen loads are issued before any is consumed (ncu showed the one-load-at-a-time
        // form stalled on long_scoreboard 4.2 cycles per issue)
        for (; i + 64 <= len; i += 64) {
            uint32_t t[16];
#pragma unroll
            for (int k = 0; k < 16; k++) t[k] = __ldg(reinterpret_cast<const uint32_t*>(p + i) + k);
#pragma unroll
            for (int k = 0; k < 16; k++) word_le(t[k]);
        }
        for (; i + 4 <= len; i += 4) word_le(__ldg(reinterpret_cast<const uint32_t*>(p + i)));
        while (i < len) byte(p[i++]);
    }
    __device__ __forceinline__ void finish(uint64_t total_bytes, uint8_t* out)
    {
        byte(0x80);
        while (nacc != 0) byte(0);
        while (fillw != 14) emit(0);
        emit((uint32_t)((total_bytes * 8) >> 32));
        emit((uint32_t)(total_bytes * 8));
        uint32_t* o = reinterpret_cast<uint32_t*>(out);
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = __byte_perm(h[k], 0, 0x0123);     // big-endian bytes
    }
};

// digests[j] = SHA-256(buf[off0 .. off0+len0) || buf[off1 ..) || buf[off2 ..))
__global__ void __launch_bounds__(128)
sha256_segments_kernel(const uint8_t* __restrict__ buf, const ShaJob* __restrict__ jobs, uint32_t n, uint8_t* __restrict__ digests)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const ShaJob job = jobs[j];
    ShaStream st;
    st.init();
#pragma unroll 1
    for (int sgi = 0; sgi < 3; sgi++) st.feed(buf + job.off[sgi], job.len[sgi]);
    st.finish((uint64_t)job.len[0] + job.len[1] + job.len[2], digests + 32 * (size_t)j);
}

}  // namespace fabgpu
