/* Synthetic workload generator (NOT the oracle, NOT product code): deterministic P-256 key pairs and
 * low-S ECDSA signatures in the shape Fabric's signers produce -- ecdsa.Sign, then ToLowS, then DER
 * (reference bccsp/sw/ecdsa.go:27-39, bccsp/utils/ecdsa.go:39-41,94-109).  Used by tests/ and bench.py
 * to build the BASELINE.json configs (seeded keys, SHA-256 digests, low-S DER).
 * Curve arithmetic by OpenSSL libcrypto; nonces from a SHA-256 counter DRBG so runs are reproducible.
 *
 * Build: make -C tools   (gcc -O2 -shared -fPIC siggen.c -lcrypto -lpthread)
 */
#define OPENSSL_SUPPRESS_DEPRECATED 1
#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/obj_mac.h>
#include <openssl/sha.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void drbg(uint64_t seed, uint64_t a, uint64_t b, const char *label, uint8_t out[32])
{
    uint8_t buf[64]; memset(buf, 0, sizeof buf);
    memcpy(buf, &seed, 8); memcpy(buf + 8, &a, 8); memcpy(buf + 16, &b, 8);
    strncpy((char *)buf + 24, label, 39);
    SHA256(buf, sizeof buf, out);
}

/* scalar in [1, n-1] from the DRBG stream (seed, a, ctr) */
static void scalar_from_drbg(uint64_t seed, uint64_t a, const char *label, const BIGNUM *order, BIGNUM *out)
{
    uint8_t h[32];
    for (uint64_t ctr = 0;; ctr++) {
        drbg(seed, a, ctr, label, h);
        BN_bin2bn(h, 32, out);
        if (!BN_is_zero(out) && BN_cmp(out, order) < 0) return;
    }
}

/* priv: K x 32 BE scalars, pub_xy: K x 64 BE affine coordinates */
int siggen_keys(uint64_t seed, int K, uint8_t *priv, uint8_t *pub_xy)
{
    EC_GROUP *g = EC_GROUP_new_by_curve_name(NID_X9_62_prime256v1);
    BN_CTX *ctx = BN_CTX_new();
    BIGNUM *order = BN_new(), *d = BN_new(), *x = BN_new(), *y = BN_new();
    EC_POINT *pt = EC_POINT_new(g);
    EC_GROUP_get_order(g, order, ctx);
    for (int k = 0; k < K; k++) {
        scalar_from_drbg(seed, (uint64_t)k, "fabgpu-key", order, d);
        EC_POINT_mul(g, pt, d, NULL, NULL, ctx);
        EC_POINT_get_affine_coordinates(g, pt, x, y, ctx);
        BN_bn2binpad(d, priv + 32 * k, 32);
        BN_bn2binpad(x, pub_xy + 64 * k, 32);
        BN_bn2binpad(y, pub_xy + 64 * k + 32, 32);
    }
    EC_POINT_free(pt); BN_free(order); BN_free(d); BN_free(x); BN_free(y); BN_CTX_free(ctx); EC_GROUP_free(g);
    return 0;
}

typedef struct {
    const uint8_t *priv; const int32_t *key_idx; const uint8_t *digests; uint64_t seed;
    int begin, end; uint8_t *r_out, *s_out;
} sjob_t;

static void *sign_worker(void *arg)
{
    sjob_t *j = (sjob_t *)arg;
    EC_GROUP *g = EC_GROUP_new_by_curve_name(NID_X9_62_prime256v1);
    BN_CTX *ctx = BN_CTX_new();
    BIGNUM *order = BN_new(), *half = BN_new(), *k = BN_new(), *kinv = BN_new(), *r = BN_new(), *s = BN_new(),
           *d = BN_new(), *e = BN_new(), *x = BN_new();
    EC_POINT *pt = EC_POINT_new(g);
    EC_GROUP_get_order(g, order, ctx);
    BN_rshift1(half, order);
    for (int i = j->begin; i < j->end; i++) {
        BN_bin2bn(j->priv + 32 * j->key_idx[i], 32, d);
        BN_bin2bn(j->digests + 32 * (size_t)i, 32, e);          /* hashToInt of a 32-byte digest */
        for (uint64_t attempt = 0;; attempt++) {
            scalar_from_drbg(j->seed ^ (attempt * 0x9E3779B97F4A7C15ull), (uint64_t)i, "fabgpu-nonce", order, k);
            EC_POINT_mul(g, pt, k, NULL, NULL, ctx);
            EC_POINT_get_affine_coordinates(g, pt, x, NULL, ctx);
            BN_nnmod(r, x, order, ctx);
            if (BN_is_zero(r)) continue;
            BN_mod_inverse(kinv, k, order, ctx);
            BN_mod_mul(s, r, d, order, ctx);
            BN_mod_add(s, s, e, order, ctx);                     /* e may exceed n: mod_add reduces */
            BN_mod_mul(s, s, kinv, order, ctx);
            if (BN_is_zero(s)) continue;
            if (BN_cmp(s, half) > 0) BN_sub(s, order, s);        /* ToLowS */
            break;
        }
        BN_bn2binpad(r, j->r_out + 32 * (size_t)i, 32);
        BN_bn2binpad(s, j->s_out + 32 * (size_t)i, 32);
    }
    EC_POINT_free(pt);
    BN_free(order); BN_free(half); BN_free(k); BN_free(kinv); BN_free(r); BN_free(s); BN_free(d); BN_free(e); BN_free(x);
    BN_CTX_free(ctx); EC_GROUP_free(g);
    return NULL;
}

/* digests: n x 32 bytes.  r_out/s_out: n x 32 BE, s already low-S. */
int siggen_sign_batch(const uint8_t *priv, const int32_t *key_idx, const uint8_t *digests, int n, uint64_t seed,
                      uint8_t *r_out, uint8_t *s_out, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    sjob_t *jobs = (sjob_t *)calloc(nthreads, sizeof(sjob_t));
    for (int t = 0; t < nthreads; t++) {
        sjob_t jb = { priv, key_idx, digests, seed, (int)((long long)n * t / nthreads),
                      (int)((long long)n * (t + 1) / nthreads), r_out, s_out };
        jobs[t] = jb;
        if (nthreads == 1) sign_worker(&jobs[t]); else pthread_create(&th[t], NULL, sign_worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}

/* n x 32 bytes of SHA-256(msg_i), msg_i = msg_len pseudo-random bytes from the DRBG (the "1 KiB message") */
int siggen_digests(uint64_t seed, int n, int msg_len, uint8_t *digests)
{
    uint8_t *msg = (uint8_t *)malloc((size_t)msg_len + 32);
    for (int i = 0; i < n; i++) {
        for (int o = 0; o < msg_len; o += 32) drbg(seed, (uint64_t)i, (uint64_t)o, "fabgpu-msg", msg + o);
        SHA256(msg, (size_t)msg_len, digests + 32 * (size_t)i);
    }
    free(msg);
    return 0;
}

/* DER-encode n (r,s) pairs (32-byte BE each) the way asn1.Marshal does (minimal INTEGERs).
 * sig_off has n+1 entries; sigs needs up to 72*n bytes.  Returns total length. */
long siggen_der(const uint8_t *r, const uint8_t *s, int n, uint8_t *sigs, uint32_t *sig_off)
{
    size_t o = 0;
    for (int i = 0; i < n; i++) {
        sig_off[i] = (uint32_t)o;
        uint8_t body[80]; size_t b = 0;
        const uint8_t *v[2] = { r + 32 * (size_t)i, s + 32 * (size_t)i };
        for (int q = 0; q < 2; q++) {
            int z = 0; while (z < 31 && v[q][z] == 0) z++;
            int pad = (v[q][z] & 0x80) ? 1 : 0;
            body[b++] = 0x02; body[b++] = (uint8_t)(32 - z + pad);
            if (pad) body[b++] = 0;
            memcpy(body + b, v[q] + z, 32 - z); b += 32 - z;
        }
        sigs[o++] = 0x30; sigs[o++] = (uint8_t)b;
        memcpy(sigs + o, body, b); o += b;
    }
    sig_off[n] = (uint32_t)o;
    return (long)o;
}
