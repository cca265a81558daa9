/* ORACLE (test infrastructure, NOT product code).
 *
 * Fast CPU restatement of the reference's bccsp/sw ECDSA-P256 verifier, used as
 *   (1) the bulk checker for the CUDA path at sizes the pure-Python oracle
 *       (oracle/p256.py, oracle/bccsp_sw.py) cannot finish in seconds, and
 *   (2) the "port" CPU baseline that bench.py times on the GPU box's host cores.
 *
 * Follows, in order (same order of checks as the reference):
 *   sw.CSP.Verify                bccsp/sw/impl.go:247-270   argument gates
 *   verifyECDSA                  bccsp/sw/ecdsa.go:41-57    DER -> low-S -> ecdsa.Verify
 *   UnmarshalECDSASignature      bccsp/utils/ecdsa.go:43-67 asn1 + R>0, S>0
 *   IsLowS                       bccsp/utils/ecdsa.go:84-92 s <= N>>1
 *   [Go 1.14.4 crypto/ecdsa.Verify, not in tree; pinned Makefile:79] r,s < N; e = leftmost 32 bytes;
 *       w = s^-1; u1 = e w; u2 = r w; R = u1 G + u2 Q; R != inf; R.x mod N == r.
 * The curve arithmetic of the last step is delegated to OpenSSL libcrypto (ECDSA_do_verify,
 * nistz256 -- the same algorithm family Go's amd64 P-256 assembly was derived from); OpenSSL accepts
 * high-S and laxer DER, which is why the explicit gates above come first.  This file is cross-checked
 * against the pure-Python restatement and the golden X.509 fixtures in tests/test_oracle.py.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC ... -lcrypto -lpthread)
 */
#define OPENSSL_SUPPRESS_DEPRECATED 1
#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/ecdsa.h>
#include <openssl/obj_mac.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum {
    ST_VALID = 0, ST_INVALID = 1, ST_ERR_NIL_KEY = 2, ST_ERR_EMPTY_SIG = 3, ST_ERR_EMPTY_DIGEST = 4,
    ST_ERR_UNMARSHAL = 5, ST_ERR_R_NOT_POSITIVE = 6, ST_ERR_S_NOT_POSITIVE = 7, ST_ERR_HIGH_S = 8,
    ST_ERR_UNSUPPORTED_KEY = 9, ST_ERR_OFF_CURVE = 10
};

/* ---- Go encoding/asn1 subset (see oracle/goasn1.py for the rule list) ---- */
typedef struct { const uint8_t *p; size_t len; int neg; } asn1_int;

static int tag_and_length(const uint8_t *b, size_t n, size_t *off, int *cls, int *compound, uint32_t *tag, size_t *length)
{
    size_t o = *off;
    if (o >= n) return -1;
    uint8_t c = b[o++];
    *cls = c >> 6; *compound = (c & 0x20) != 0; *tag = c & 0x1f;
    if (*tag == 0x1f) {                       /* base-128 tag */
        uint64_t t = 0; int shifted = 0;
        for (;;) {
            if (o >= n) return -1;
            if (shifted == 5) return -1;
            c = b[o++];
            if (shifted == 0 && c == 0x80) return -1;
            t = (t << 7) | (c & 0x7f); shifted++;
            if (!(c & 0x80)) break;
        }
        if (t > 0x7fffffffu || t < 0x1f) return -1;
        *tag = (uint32_t)t;
    }
    if (o >= n) return -1;
    c = b[o++];
    if (!(c & 0x80)) { *length = c & 0x7f; }
    else {
        int nb = c & 0x7f; size_t L = 0;
        if (nb == 0) return -1;               /* indefinite */
        for (int i = 0; i < nb; i++) {
            if (o >= n) return -1;
            c = b[o++];
            if (L >= (1u << 23)) return -1;   /* length too large */
            L = (L << 8) | c;
            if (L == 0) return -1;            /* superfluous leading zeros */
        }
        if (L < 0x80) return -1;              /* non-minimal */
        *length = L;
    }
    *off = o;
    return 0;
}

static int parse_int_field(const uint8_t *b, size_t n, size_t *off, asn1_int *out)
{
    int cls, compound; uint32_t tag; size_t len;
    if (*off == n) return -1;                 /* sequence truncated */
    if (tag_and_length(b, n, off, &cls, &compound, &tag, &len)) return -1;
    if (*off + len > n) return -1;            /* data truncated */
    if (cls != 0 || tag != 2 || compound) return -1;
    const uint8_t *p = b + *off;
    if (len == 0) return -1;
    if (len > 1 && ((p[0] == 0 && !(p[1] & 0x80)) || (p[0] == 0xff && (p[1] & 0x80)))) return -1;
    out->p = p; out->len = len; out->neg = (p[0] & 0x80) != 0;
    *off += len;
    return 0;
}

/* returns 0 ok, -1 unmarshal error */
static int unmarshal_sig(const uint8_t *raw, size_t n, asn1_int *r, asn1_int *s)
{
    size_t off = 0, len; int cls, compound; uint32_t tag;
    if (n == 0) return -1;
    if (tag_and_length(raw, n, &off, &cls, &compound, &tag, &len)) return -1;
    if (off + len > n) return -1;
    if (cls != 0 || tag != 16 || !compound) return -1;
    const uint8_t *inner = raw + off; size_t ioff = 0;
    if (parse_int_field(inner, len, &ioff, r)) return -1;
    if (parse_int_field(inner, len, &ioff, s)) return -1;
    return 0;                                 /* extra bytes inside/after the SEQUENCE are ignored */
}

static int is_zero_int(const asn1_int *a)
{
    for (size_t i = 0; i < a->len; i++) if (a->p[i]) return 0;
    return 1;
}

/* ---- batch driver ---- */
typedef struct {
    const uint8_t *keys_xy; int K; const int32_t *key_idx;
    const uint8_t *digests; const uint32_t *dig_off;
    const uint8_t *sigs; const uint32_t *sig_off;
    int begin, end; uint8_t *status;
    EC_KEY **keys; uint8_t *key_state; /* per-worker lazily imported keys: 0 unseen, 1 ok, 2 off-curve */
} job_t;

static BIGNUM *g_half_n = NULL;

static uint8_t one_status(const job_t *j, int i, BN_CTX *ctx)
{
    int32_t ki = j->key_idx[i];
    if (ki < 0) return ST_ERR_NIL_KEY;
    const uint8_t *sig = j->sigs + j->sig_off[i]; size_t siglen = j->sig_off[i + 1] - j->sig_off[i];
    const uint8_t *dg = j->digests + j->dig_off[i]; size_t dglen = j->dig_off[i + 1] - j->dig_off[i];
    if (siglen == 0) return ST_ERR_EMPTY_SIG;
    if (dglen == 0) return ST_ERR_EMPTY_DIGEST;
    if (ki >= j->K) return ST_ERR_UNSUPPORTED_KEY;
    asn1_int r, s;
    if (unmarshal_sig(sig, siglen, &r, &s)) return ST_ERR_UNMARSHAL;
    if (r.neg || is_zero_int(&r)) return ST_ERR_R_NOT_POSITIVE;
    if (s.neg || is_zero_int(&s)) return ST_ERR_S_NOT_POSITIVE;
    uint8_t st = ST_INVALID;
    BIGNUM *br = BN_bin2bn(r.p, (int)r.len, NULL), *bs = BN_bin2bn(s.p, (int)s.len, NULL);
    if (BN_cmp(bs, g_half_n) > 0) { st = ST_ERR_HIGH_S; goto done; }
    if (j->key_state[ki] == 0) {             /* import once per worker, like KeyImport once per identity */
        EC_KEY *ek = EC_KEY_new_by_curve_name(NID_X9_62_prime256v1);
        BIGNUM *x = BN_bin2bn(j->keys_xy + 64 * ki, 32, NULL), *y = BN_bin2bn(j->keys_xy + 64 * ki + 32, 32, NULL);
        if (EC_KEY_set_public_key_affine_coordinates(ek, x, y) != 1) { EC_KEY_free(ek); ek = NULL; } /* off curve */
        BN_free(x); BN_free(y);
        j->keys[ki] = ek; j->key_state[ki] = ek ? 1 : 2;
    }
    if (j->key_state[ki] == 2) { st = ST_ERR_OFF_CURVE; goto done; }
    {
        ECDSA_SIG *es = ECDSA_SIG_new();
        ECDSA_SIG_set0(es, br, bs); br = bs = NULL;
        /* ECDSA_do_verify: 1 valid, 0 invalid (incl. r >= N), -1 error */
        int rc = ECDSA_do_verify(dg, (int)dglen, es, j->keys[ki]);
        st = (rc == 1) ? ST_VALID : ST_INVALID;
        ECDSA_SIG_free(es);
    }
done:
    if (br) BN_free(br);
    if (bs) BN_free(bs);
    (void)ctx;
    return st;
}

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    BN_CTX *ctx = BN_CTX_new();
    j->keys = (EC_KEY **)calloc(j->K > 0 ? j->K : 1, sizeof(EC_KEY *));
    j->key_state = (uint8_t *)calloc(j->K > 0 ? j->K : 1, 1);
    for (int i = j->begin; i < j->end; i++) j->status[i] = one_status(j, i, ctx);
    for (int k = 0; k < j->K; k++) if (j->keys[k]) EC_KEY_free(j->keys[k]);
    free(j->keys); free(j->key_state);
    BN_CTX_free(ctx);
    return NULL;
}

/* keys_xy: K x 64 bytes (X||Y big-endian).  key_idx[i] < 0 stands for a nil key.
 * digests/sigs are concatenated byte strings indexed by (n+1)-entry offset tables.
 * status[i] receives one ST_* code.  Returns 0, or -1 on allocation failure. */
int oracle_verify_batch(const uint8_t *keys_xy, int K, const int32_t *key_idx,
                        const uint8_t *digests, const uint32_t *dig_off,
                        const uint8_t *sigs, const uint32_t *sig_off,
                        int n, uint8_t *status, int nthreads)
{
    if (!g_half_n) {
        EC_GROUP *g = EC_GROUP_new_by_curve_name(NID_X9_62_prime256v1);
        BIGNUM *h = BN_new();
        EC_GROUP_get_order(g, h, NULL);
        BN_rshift1(h, h);
        g_half_n = h;
        EC_GROUP_free(g);
    }
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    job_t *jobs = (job_t *)calloc(nthreads, sizeof(job_t));
    for (int t = 0; t < nthreads; t++) {
        job_t jb = { keys_xy, K, key_idx, digests, dig_off, sigs, sig_off,
                     (int)((long long)n * t / nthreads), (int)((long long)n * (t + 1) / nthreads), status, NULL, NULL };
        jobs[t] = jb;
        if (nthreads == 1) worker(&jobs[t]);
        else pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}
