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
 * The steps of ecdsa.Verify are spelled out below with OpenSSL libcrypto primitives (BN_mod_inverse, BN_mod_mul,
 * EC_POINT_mul = nistz256 -- the algorithm family Go's amd64 P-256 assembly was derived from).  Each worker thread owns
 * its EC_GROUP, BN_CTX and imported keys, so threads share no OpenSSL object (ECDSA_do_verify on shared keys scaled
 * badly on the 128-thread GPU hosts).  This file is cross-checked
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
    /* per-worker state: nothing below is shared between threads */
    EC_GROUP *group; BN_CTX *ctx; BIGNUM *order, *half, *r, *s, *e, *w, *u1, *u2, *x;
    EC_POINT *R; EC_POINT **keys; uint8_t *key_state; /* lazily imported keys: 0 unseen, 1 ok, 2 off-curve */
} job_t;

static uint8_t one_status(job_t *j, int i)
{
    int32_t ki = j->key_idx[i];
    if (ki < 0) return ST_ERR_NIL_KEY;                                   /* bccsp/sw/impl.go:249-251 */
    const uint8_t *sig = j->sigs + j->sig_off[i]; size_t siglen = j->sig_off[i + 1] - j->sig_off[i];
    const uint8_t *dg = j->digests + j->dig_off[i]; size_t dglen = j->dig_off[i + 1] - j->dig_off[i];
    if (siglen == 0) return ST_ERR_EMPTY_SIG;                            /* :252-254 */
    if (dglen == 0) return ST_ERR_EMPTY_DIGEST;                          /* :255-257 */
    if (ki >= j->K) return ST_ERR_UNSUPPORTED_KEY;
    asn1_int r, s;
    if (unmarshal_sig(sig, siglen, &r, &s)) return ST_ERR_UNMARSHAL;     /* bccsp/utils/ecdsa.go:46-49 */
    if (r.neg || is_zero_int(&r)) return ST_ERR_R_NOT_POSITIVE;          /* :59-61 */
    if (s.neg || is_zero_int(&s)) return ST_ERR_S_NOT_POSITIVE;          /* :62-64 */
    BN_bin2bn(r.p, (int)r.len, j->r);
    BN_bin2bn(s.p, (int)s.len, j->s);
    if (BN_cmp(j->s, j->half) > 0) return ST_ERR_HIGH_S;                 /* bccsp/sw/ecdsa.go:47-54 */
    if (j->key_state[ki] == 0) {                                         /* import once per worker, like KeyImport per identity */
        EC_POINT *q = EC_POINT_new(j->group);
        BIGNUM *x = BN_bin2bn(j->keys_xy + 64 * ki, 32, NULL), *y = BN_bin2bn(j->keys_xy + 64 * ki + 32, 32, NULL);
        int ok = EC_POINT_set_affine_coordinates(j->group, q, x, y, j->ctx) == 1;   /* fails for off-curve points */
        BN_free(x); BN_free(y);
        if (!ok) { EC_POINT_free(q); q = NULL; }
        j->keys[ki] = q; j->key_state[ki] = q ? 1 : 2;
    }
    if (j->key_state[ki] == 2) return ST_ERR_OFF_CURVE;
    /* Go 1.14 ecdsa.Verify (reached at bccsp/sw/ecdsa.go:56) */
    if (BN_cmp(j->r, j->order) >= 0 || BN_cmp(j->s, j->order) >= 0) return ST_INVALID;
    BN_bin2bn(dg, (int)(dglen > 32 ? 32 : dglen), j->e);                 /* hashToInt */
    if (!BN_mod_inverse(j->w, j->s, j->order, j->ctx)) return ST_INVALID;
    BN_mod_mul(j->u1, j->e, j->w, j->order, j->ctx);
    BN_mod_mul(j->u2, j->r, j->w, j->order, j->ctx);
    if (!EC_POINT_mul(j->group, j->R, j->u1, j->keys[ki], j->u2, j->ctx)) return ST_INVALID;
    if (EC_POINT_is_at_infinity(j->group, j->R)) return ST_INVALID;
    if (!EC_POINT_get_affine_coordinates(j->group, j->R, j->x, NULL, j->ctx)) return ST_INVALID;
    BN_nnmod(j->x, j->x, j->order, j->ctx);
    return BN_cmp(j->x, j->r) == 0 ? ST_VALID : ST_INVALID;
}

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    j->group = EC_GROUP_new_by_curve_name(NID_X9_62_prime256v1);
    j->ctx = BN_CTX_new();
    j->order = BN_new(); j->half = BN_new(); j->r = BN_new(); j->s = BN_new(); j->e = BN_new(); j->w = BN_new();
    j->u1 = BN_new(); j->u2 = BN_new(); j->x = BN_new();
    EC_GROUP_get_order(j->group, j->order, j->ctx);
    BN_rshift1(j->half, j->order);                                       /* bccsp/utils/ecdsa.go:27-32 */
    j->R = EC_POINT_new(j->group);
    j->keys = (EC_POINT **)calloc(j->K > 0 ? j->K : 1, sizeof(EC_POINT *));
    j->key_state = (uint8_t *)calloc(j->K > 0 ? j->K : 1, 1);
    for (int i = j->begin; i < j->end; i++) j->status[i] = one_status(j, i);
    for (int k = 0; k < j->K; k++) if (j->keys[k]) EC_POINT_free(j->keys[k]);
    free(j->keys); free(j->key_state);
    EC_POINT_free(j->R);
    BN_free(j->order); BN_free(j->half); BN_free(j->r); BN_free(j->s); BN_free(j->e); BN_free(j->w);
    BN_free(j->u1); BN_free(j->u2); BN_free(j->x);
    BN_CTX_free(j->ctx); EC_GROUP_free(j->group);
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
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    job_t *jobs = (job_t *)calloc(nthreads, sizeof(job_t));
    for (int t = 0; t < nthreads; t++) {
        job_t jb;
        memset(&jb, 0, sizeof jb);
        jb.keys_xy = keys_xy; jb.K = K; jb.key_idx = key_idx; jb.digests = digests; jb.dig_off = dig_off; jb.sigs = sigs; jb.sig_off = sig_off;
        jb.begin = (int)((long long)n * t / nthreads); jb.end = (int)((long long)n * (t + 1) / nthreads); jb.status = status;
        jobs[t] = jb;
        if (nthreads == 1) worker(&jobs[t]);
        else pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}
