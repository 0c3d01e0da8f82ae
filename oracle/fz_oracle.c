/*
 * fz_oracle.c -- CPU ORACLE for the Fusion hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's algorithm
 * (geometry-labs/fusion-cryptography, pure Python).  It exists so that the
 * CUDA product can be checked bit-for-bit on inputs far larger than the
 * Python reference could process.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (fusion_cryptography_b200/) never imports, links or executes this file.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py runs the unmodified
 * Python reference in the build container and the oracle is checked against
 * those fixtures and against the reference's valid KAT CSVs
 * (tests/test_oracle_golden.py).
 *
 * Third-party arithmetic the reference takes from CPython's stdlib is restated
 * from the published algorithms: Keccak-f[1600]/SHA3-256/SHAKE256 (FIPS 202,
 * reference call sites fusion/fusion.py:408,419,591) and MT19937 with CPython's
 * seeding and randrange (reference call sites algebra/polynomials.py:448-458,
 * 479-480).  Both are additionally pinned against hashlib/random in the tests.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).
 */
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FZO_Q 2147465729LL /* fusion/fusion.py:17 */
#define FZO_HALF (FZO_Q / 2)
#define FZO_MAXD 256

typedef struct {
    int secpar, d, ell, capacity;
    int omega_ch, omega_ag, beta_ch, beta_ag, beta_sk, omega_sk;
    int bytes_coef_ch, bytes_coef_ag, bytes_idx; /* fusion.py:123-141, 433-434 */
    int64_t beta_vf;
    int64_t root, inv_root, n_inv;
    int root_order;
    int64_t tw[FZO_MAXD], itw[FZO_MAXD];
    int ready;
} fzo_params;

static fzo_params g_p128, g_p256;
static int g_threads = 1;

void fzo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int fzo_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* floor-mod like Python's % (algebra/ntt.py:120) */
static inline int64_t pymod(int64_t v, int64_t m) {
    int64_t r = v % m;
    return r < 0 ? r + m : r;
}

/* algebra/ntt.py:93-123: y = val % q; z = y - (1 + ((y - q//2 - 1) >> logmod)) * q.
 * For 0 <= y < q the shifted term is -1 when y <= q//2 and 0 otherwise, so the
 * result is y when y <= q//2 and y - q otherwise (independent of logmod 30/31). */
int64_t fzo_cent(int64_t val) {
    int64_t y = pymod(val, FZO_Q);
    int64_t iv = y - FZO_HALF - 1;
    int64_t sh = iv >> 31; /* arithmetic shift, as Python's >> on a negative int */
    return y - (1 + sh) * FZO_Q;
}

static int64_t powmod(int64_t b, int64_t e, int64_t m) {
    __int128 r = 1, x = b % m;
    while (e > 0) {
        if (e & 1) r = (r * x) % m;
        x = (x * x) % m;
        e >>= 1;
    }
    return (int64_t)r;
}

static int brv(int i, int bits) {
    int r = 0;
    for (int b = 0; b < bits; b++)
        if (i & (1 << b)) r |= 1 << (bits - 1 - b);
    return r;
}

static int ilog2(int n) {
    int k = 0;
    while ((1 << k) < n) k++;
    return k;
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

/* fusion/fusion.py:16-141 (constants), algebra/polynomials.py:396-397,414-417
 * (root powers), algebra/ntt.py:74-90 (bit_reverse_copy). */
static fzo_params *get_params(int secpar) {
    fzo_params *p = secpar == 128 ? &g_p128 : secpar == 256 ? &g_p256 : NULL;
    if (!p) return NULL;
    if (p->ready) return p;
    p->secpar = secpar;
    if (secpar == 128) {
        p->d = 64; p->ell = 195; p->capacity = 1796; p->omega_ch = 27; p->omega_ag = 35;
        p->root = 23584283;
    } else {
        p->d = 256; p->ell = 83; p->capacity = 2818; p->omega_ch = 60; p->omega_ag = 60;
        p->root = 3337519;
    }
    p->beta_ch = 1; p->beta_ag = 1; p->beta_sk = 52; p->omega_sk = p->d;
    p->root_order = 2 * p->d;
    p->inv_root = powmod(p->root, FZO_Q - 2, FZO_Q);
    p->n_inv = powmod(p->d, FZO_Q - 2, FZO_Q); /* algebra/ntt.py:353 */
    {   /* fusion.py:57-68 */
        int ch_bd = secpar == 128 ? 3 : 1, ag_bd = secpar == 128 ? 2 : 1;
        int64_t inter = 52LL * (1 + (int64_t)(p->d < p->omega_ch ? p->d : p->omega_ch) * ch_bd);
        p->beta_vf = (int64_t)p->capacity * (p->d < p->omega_ag ? p->d : p->omega_ag) * ag_bd * inter;
    }
    /* fusion.py:433-434 with bound = 1 (log2(1) = 0): ceil((0 + 1 + secpar)/8) */
    p->bytes_coef_ch = ceil_div(1 + secpar, 8);
    p->bytes_coef_ag = ceil_div(1 + secpar, 8);
    p->bytes_idx = ceil_div(ilog2(p->d) + secpar, 8);
    int bits = ilog2(p->d);
    int64_t pw = 1, ipw = 1;
    int64_t fw[FZO_MAXD], iw[FZO_MAXD];
    for (int i = 0; i < p->d; i++) {
        fw[i] = pw; iw[i] = ipw;
        pw = (int64_t)(((__int128)pw * p->root) % FZO_Q);
        ipw = (int64_t)(((__int128)ipw * p->inv_root) % FZO_Q);
    }
    for (int i = 0; i < p->d; i++) {
        p->tw[i] = fw[brv(i, bits)];
        p->itw[i] = iw[brv(i, bits)];
    }
    p->ready = 1;
    return p;
}

int fzo_param(int secpar, const char *name, int64_t *out) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
#define P(n) if (!strcmp(name, #n)) { *out = (int64_t)p->n; return 0; }
    P(d) P(ell) P(capacity) P(omega_ch) P(omega_ag) P(beta_ch) P(beta_ag) P(beta_sk)
    P(bytes_coef_ch) P(bytes_coef_ag) P(bytes_idx) P(beta_vf) P(root) P(inv_root) P(n_inv) P(root_order)
#undef P
    return -2;
}

int fzo_twiddles(int secpar, int64_t *tw, int64_t *itw) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    memcpy(tw, p->tw, sizeof(int64_t) * p->d);
    memcpy(itw, p->itw, sizeof(int64_t) * p->d);
    return 0;
}

/* algebra/ntt.py:271-291 (cooley_tukey_ntt main loop), one polynomial in place. */
static void ntt_fwd_one(const fzo_params *p, int32_t *val) {
    int n = p->d, t = n, m = 1;
    while (m < n) {
        t /= 2;
        for (int i = 0; i < m; i++) {
            int j1 = 2 * i * t, j2 = j1 + t - 1;
            int64_t s = p->tw[m + i];
            for (int j = j1; j <= j2; j++) {
                int64_t u = val[j], v = (int64_t)val[j + t] * s;
                val[j] = (int32_t)fzo_cent(u + v);
                val[j + t] = (int32_t)fzo_cent(u - v);
            }
        }
        m *= 2;
    }
}

/* algebra/ntt.py:350-377 (gentleman_sande_intt main loop + n^-1 scaling). */
static void ntt_inv_one(const fzo_params *p, int32_t *val) {
    int n = p->d, t = 1, m = n;
    while (m > 1) {
        int j1 = 0, h = m / 2;
        for (int i = 0; i < h; i++) {
            int j2 = j1 + t - 1;
            int64_t s = p->itw[h + i];
            for (int j = j1; j <= j2; j++) {
                int64_t u = val[j], v = val[j + t];
                val[j] = (int32_t)fzo_cent(u + v);
                val[j + t] = (int32_t)fzo_cent((u - v) * s);
            }
            j1 += 2 * t;
        }
        t *= 2;
        m = h;
    }
    for (int j = 0; j < n; j++) val[j] = (int32_t)fzo_cent((int64_t)val[j] * p->n_inv);
}

/* algebra/polynomials.py:394-411 (transform, coefficient -> NTT), batch of n polys. */
int fzo_ntt_fwd(int secpar, int32_t *polys, size_t n) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long long i = 0; i < (long long)n; i++) ntt_fwd_one(p, polys + (size_t)i * p->d);
    return 0;
}

/* algebra/polynomials.py:412-431 (transform, NTT -> coefficient). */
int fzo_ntt_inv(int secpar, int32_t *polys, size_t n) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long long i = 0; i < (long long)n; i++) ntt_inv_one(p, polys + (size_t)i * p->d);
    return 0;
}

/* algebra/polynomials.py:341-385 (__mul__, pointwise cent(x*y)). */
int fzo_pointwise_mul(const int32_t *a, const int32_t *b, int32_t *out, size_t n_coef) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long long i = 0; i < (long long)n_coef; i++)
        out[i] = (int32_t)fzo_cent((int64_t)a[i] * b[i]);
    return 0;
}

/* __mul__ then __add__ (polynomials.py:341-385, 282-318): cent(cent(x*y) + c). */
int fzo_pointwise_muladd(const int32_t *a, const int32_t *b, const int32_t *c, int32_t *out, size_t n_coef) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long long i = 0; i < (long long)n_coef; i++)
        out[i] = (int32_t)fzo_cent(fzo_cent((int64_t)a[i] * b[i]) + c[i]);
    return 0;
}

/* INTT(NTT(a) o NTT(b)): algebra/ntt.py:450-471 (ntt_poly_mult, result only). */
int fzo_polymul(int secpar, const int32_t *a, const int32_t *b, int32_t *out, size_t n) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long long i = 0; i < (long long)n; i++) {
        int32_t fa[FZO_MAXD], fb[FZO_MAXD];
        memcpy(fa, a + (size_t)i * p->d, sizeof(int32_t) * p->d);
        memcpy(fb, b + (size_t)i * p->d, sizeof(int32_t) * p->d);
        ntt_fwd_one(p, fa);
        ntt_fwd_one(p, fb);
        for (int k = 0; k < p->d; k++) fa[k] = (int32_t)fzo_cent((int64_t)fa[k] * fb[k]);
        ntt_inv_one(p, fa);
        memcpy(out + (size_t)i * p->d, fa, sizeof(int32_t) * p->d);
    }
    return 0;
}

/* ---- any ring Z_q[X]/(X^d + 1), q an odd prime below 2^31, caller-supplied tables ---------------------------------
 * The same loops as above with the modulus, the degree and the bit-reversed power table as arguments, which is the
 * signature the reference's functions really have (algebra/ntt.py:216-219, :294-297); the reference's own unit tests
 * call them on small moduli (tests/test_ntt.py:18-50). */

/* algebra/ntt.py:93-123 with halfmod = q // 2: for 0 <= y < q the shifted term is -1 up to q//2 and 0 above, for any
 * logmod with 2^logmod >= q/2 (callers pass ceil(log2 q), bit_length or bit_length - 1: ntt.py:272, :448, polynomials.py:61) */
int64_t fzo_ring_cent(int64_t val, int64_t q) {
    int64_t y = pymod(val, q);
    return y > q / 2 ? y - q : y;
}

/* algebra/ntt.py:271-291 */
int fzo_ring_ntt_fwd(int64_t q, int d, const int64_t *table, int32_t *polys, size_t n) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long long pi = 0; pi < (long long)n; pi++) {
        int32_t *val = polys + (size_t)pi * d;
        int t = d, m = 1;
        while (m < d) {
            t /= 2;
            for (int i = 0; i < m; i++) {
                int j1 = 2 * i * t, j2 = j1 + t - 1;
                int64_t s = pymod(table[m + i], q);
                for (int j = j1; j <= j2; j++) {
                    int64_t u = val[j], v = pymod((int64_t)val[j + t], q) * s;
                    val[j] = (int32_t)fzo_ring_cent(u + v, q);
                    val[j + t] = (int32_t)fzo_ring_cent(u - v, q);
                }
            }
            m *= 2;
        }
    }
    return 0;
}

/* algebra/ntt.py:350-377, n_inv = pow(n, q - 2, q) (:353) */
int fzo_ring_ntt_inv(int64_t q, int d, const int64_t *table, int32_t *polys, size_t n) {
    const int64_t n_inv = powmod(d, q - 2, q);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long long pi = 0; pi < (long long)n; pi++) {
        int32_t *val = polys + (size_t)pi * d;
        int t = 1, m = d;
        while (m > 1) {
            int j1 = 0, h = m / 2;
            for (int i = 0; i < h; i++) {
                int j2 = j1 + t - 1;
                int64_t s = pymod(table[h + i], q);
                for (int j = j1; j <= j2; j++) {
                    int64_t u = val[j], v = val[j + t];
                    val[j] = (int32_t)fzo_ring_cent(u + v, q);
                    val[j + t] = (int32_t)fzo_ring_cent(pymod(u - v, q) * s, q);
                }
                j1 += 2 * t;
            }
            t *= 2;
            m = h;
        }
        for (int j = 0; j < d; j++) val[j] = (int32_t)fzo_ring_cent(pymod(val[j], q) * n_inv, q);
    }
    return 0;
}

/* algebra/polynomials.py:341-385 / :282-318: cent(x * y) (mul = 1) or cent(x + y) (mul = 0) */
int fzo_ring_pointwise(int64_t q, int mul, const int32_t *a, const int32_t *b, int32_t *out, size_t n_coef) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long long i = 0; i < (long long)n_coef; i++)
        out[i] = (int32_t)fzo_ring_cent(mul ? pymod(a[i], q) * pymod(b[i], q) : (int64_t)a[i] + b[i], q);
    return 0;
}

/* algebra/matrices.py:121-131: (1 x ell) . (ell x 1), accumulated left to right
 * with cent after every mul and every add.  A: [ell][d]; s: [batch][ell][d]; out: [batch][d]. */
static void matvec_one(const fzo_params *p, const int32_t *A, const int32_t *s, int32_t *out) {
    for (int k = 0; k < p->d; k++) {
        int64_t acc = fzo_cent((int64_t)A[k] * s[k]);
        for (int r = 1; r < p->ell; r++)
            acc = fzo_cent(acc + fzo_cent((int64_t)A[(size_t)r * p->d + k] * s[(size_t)r * p->d + k]));
        out[k] = (int32_t)acc;
    }
}

int fzo_matvec(int secpar, const int32_t *A, const int32_t *s, int32_t *out, size_t batch) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long long b = 0; b < (long long)batch; b++)
        matvec_one(p, A, s + (size_t)b * p->ell * p->d, out + (size_t)b * p->d);
    return 0;
}

/* ------------------------------------------------------------------ MT19937
 * CPython Modules/_randommodule.c (init_genrand, init_by_array, genrand_uint32)
 * and Lib/random.py (_randbelow_with_getrandbits); the reference reaches them
 * through random.seed / randrange at algebra/polynomials.py:448,452,458,479-480. */
typedef struct { uint32_t mt[624]; int idx; } fzo_mt;

static void mt_init_genrand(fzo_mt *s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        s->mt[i] = 1812433253U * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}

static void mt_init_by_array(fzo_mt *s, const uint32_t *key, int klen) {
    mt_init_genrand(s, 19650218U);
    int i = 1, j = 0;
    int k = 624 > klen ? 624 : klen;
    for (; k; k--) {
        s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1664525U)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
        if (j >= klen) j = 0;
    }
    for (k = 623; k; k--) {
        s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1566083941U)) - (uint32_t)i;
        i++;
        if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
    }
    s->mt[0] = 0x80000000U;
    s->idx = 624;
}

static uint32_t mt_next(fzo_mt *s) {
    if (s->idx >= 624) {
        uint32_t *mt = s->mt;
        int kk;
        for (kk = 0; kk < 624; kk++) {
            uint32_t y = (mt[kk] & 0x80000000U) | (mt[(kk + 1) % 624] & 0x7fffffffU);
            mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680U;
    y ^= (y << 15) & 0xefc60000U;
    y ^= y >> 18;
    return y;
}

static int bit_length_u32(uint32_t n) {
    int k = 0;
    while (n) { k++; n >>= 1; }
    return k;
}

/* Lib/random.py _randbelow_with_getrandbits; n >= 1, n < 2^32. */
static uint32_t mt_randbelow(fzo_mt *s, uint32_t n) {
    int k = bit_length_u32(n);
    uint32_t r = mt_next(s) >> (32 - k);
    while (r >= n) r = mt_next(s) >> (32 - k);
    return r;
}

/* raw stream, for pinning against random.getrandbits(32) */
int fzo_mt_stream(const uint32_t *key, int klen, uint32_t *out, int n) {
    fzo_mt s;
    mt_init_by_array(&s, key, klen);
    for (int i = 0; i < n; i++) out[i] = mt_next(&s);
    return 0;
}

/* algebra/polynomials.py:436-467 with weight_bound = omega_sk = degree (no shuffle branch
 * is taken when weight == degree, :455); general weight handled for completeness. */
int fzo_sample_coef_poly(int secpar, const uint32_t *key, int klen, int norm_bound, int weight_bound, int32_t *out) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    fzo_mt s;
    mt_init_by_array(&s, key, klen);
    int ncoef = weight_bound < p->d ? weight_bound : p->d;
    if (ncoef < 0) ncoef = 0;
    int64_t bound = norm_bound < FZO_HALF ? norm_bound : FZO_HALF;
    if (bound < 0) bound = 0;
    for (int i = 0; i < ncoef; i++) {
        int32_t mag = 1 + (int32_t)mt_randbelow(&s, (uint32_t)bound);
        int32_t sg = 1 - 2 * (int32_t)mt_randbelow(&s, 2);
        out[i] = mag * sg;
    }
    for (int i = ncoef; i < p->d; i++) out[i] = 0;
    if (ncoef < p->d) {
        for (int i = p->d - 1; i > 0; i--) {
            int j = (int)mt_randbelow(&s, (uint32_t)(i + 1));
            int32_t t = out[i]; out[i] = out[j]; out[j] = t;
        }
    }
    return 0;
}

/* algebra/polynomials.py:470-488: values = randrange(q) - q//2. */
int fzo_sample_ntt_poly(int secpar, const uint32_t *key, int klen, int32_t *out) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    fzo_mt s;
    mt_init_by_array(&s, key, klen);
    for (int i = 0; i < p->d; i++)
        out[i] = (int32_t)((int64_t)mt_randbelow(&s, (uint32_t)FZO_Q) - FZO_HALF);
    return 0;
}

/* ------------------------------------------------------------------ Keccak (FIPS 202) */
static const uint64_t KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

static inline uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

static void keccak_f(uint64_t a[25]) {
    for (int r = 0; r < 24; r++) {
        uint64_t c[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) {
            uint64_t dd = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
            for (int y = 0; y < 25; y += 5) a[y + x] ^= dd;
        }
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++)
                b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(a[x + 5 * y], KROT[x + 5 * y]);
        for (int y = 0; y < 25; y += 5)
            for (int x = 0; x < 5; x++) a[y + x] = b[y + x] ^ (~b[y + (x + 1) % 5] & b[y + (x + 2) % 5]);
        a[0] ^= KRC[r];
    }
}

typedef struct { uint64_t a[25]; int pos; int squeezing; uint8_t suffix; } fzo_sponge; /* rate 136 */

static void sp_init(fzo_sponge *s, uint8_t suffix) { memset(s, 0, sizeof(*s)); s->suffix = suffix; }
static void sp_absorb(fzo_sponge *s, const uint8_t *in, size_t len) {
    for (size_t i = 0; i < len; i++) {
        s->a[s->pos >> 3] ^= (uint64_t)in[i] << (8 * (s->pos & 7));
        if (++s->pos == 136) { keccak_f(s->a); s->pos = 0; }
    }
}
static void sp_finish(fzo_sponge *s) {
    s->a[s->pos >> 3] ^= (uint64_t)s->suffix << (8 * (s->pos & 7));
    s->a[16] ^= 0x8000000000000000ULL; /* byte 135 ^= 0x80 */
    keccak_f(s->a);
    s->pos = 0;
    s->squeezing = 1;
}
static void sp_squeeze(fzo_sponge *s, uint8_t *out, size_t len) {
    if (!s->squeezing) sp_finish(s);
    for (size_t i = 0; i < len; i++) {
        if (s->pos == 136) { keccak_f(s->a); s->pos = 0; }
        out[i] = (uint8_t)(s->a[s->pos >> 3] >> (8 * (s->pos & 7)));
        s->pos++;
    }
}

int fzo_sha3_256(const uint8_t *in, size_t len, uint8_t *out32) {
    fzo_sponge s; sp_init(&s, 0x06); sp_absorb(&s, in, len); sp_squeeze(&s, out32, 32); return 0;
}
int fzo_shake256(const uint8_t *in, size_t len, uint8_t *out, size_t outlen) {
    fzo_sponge s; sp_init(&s, 0x1F); sp_absorb(&s, in, len); sp_squeeze(&s, out, outlen); return 0;
}

/* ------------------------------------------------------------------ repr formatting
 * algebra/polynomials.py:257-258, algebra/matrices.py:40-41, fusion/fusion.py:328-329, 382-383. */
static size_t fmt_poly(const fzo_params *p, const int32_t *v, char *out) {
    char *o = out;
    o += sprintf(o, "PolynomialNTTRepresentation(modulus=%lld, degree=%d, root=%lld, inv_root=%lld, root_order=%d, values=[",
                 (long long)FZO_Q, p->d, (long long)p->root, (long long)p->inv_root, p->root_order);
    for (int i = 0; i < p->d; i++) o += sprintf(o, i ? ", %d" : "%d", v[i]);
    o += sprintf(o, "])");
    return (size_t)(o - out);
}

static size_t fmt_gm1(const fzo_params *p, const int32_t *v, char *out) {
    char *o = out;
    o += sprintf(o, "GeneralMatrix(elem_class=<class 'algebra.polynomials.PolynomialNTTRepresentation'>, matrix=[[");
    o += fmt_poly(p, v, o);
    o += sprintf(o, "]])");
    return (size_t)(o - out);
}

static size_t fmt_vk(const fzo_params *p, const int32_t *vl, const int32_t *vr, char *out) {
    char *o = out;
    o += sprintf(o, "OneTimeVerificationKey(left_vk_hat=");
    o += fmt_gm1(p, vl, o);
    o += sprintf(o, ", right_vk_hat=");
    o += fmt_gm1(p, vr, o);
    o += sprintf(o, ")");
    return (size_t)(o - out);
}

static size_t fmt_ch(const fzo_params *p, const int32_t *c, char *out) {
    char *o = out;
    o += sprintf(o, "SignatureChallenge(c_hat=");
    o += fmt_poly(p, c, o);
    o += sprintf(o, ")");
    return (size_t)(o - out);
}

/* decimal digits of int.from_bytes(digest, "little") (fusion.py:409, then str() at :417, :589) */
static size_t fmt_le256(const uint8_t *dg, char *out) {
    uint32_t limb[8];
    for (int i = 0; i < 8; i++)
        limb[i] = (uint32_t)dg[4 * i] | ((uint32_t)dg[4 * i + 1] << 8) | ((uint32_t)dg[4 * i + 2] << 16) | ((uint32_t)dg[4 * i + 3] << 24);
    char tmp[96];
    int n = 0;
    for (;;) {
        int nz = 0;
        for (int i = 0; i < 8; i++) nz |= limb[i] != 0;
        if (!nz) break;
        uint64_t rem = 0;
        for (int i = 7; i >= 0; i--) {
            uint64_t cur = (rem << 32) | limb[i];
            limb[i] = (uint32_t)(cur / 10);
            rem = cur % 10;
        }
        tmp[n++] = (char)('0' + rem);
    }
    if (!n) tmp[n++] = '0';
    for (int i = 0; i < n; i++) out[i] = tmp[n - 1 - i];
    return (size_t)n;
}

#define FZO_VK_MAX 8192
#define FZO_PIECE_MAX 12288

size_t fzo_format_vk(int secpar, const int32_t *vl, const int32_t *vr, char *out) {
    return fmt_vk(get_params(secpar), vl, vr, out);
}
size_t fzo_format_ch(int secpar, const int32_t *c, char *out) { return fmt_ch(get_params(secpar), c, out); }
size_t fzo_format_prehash(const uint8_t *dg, char *out) { return fmt_le256(dg, out); }

/* fusion/fusion.py:405-409: sha3_256(dst + "," + message); dst = (3 or 1, 0) (:38-43). */
int fzo_hash_msg(int secpar, const uint8_t *msg, size_t len, uint8_t *out32) {
    fzo_sponge s;
    uint8_t pre[3] = {(uint8_t)(secpar == 256 ? 3 : 1), 0, ','};
    sp_init(&s, 0x06);
    sp_absorb(&s, pre, 3);
    sp_absorb(&s, msg, len);
    sp_squeeze(&s, out32, 32);
    return 0;
}

/* fusion/fusion.py:422-481.  Returns 0, or -1 for "Too few bytes" (:439-442).
 * Python slicing past the end of a bytes object yields a shorter (possibly empty)
 * slice and int.from_bytes(b"") == 0; both are reproduced. */
int fzo_decode_poly(int secpar, const uint8_t *b, size_t blen, int log2_bias, int norm_bound, int weight_bound, int32_t *coefs) {
    fzo_params *p = get_params(secpar);
    if (!p) return -2;
    int d = p->d;
    int num_coefs = weight_bound < d ? weight_bound : d;
    if (num_coefs < 1) num_coefs = 1;
    int64_t bound = norm_bound < FZO_HALF ? norm_bound : FZO_HALF;
    if (bound < 1) bound = 1;
    /* ceil((log2(bound) + 1 + bias)/8): exact for bound a power of two; otherwise
     * floor(log2(bound)) + 1 more bit when bound is not a power of two. */
    int lb = 0;
    while ((1LL << (lb + 1)) <= bound) lb++;
    int frac = ((1LL << lb) != bound);
    int bpc = ceil_div(lb + frac + 1 + log2_bias, 8);
    int bpi = ceil_div(ilog2(d) + log2_bias, 8);
    int bfs = ceil_div(weight_bound, 8);
    size_t total = (size_t)bfs + (size_t)(bpc + bpi) * weight_bound;
    if (blen < total) return -1;
    size_t pos = 0;
    /* signums (:446-453): bit i, LSB first, of the big-endian integer of b[:bfs] */
    for (int i = 0; i < d; i++) coefs[i] = 0;
    int sg[FZO_MAXD];
    for (int i = 0; i < weight_bound && i < FZO_MAXD; i++) {
        int bit = (b[bfs - 1 - i / 8] >> (i % 8)) & 1;
        sg[i] = 2 * bit - 1;
    }
    pos = bfs;
    /* magnitudes (:458-465): (int(next bpc bytes, big endian) % bound + 1) * signum */
    for (int i = 0; i < weight_bound; i++) {
        uint64_t r = 0;
        for (int k = 0; k < bpc; k++) {
            if (pos < blen) { r = ((r << 8) | b[pos]) % (uint64_t)bound; pos++; }
        }
        if (i < d) coefs[i] = (int32_t)((int64_t)(r + 1) * sg[i]);
    }
    /* partial Fisher-Yates (:472-480) */
    if (num_coefs < d) {
        for (int i = d - 1; i > weight_bound; i--) {
            uint64_t r = 0;
            uint64_t m = (uint64_t)(i + 1);
            for (int k = 0; k < bpi; k++) {
                if (pos < blen) { r = ((r << 8) | b[pos]) % m; pos++; }
            }
            int j = (int)r;
            int32_t t = coefs[i]; coefs[i] = coefs[j]; coefs[j] = t;
        }
    }
    return 0;
}

static int ch_out_len(const fzo_params *p) {
    /* fusion.py:515-524: n = bfs + bpc*num_coefs + degree*bpi */
    int num_coefs = p->d < p->omega_ch ? p->d : p->omega_ch;
    return ceil_div(p->omega_ch, 8) + p->bytes_coef_ch * num_coefs + p->d * p->bytes_idx;
}
static int ag_out_len(const fzo_params *p) {
    /* fusion.py:579-585 */
    return ceil_div(p->omega_ag, 8) + (p->bytes_coef_ag + p->bytes_idx) * p->omega_ag;
}

/* fusion.py:412-419: shake_256(dst + "," + str(vk) + "," + str(i)).digest(n) */
int fzo_hash_vk_and_int_to_bytes(int secpar, const int32_t *vl, const int32_t *vr, const uint8_t *digest32, uint8_t *out, size_t n) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    char *buf = (char *)malloc(FZO_VK_MAX + 128);
    size_t o = 0;
    buf[o++] = (char)(secpar == 256 ? 3 : 1); buf[o++] = 1; buf[o++] = ',';
    o += fmt_vk(p, vl, vr, buf + o);
    buf[o++] = ',';
    o += fmt_le256(digest32, buf + o);
    fzo_shake256((const uint8_t *)buf, o, out, n);
    free(buf);
    return 0;
}

/* fusion.py:511-531 (hash_ch) = :405-409 + :412-419 + :484-508 (parse_challenge: decode + NTT).
 * c_coef (optional) receives the coefficient-domain challenge. */
int fzo_hash_ch(int secpar, const int32_t *vl, const int32_t *vr, const uint8_t *msg, size_t mlen, int32_t *c_hat, int32_t *c_coef, uint8_t *digest_out) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    uint8_t dg[32];
    fzo_hash_msg(secpar, msg, mlen, dg);
    if (digest_out) memcpy(digest_out, dg, 32);
    int n = ch_out_len(p);
    uint8_t *xof = (uint8_t *)malloc(n);
    fzo_hash_vk_and_int_to_bytes(secpar, vl, vr, dg, xof, n);
    int rc = fzo_decode_poly(secpar, xof, n, secpar, p->beta_ch, p->omega_ch, c_hat);
    free(xof);
    if (rc) return rc;
    if (c_coef) memcpy(c_coef, c_hat, sizeof(int32_t) * p->d);
    ntt_fwd_one(p, c_hat);
    return 0;
}

/* limbs of abs(seed) for u64 seeds (CPython random_seed: 32-bit little-endian limbs, >= 1 limb) */
static int u64_limbs(uint64_t s, uint32_t *key) {
    key[0] = (uint32_t)s; key[1] = (uint32_t)(s >> 32);
    return key[1] ? 2 : 1;
}

/* fusion.py:176-201 + :273-282: A is 1 x ell, every entry sampled with the SAME seed
 *  => ell identical polynomials.  A out: [ell][d]. */
int fzo_setup_A(int secpar, const uint32_t *key, int klen, int32_t *A) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    fzo_sample_ntt_poly(secpar, key, klen, A);
    for (int r = 1; r < p->ell; r++) memcpy(A + (size_t)r * p->d, A, sizeof(int32_t) * p->d);
    return 0;
}

/* fusion.py:338-373.  keyL/keyR = limbs of abs(seed), abs(seed+1).  sk_*: [ell][d], vk_*: [d]. */
int fzo_keygen(int secpar, const int32_t *A, const uint32_t *keyL, int klenL, const uint32_t *keyR, int klenR,
               int32_t *sk_l, int32_t *sk_r, int32_t *vk_l, int32_t *vk_r) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    for (int side = 0; side < 2; side++) {
        int32_t *sk = side ? sk_r : sk_l;
        for (int r = 0; r < p->ell; r++) { /* every row re-seeds with the same seed (Q1) */
            fzo_sample_coef_poly(secpar, side ? keyR : keyL, side ? klenR : klenL, p->beta_sk, p->omega_sk, sk + (size_t)r * p->d);
            ntt_fwd_one(p, sk + (size_t)r * p->d);
        }
        matvec_one(p, A, sk, side ? vk_r : vk_l);
    }
    return 0;
}

/* batch keygen with u64 seeds (seed_i, seed_i + 1).  sk_l/sk_r may be NULL. */
int fzo_keygen_batch(int secpar, const int32_t *A, const uint64_t *seeds, size_t batch,
                     int32_t *sk_l, int32_t *sk_r, int32_t *vk_l, int32_t *vk_r) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    size_t skn = (size_t)p->ell * p->d;
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 1)
    for (long long b = 0; b < (long long)batch; b++) {
        uint32_t kL[3], kR[3];
        int nL = u64_limbs(seeds[b], kL), nR;
        if (seeds[b] == UINT64_MAX) { kR[0] = 0; kR[1] = 0; kR[2] = 1; nR = 3; }
        else nR = u64_limbs(seeds[b] + 1, kR);
        int32_t *tl = sk_l ? sk_l + (size_t)b * skn : (int32_t *)malloc(skn * 4);
        int32_t *tr = sk_r ? sk_r + (size_t)b * skn : (int32_t *)malloc(skn * 4);
        fzo_keygen(secpar, A, kL, nL, kR, nR, tl, tr, vk_l + (size_t)b * p->d, vk_r + (size_t)b * p->d);
        if (!sk_l) free(tl);
        if (!sk_r) free(tr);
    }
    return 0;
}

/* fusion.py:534-557: z = sk_l * c_hat + sk_r  (matrices.py:109-114 then :74-91). */
int fzo_sign(int secpar, const int32_t *sk_l, const int32_t *sk_r, const int32_t *vk_l, const int32_t *vk_r,
             const uint8_t *msg, size_t mlen, int32_t *sig) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    int32_t c_hat[FZO_MAXD];
    int rc = fzo_hash_ch(secpar, vk_l, vk_r, msg, mlen, c_hat, NULL, NULL);
    if (rc) return rc;
    for (int r = 0; r < p->ell; r++)
        for (int k = 0; k < p->d; k++) {
            size_t o = (size_t)r * p->d + k;
            sig[o] = (int32_t)fzo_cent(fzo_cent((int64_t)sk_l[o] * c_hat[k]) + sk_r[o]);
        }
    return 0;
}

int fzo_sign_batch(int secpar, const int32_t *sk_l, const int32_t *sk_r, const int32_t *vk_l, const int32_t *vk_r,
                   const uint8_t *msgs, const uint64_t *offs, size_t batch, int32_t *sigs) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    size_t skn = (size_t)p->ell * p->d;
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 1)
    for (long long b = 0; b < (long long)batch; b++)
        fzo_sign(secpar, sk_l + b * skn, sk_r + b * skn, vk_l + (size_t)b * p->d, vk_r + (size_t)b * p->d,
                 msgs + offs[b], (size_t)(offs[b + 1] - offs[b]), sigs + b * skn);
    return 0;
}

/* sort key of fusion.py:661-663 / :693: str(vk), Python str comparison = memcmp, shorter first; stable. */
typedef struct { char *s; size_t len; size_t idx; } fzo_sortrec;
static int cmp_rec(const void *a, const void *b) {
    const fzo_sortrec *x = (const fzo_sortrec *)a, *y = (const fzo_sortrec *)b;
    size_t m = x->len < y->len ? x->len : y->len;
    int c = memcmp(x->s, y->s, m);
    if (c) return c;
    if (x->len != y->len) return x->len < y->len ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0); /* stability */
}

int fzo_sort_by_repr(int secpar, const int32_t *vk_l, const int32_t *vk_r, size_t n, uint32_t *order) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    fzo_sortrec *rec = (fzo_sortrec *)malloc(sizeof(fzo_sortrec) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) {
        rec[i].s = (char *)malloc(FZO_VK_MAX);
        rec[i].len = fmt_vk(p, vk_l + i * p->d, vk_r + i * p->d, rec[i].s);
        rec[i].idx = i;
    }
    qsort(rec, n, sizeof(fzo_sortrec), cmp_rec);
    for (size_t i = 0; i < n; i++) { order[i] = (uint32_t)rec[i].idx; free(rec[i].s); }
    free(rec);
    return 0;
}

/* fusion.py:632-652 (hash_ag) on ALREADY SORTED inputs: prehash each message, hash_ch each key,
 * one SHAKE256 over dst + "," + str(list(zip(keys, prehashes, challenges))) (:573-591),
 * slice and decode (:594-629).  alpha_hat: [n][d]; c_hat_out (optional): [n][d]. */
int fzo_hash_ag(int secpar, const int32_t *vk_l, const int32_t *vk_r, const uint8_t *msgs, const uint64_t *offs,
                size_t n, int32_t *alpha_hat, int32_t *c_hat_out) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    int d = p->d;
    int32_t *ch = c_hat_out ? c_hat_out : (int32_t *)malloc(sizeof(int32_t) * d * (n ? n : 1));
    uint8_t *dgs = (uint8_t *)malloc(32 * (n ? n : 1));
    int bad = 0;
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 1)
    for (long long i = 0; i < (long long)n; i++) {
        if (fzo_hash_ch(secpar, vk_l + (size_t)i * d, vk_r + (size_t)i * d, msgs + offs[i], (size_t)(offs[i + 1] - offs[i]),
                        ch + (size_t)i * d, NULL, dgs + 32 * i))
            bad = 1;
    }
    fzo_sponge s;
    sp_init(&s, 0x1F);
    uint8_t pre[3] = {(uint8_t)(secpar == 256 ? 3 : 1), 2, ','};
    sp_absorb(&s, pre, 3);
    sp_absorb(&s, (const uint8_t *)"[", 1);
    char *piece = (char *)malloc(FZO_PIECE_MAX);
    for (size_t i = 0; i < n; i++) {
        size_t o = 0;
        if (i) { piece[o++] = ','; piece[o++] = ' '; }
        piece[o++] = '(';
        o += fmt_vk(p, vk_l + i * d, vk_r + i * d, piece + o);
        piece[o++] = ','; piece[o++] = ' ';
        o += fmt_le256(dgs + 32 * i, piece + o);
        piece[o++] = ','; piece[o++] = ' ';
        o += fmt_ch(p, ch + i * d, piece + o);
        piece[o++] = ')';
        sp_absorb(&s, (const uint8_t *)piece, o);
    }
    sp_absorb(&s, (const uint8_t *)"]", 1);
    free(piece);
    int nag = ag_out_len(p);
    uint8_t *xof = (uint8_t *)malloc((size_t)nag * (n ? n : 1));
    sp_squeeze(&s, xof, (size_t)nag * n);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long long i = 0; i < (long long)n; i++) {
        if (fzo_decode_poly(secpar, xof + (size_t)i * nag, nag, secpar, p->beta_ag, p->omega_ag, alpha_hat + (size_t)i * d)) bad = 1;
        ntt_fwd_one(p, alpha_hat + (size_t)i * d);
    }
    free(xof); free(dgs);
    if (!c_hat_out) free(ch);
    return bad ? -3 : 0;
}

static void gather_sorted(const fzo_params *p, const uint32_t *order, size_t n, const int32_t *vk_l, const int32_t *vk_r,
                          const uint8_t *msgs, const uint64_t *offs, int32_t *svl, int32_t *svr, uint8_t *smsgs, uint64_t *soffs) {
    soffs[0] = 0;
    for (size_t i = 0; i < n; i++) {
        size_t k = order[i];
        memcpy(svl + i * p->d, vk_l + k * p->d, sizeof(int32_t) * p->d);
        memcpy(svr + i * p->d, vk_r + k * p->d, sizeof(int32_t) * p->d);
        size_t len = (size_t)(offs[k + 1] - offs[k]);
        memcpy(smsgs + soffs[i], msgs + offs[k], len);
        soffs[i + 1] = soffs[i] + len;
    }
}

/* fusion.py:655-677.  sigs: [n][ell][d]; out: [ell][d].  n >= 1. */
int fzo_aggregate(int secpar, const int32_t *vk_l, const int32_t *vk_r, const uint8_t *msgs, const uint64_t *offs,
                  const int32_t *sigs, size_t n, int32_t *out) {
    fzo_params *p = get_params(secpar);
    if (!p || n == 0) return -1;
    int d = p->d;
    size_t skn = (size_t)p->ell * d;
    uint32_t *order = (uint32_t *)malloc(4 * n);
    fzo_sort_by_repr(secpar, vk_l, vk_r, n, order);
    int32_t *svl = (int32_t *)malloc(4 * d * n), *svr = (int32_t *)malloc(4 * d * n), *alpha = (int32_t *)malloc(4 * d * n);
    uint8_t *smsgs = (uint8_t *)malloc((size_t)offs[n] + 1);
    uint64_t *soffs = (uint64_t *)malloc(8 * (n + 1));
    gather_sorted(p, order, n, vk_l, vk_r, msgs, offs, svl, svr, smsgs, soffs);
    int rc = fzo_hash_ag(secpar, svl, svr, smsgs, soffs, n, alpha, NULL);
    /* running sum, left to right (:670-677); cent after each mul and add */
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long long o = 0; o < (long long)skn; o++) {
        int k = (int)(o % d);
        int64_t acc = fzo_cent((int64_t)sigs[(size_t)order[0] * skn + o] * alpha[k]);
        for (size_t i = 1; i < n; i++)
            acc = fzo_cent(acc + fzo_cent((int64_t)sigs[(size_t)order[i] * skn + o] * alpha[i * d + k]));
        out[o] = (int32_t)acc;
    }
    free(order); free(svl); free(svr); free(alpha); free(smsgs); free(soffs);
    return rc;
}

/* fusion.py:680-728.  Return code: 0 ok; 1 too many keys; 2 count mismatch; 3 target mismatch;
 * 4 norm too large; 5 weight too large (unreachable: omega_vf = d). */
int fzo_verify(int secpar, const int32_t *A, const int32_t *vk_l, const int32_t *vk_r, const uint8_t *msgs, const uint64_t *offs,
               size_t n_keys, size_t n_msgs, const int32_t *agg) {
    fzo_params *p = get_params(secpar);
    if (!p) return -1;
    int d = p->d;
    if (n_keys > (size_t)p->capacity) return 1;
    if (n_keys != n_msgs) return 2;
    size_t n = n_keys;
    if (n == 0) return -1; /* reference raises IndexError at :704 */
    size_t skn = (size_t)p->ell * d;
    int32_t *coef = (int32_t *)malloc(4 * skn);
    memcpy(coef, agg, 4 * skn);
    for (int r = 0; r < p->ell; r++) ntt_inv_one(p, coef + (size_t)r * d);
    uint32_t *order = (uint32_t *)malloc(4 * n);
    fzo_sort_by_repr(secpar, vk_l, vk_r, n, order);
    int32_t *svl = (int32_t *)malloc(4 * d * n), *svr = (int32_t *)malloc(4 * d * n);
    int32_t *alpha = (int32_t *)malloc(4 * d * n), *ch = (int32_t *)malloc(4 * d * n);
    uint8_t *smsgs = (uint8_t *)malloc((size_t)offs[n] + 1);
    uint64_t *soffs = (uint64_t *)malloc(8 * (n + 1));
    gather_sorted(p, order, n, vk_l, vk_r, msgs, offs, svl, svr, smsgs, soffs);
    fzo_hash_ag(secpar, svl, svr, smsgs, soffs, n, alpha, ch); /* ch == the hash_ch values of :695-698 */
    int code = 0;
    int32_t target[FZO_MAXD], observed[FZO_MAXD];
    for (int k = 0; k < d; k++) { /* :706-714 */
        int64_t acc = 0;
        for (size_t i = 0; i < n; i++) {
            int64_t t = fzo_cent(fzo_cent((int64_t)svl[i * d + k] * ch[i * d + k]) + svr[i * d + k]);
            t = fzo_cent(t * alpha[i * d + k]);
            acc = i ? fzo_cent(acc + t) : t;
        }
        target[k] = (int32_t)acc;
    }
    matvec_one(p, A, agg, observed); /* :715-717 */
    for (int k = 0; k < d; k++)
        if (pymod((int64_t)target[k] - observed[k], FZO_Q) != 0) { code = 3; break; }
    if (!code) {
        int64_t mx = 0;
        for (size_t o = 0; o < skn; o++) { int64_t a = coef[o] < 0 ? -(int64_t)coef[o] : coef[o]; if (a > mx) mx = a; }
        if (mx > p->beta_vf) code = 4; /* :722-725 */
    }
    if (!code) { /* :726-727 weight = count of coefficients != 0 mod q, per polynomial */
        for (int r = 0; r < p->ell && !code; r++) {
            int w = 0;
            for (int k = 0; k < d; k++) w += pymod(coef[(size_t)r * d + k], FZO_Q) != 0;
            if (w > d) code = 5;
        }
    }
    free(coef); free(order); free(svl); free(svr); free(alpha); free(ch); free(smsgs); free(soffs);
    return code;
}
