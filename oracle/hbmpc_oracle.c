/*
 * hbmpc_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file.  The product (mpc-protocols_b200/csrc) never links it and has no
 * CPU fallback.
 *
 * PARITY UNPINNED at the arkworks byte level.  The reference (Stoffel-Labs/mpc-protocols) is Rust;
 * this environment has no rustc/cargo and the arithmetic lives in crates that are not vendored
 * under /root/reference: ark-ff 0.5.0 (Fp Montgomery arithmetic, FftField), ark-poly 0.5.0
 * (DensePolynomial, GeneralEvaluationDomain), ark-bls12-381 0.5.0 (Fr constants).  The oracle
 * restates their published algorithms and is pinned (tests/test_oracle_*.py) against
 *   - every known-answer test the reference's own test-suite holds for this path
 *     (SURVEY.md section 8c items 1-9), and
 *   - an independent Python big-int model (oracle/pymodel.py) on seeded inputs.
 * Residues mod r are unique, so any correct arithmetic agrees with arkworks on canonical values;
 * what is mirrored here is the reference's control flow.
 *
 * All paths below are relative to /root/reference/mpc/src.  Boundary format everywhere: canonical
 * (non-Montgomery) value as 4 x uint64 little-endian limbs == `U256` (ffi/c_bindings/mod.rs:17-49).
 *
 * "Faithful" complexity: share generation is a radix-2 FFT over the size-N domain, matrices are
 * dense mat-vecs, Gao uses the naive O(m^3) Lagrange of common/mod.rs:134-165, OEC retries per
 * round -- the same asymptotics the reference pays, so this file doubles as the CPU baseline.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fr; /* Montgomery form internally */

/* ShareErrorCode: ffi/c_bindings/share/mod.rs:18-37 */
enum { ORC_OK = 0, ORC_INSUFFICIENT = 1, ORC_DEGREE_MISMATCH = 2, ORC_ID_MISMATCH = 3, ORC_INVALID_INPUT = 4,
       ORC_TYPE_MISMATCH = 5, ORC_NO_DOMAIN = 6, ORC_POLY_OP = 7, ORC_DECODING = 8 };

static const uint64_t MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const fr FR_ONE = {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}}; /* R mod r */
static const fr FR_R2 = {{0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}};
static const fr FR_ZERO = {{0, 0, 0, 0}};
/* 2^32-th root of unity 7^((r-1)/2^32), Montgomery form (ark-bls12-381 Fr TWO_ADIC_ROOT_OF_UNITY) */
static const fr FR_ROOT32 = {{0xb9b58d8c5f0e466aULL, 0x5b1b4c801819d7ecULL, 0x0af53ae352a31e64ULL, 0x5bf3adda19e9b27bULL}};
static const uint64_t MONT_INV = 0xfffffffeffffffffULL; /* -r^{-1} mod 2^64 */

#define MAXN 256          /* n <= 255 (honeybadger/mod.rs:441-444) -> domain size <= 256 */
#define MAXP (2 * MAXN + 8)


/* ------------------------------------------------------------------ tiny pthread parallel-for (no OpenMP runtime in this image) */
typedef void (*range_fn)(void *ctx, long lo, long hi);
typedef struct { range_fn fn; void *ctx; long B; long grain; long *next; } par_job;
static void *par_worker(void *arg) {
    par_job *j = (par_job *)arg;
    for (;;) {
        long lo = __atomic_fetch_add(j->next, j->grain, __ATOMIC_RELAXED);
        if (lo >= j->B) break;
        long hi = lo + j->grain < j->B ? lo + j->grain : j->B;
        j->fn(j->ctx, lo, hi);
    }
    return NULL;
}
static void par_for(long B, int threads, long grain, range_fn fn, void *ctx) {
    if (threads <= 1 || B <= grain) { fn(ctx, 0, B); return; }
    if (threads > 256) threads = 256;
    long next = 0;
    par_job job = {fn, ctx, B, grain, &next};
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < threads - 1; ++i) if (pthread_create(&th[started], NULL, par_worker, &job) == 0) started++;
    par_worker(&job);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

/* ------------------------------------------------------------------ field */
static inline int fr_is_zero(const fr *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fr_eq(const fr *a, const fr *b) { return a->l[0] == b->l[0] && a->l[1] == b->l[1] && a->l[2] == b->l[2] && a->l[3] == b->l[3]; }
static inline int geq_mod(const uint64_t *a) {
    for (int i = 3; i >= 0; --i) { if (a[i] > MOD[i]) return 1; if (a[i] < MOD[i]) return 0; }
    return 1;
}
static inline void sub_mod_raw(uint64_t *a) {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) { u128 d = (u128)a[i] - MOD[i] - br; a[i] = (uint64_t)d; br = (d >> 64) & 1; }
}
static inline fr fr_add(fr a, fr b) {
    fr c; u128 cy = 0;
    for (int i = 0; i < 4; ++i) { u128 s = (u128)a.l[i] + b.l[i] + cy; c.l[i] = (uint64_t)s; cy = s >> 64; }
    if (cy || geq_mod(c.l)) sub_mod_raw(c.l);
    return c;
}
static inline fr fr_sub(fr a, fr b) {
    fr c; u128 br = 0;
    for (int i = 0; i < 4; ++i) { u128 d = (u128)a.l[i] - b.l[i] - br; c.l[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) { u128 cy = 0; for (int i = 0; i < 4; ++i) { u128 s = (u128)c.l[i] + MOD[i] + cy; c.l[i] = (uint64_t)s; cy = s >> 64; } }
    return c;
}
static inline fr fr_neg(fr a) { return fr_is_zero(&a) ? a : fr_sub(FR_ZERO, a); }
/* CIOS Montgomery product a*b*R^-1 mod r */
static inline fr fr_mul(fr a, fr b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 cy = 0;
        for (int j = 0; j < 4; ++j) { u128 s = (u128)a.l[j] * b.l[i] + t[j] + cy; t[j] = (uint64_t)s; cy = s >> 64; }
        u128 s = (u128)t[4] + cy; t[4] = (uint64_t)s; t[5] = (uint64_t)(s >> 64);
        uint64_t m = t[0] * MONT_INV;
        cy = ((u128)m * MOD[0] + t[0]) >> 64;
        for (int j = 1; j < 4; ++j) { u128 s2 = (u128)m * MOD[j] + t[j] + cy; t[j - 1] = (uint64_t)s2; cy = s2 >> 64; }
        s = (u128)t[4] + cy; t[3] = (uint64_t)s; t[4] = t[5] + (uint64_t)(s >> 64);
    }
    fr c = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_mod(c.l)) sub_mod_raw(c.l);
    return c;
}
static fr fr_pow_u64(fr b, uint64_t e) {
    fr acc = FR_ONE;
    while (e) { if (e & 1) acc = fr_mul(acc, b); b = fr_mul(b, b); e >>= 1; }
    return acc;
}
static fr fr_inv(fr a) { /* Fermat: a^(r-2) */
    uint64_t e[4] = {MOD[0] - 2, MOD[1], MOD[2], MOD[3]};
    fr acc = FR_ONE;
    for (int i = 255; i >= 0; --i) {
        acc = fr_mul(acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = fr_mul(acc, a);
    }
    return acc;
}
static inline fr fr_from_u64(uint64_t v) { fr a = {{v, 0, 0, 0}}; return fr_mul(a, FR_R2); }
/* canonical limbs -> Montgomery; returns 0 if the value is not < r (from_bigint(..).unwrap() panics there,
 * ffi/c_bindings/mod.rs:37-41) */
static inline int fr_from_canon(const uint64_t *u, fr *out) {
    if (geq_mod(u)) return 0;
    fr a = {{u[0], u[1], u[2], u[3]}};
    *out = fr_mul(a, FR_R2);
    return 1;
}
static inline void fr_to_canon(fr a, uint64_t *u) {
    fr one = {{1, 0, 0, 0}};
    fr c = fr_mul(a, one);
    memcpy(u, c.l, 32);
}

/* ------------------------------------------------------------------ domain (common/mod.rs:51-68) */
static int domain_size(size_t n) { /* GeneralEvaluationDomain::new(n): radix-2, size next_pow2(n) */
    if (n == 0 || n > MAXN) return 0;
    int N = 1; while ((size_t)N < n) N <<= 1;
    return N;
}
static fr domain_gen(int N) { /* group_gen = root32^(2^32 / N) */
    fr w = FR_ROOT32;
    for (uint64_t k = (1ULL << 32) / (uint64_t)N; k > 1; k >>= 1) w = fr_mul(w, w);
    return w;
}
static void domain_elements(size_t n, fr *xs /* [n] */) { /* element(j) = gen^j */
    int N = domain_size(n); fr w = domain_gen(N); fr p = FR_ONE;
    for (size_t j = 0; j < n; ++j) { xs[j] = p; p = fr_mul(p, w); }
}

/* ------------------------------------------------------------------ DensePolynomial (ark-poly 0.5.0 semantics) */
typedef struct { int len; fr c[MAXP]; } poly; /* normalised: len == 0 or c[len-1] != 0 */
static void p_norm(poly *p) { while (p->len > 0 && fr_is_zero(&p->c[p->len - 1])) p->len--; }
static int p_degree(const poly *p) { return p->len == 0 ? 0 : p->len - 1; } /* degree(zero) == 0 */
static void p_set_one(poly *p) { p->len = 1; p->c[0] = FR_ONE; }
static void p_add(poly *o, const poly *a, const poly *b) {
    int n = a->len > b->len ? a->len : b->len;
    for (int i = 0; i < n; ++i) o->c[i] = fr_add(i < a->len ? a->c[i] : FR_ZERO, i < b->len ? b->c[i] : FR_ZERO);
    o->len = n; p_norm(o);
}
static void p_sub(poly *o, const poly *a, const poly *b) {
    int n = a->len > b->len ? a->len : b->len;
    for (int i = 0; i < n; ++i) o->c[i] = fr_sub(i < a->len ? a->c[i] : FR_ZERO, i < b->len ? b->c[i] : FR_ZERO);
    o->len = n; p_norm(o);
}
static void p_mul(poly *o, const poly *a, const poly *b) { /* o must not alias a or b */
    if (a->len == 0 || b->len == 0) { o->len = 0; return; }
    int n = a->len + b->len - 1;
    for (int i = 0; i < n; ++i) o->c[i] = FR_ZERO;
    for (int i = 0; i < a->len; ++i) for (int j = 0; j < b->len; ++j) o->c[i + j] = fr_add(o->c[i + j], fr_mul(a->c[i], b->c[j]));
    o->len = n; p_norm(o);
}
static void p_mul_linear(poly *p, fr x) { /* p *= (X - x) */
    if (p->len == 0) return;
    fr nx = fr_neg(x);
    p->c[p->len] = FR_ZERO;
    for (int i = p->len; i >= 1; --i) p->c[i] = fr_add(p->c[i - 1], fr_mul(p->c[i], nx));
    p->c[0] = fr_mul(p->c[0], nx);
    p->len++; p_norm(p);
}
static void p_scale(poly *o, const poly *a, fr s) {
    for (int i = 0; i < a->len; ++i) o->c[i] = fr_mul(a->c[i], s);
    o->len = a->len; p_norm(o);
}
static fr p_eval(const poly *p, fr x) {
    fr acc = FR_ZERO;
    for (int i = p->len - 1; i >= 0; --i) acc = fr_add(fr_mul(acc, x), p->c[i]);
    return acc;
}
/* divide_with_q_and_r; returns 0 on zero divisor (the reference panics / maps to PolynomialOperationError) */
static int p_divmod(poly *q, poly *r, const poly *a, const poly *b) {
    if (b->len == 0) return 0;
    *r = *a; q->len = 0;
    if (a->len < b->len) return 1;
    q->len = a->len - b->len + 1;
    for (int i = 0; i < q->len; ++i) q->c[i] = FR_ZERO;
    fr li = fr_inv(b->c[b->len - 1]);
    while (r->len >= b->len && r->len > 0) {
        fr cq = fr_mul(r->c[r->len - 1], li);
        int sh = r->len - b->len;
        q->c[sh] = cq;
        for (int i = 0; i < b->len; ++i) r->c[sh + i] = fr_sub(r->c[sh + i], fr_mul(cq, b->c[i]));
        p_norm(r);
    }
    p_norm(q);
    return 1;
}

/* ------------------------------------------------------------------ lagrange_interpolate (common/mod.rs:134-165), naive */
static int lagrange_interpolate(const fr *xs, const fr *ys, int k, poly *result) {
    for (int i = 0; i < k; ++i) for (int j = i + 1; j < k; ++j) if (fr_eq(&xs[i], &xs[j])) return ORC_INVALID_INPUT;
    poly *num = (poly *)malloc(sizeof(poly)), *term = (poly *)malloc(sizeof(poly)), *tmp = (poly *)malloc(sizeof(poly));
    result->len = 0;
    for (int j = 0; j < k; ++j) {
        p_set_one(num);
        fr den = FR_ONE;
        for (int m = 0; m < k; ++m) if (m != j) { p_mul_linear(num, xs[m]); den = fr_mul(den, fr_sub(xs[j], xs[m])); }
        p_scale(term, num, fr_mul(ys[j], fr_inv(den)));
        p_add(tmp, result, term);
        *result = *tmp;
    }
    free(num); free(term); free(tmp);
    return ORC_OK;
}

/* ------------------------------------------------------------------ FFT share generation */
/* compute_shares: robust_interpolate.rs:52-82, shamir.rs:158-196: evals = domain.fft(&poly), take(n) */
static void fft_natural(fr *a, int N, fr w) { /* in-place radix-2 DIT, natural-order output */
    int lg = 0; while ((1 << lg) < N) lg++;
    for (int i = 0; i < N; ++i) {
        int r = 0; for (int b = 0; b < lg; ++b) if (i & (1 << b)) r |= 1 << (lg - 1 - b);
        if (r > i) { fr t = a[i]; a[i] = a[r]; a[r] = t; }
    }
    for (int len = 2; len <= N; len <<= 1) {
        fr wl = w; for (int k = N / len; k > 1; k >>= 1) wl = fr_mul(wl, wl);
        for (int s = 0; s < N; s += len) {
            fr tw = FR_ONE;
            for (int j = 0; j < len / 2; ++j) {
                fr u = a[s + j], v = fr_mul(a[s + j + len / 2], tw);
                a[s + j] = fr_add(u, v); a[s + j + len / 2] = fr_sub(u, v);
                tw = fr_mul(tw, wl);
            }
        }
    }
}

typedef struct { size_t n, d; int N; fr w; const uint64_t *coeffs; uint64_t *shares; int bad; } cs_ctx;
static void cs_range(void *vc, long lo, long hi) {
    cs_ctx *c = (cs_ctx *)vc;
    for (long b = lo; b < hi; ++b) {
        fr a[MAXN];
        for (int i = 0; i < c->N; ++i) a[i] = FR_ZERO;
        for (size_t k = 0; k <= c->d; ++k) if (!fr_from_canon(c->coeffs + ((size_t)b * (c->d + 1) + k) * 4, &a[k])) c->bad = 1;
        fft_natural(a, c->N, c->w);
        for (size_t j = 0; j < c->n; ++j) fr_to_canon(a[j], c->shares + ((size_t)b * c->n + j) * 4);
    }
}
int orc_compute_shares(size_t n, size_t d, size_t B, const uint64_t *coeffs /*[B][d+1]*/, uint64_t *shares /*[B][n]*/, int threads) {
    if (n <= d) return ORC_INVALID_INPUT;                   /* robust_interpolate.rs:59-64 */
    int N = domain_size(n);
    if (!N) return ORC_NO_DOMAIN;                           /* :65-66 */
    cs_ctx c = {n, d, N, domain_gen(N), coeffs, shares, 0};
    par_for((long)B, threads, 256, cs_range, &c);
    return c.bad ? ORC_INVALID_INPUT : ORC_OK;
}

/* make_vandermonde: common/share/mod.rs:31-45 */
int orc_make_vandermonde(size_t n, size_t t, uint64_t *V /*[n][t+1]*/) {
    if (!domain_size(n)) return ORC_NO_DOMAIN;
    fr xs[MAXN]; domain_elements(n, xs);
    for (size_t j = 0; j < n; ++j) {
        fr p = FR_ONE;
        for (size_t k = 0; k <= t; ++k) { fr_to_canon(p, V + (j * (t + 1) + k) * 4); p = fr_mul(p, xs[j]); }
    }
    return ORC_OK;
}

/* apply_vandermonde: common/share/mod.rs:50-76, looped over B chunks as batch_recon.rs:160-165 does.
 * recipient_major != 0 writes out[row][chunk] (the transposition of batch_recon.rs:158-165). */
typedef struct { size_t rows, cols, B; const fr *M; const uint64_t *in; uint64_t *out; int recipient_major; int bad; } am_ctx;
static void am_range(void *vc, long lo, long hi) {
    am_ctx *c = (am_ctx *)vc;
    for (long b = lo; b < hi; ++b) {
        fr v[MAXN];
        for (size_t k = 0; k < c->cols; ++k) if (!fr_from_canon(c->in + ((size_t)b * c->cols + k) * 4, &v[k])) c->bad = 1;
        for (size_t j = 0; j < c->rows; ++j) {
            fr acc = fr_mul(v[0], c->M[j * c->cols]);
            for (size_t k = 1; k < c->cols; ++k) acc = fr_add(acc, fr_mul(v[k], c->M[j * c->cols + k]));
            fr_to_canon(acc, c->out + (c->recipient_major ? (j * c->B + (size_t)b) : ((size_t)b * c->rows + j)) * 4);
        }
    }
}
int orc_apply_matrix(size_t rows, size_t cols, const uint64_t *Mx /*[rows][cols]*/, size_t B, const uint64_t *in /*[B][cols]*/,
                     uint64_t *out, int recipient_major, int threads) {
    if (cols == 0 || cols > MAXN) return ORC_INVALID_INPUT;
    fr *M = (fr *)malloc(sizeof(fr) * rows * cols);
    int bad = 0;
    for (size_t i = 0; i < rows * cols; ++i) if (!fr_from_canon(Mx + i * 4, &M[i])) bad = 1;
    am_ctx c = {rows, cols, B, M, in, out, recipient_major, 0};
    par_for((long)B, threads, 256, am_range, &c);
    free(M);
    return (bad || c.bad) ? ORC_INVALID_INPUT : ORC_OK;
}

int orc_apply_vandermonde(size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *out, int recipient_major, int threads) {
    if (cols == 0 || cols > MAXN) return ORC_INVALID_INPUT;
    if (!domain_size(n)) return ORC_NO_DOMAIN;
    uint64_t *V = (uint64_t *)malloc(32 * n * cols);
    orc_make_vandermonde(n, cols - 1, V);
    int rc = orc_apply_matrix(n, cols, V, B, in, out, recipient_major, threads);
    free(V);
    return rc;
}

/* ------------------------------------------------------------------ robust decoding */
typedef struct { size_t id; fr y; } pt;
static int cmp_pt(const void *a, const void *b) { size_t x = ((const pt *)a)->id, y = ((const pt *)b)->id; return x < y ? -1 : x > y; }

/* robust_interpolate_fnt: robust_interpolate.rs:206-266 */
static int interpolate_fnt(size_t t, size_t n, const pt *sh, size_t cnt, size_t degree, poly *out) {
    fr dom[MAXN]; domain_elements(n, dom);
    size_t m = degree + 1;
    poly *a_poly = (poly *)malloc(sizeof(poly)), *a_der = (poly *)malloc(sizeof(poly)), *basis = (poly *)malloc(sizeof(poly)),
         *rem = (poly *)malloc(sizeof(poly)), *lin = (poly *)malloc(sizeof(poly)), *tmp = (poly *)malloc(sizeof(poly)), *sc = (poly *)malloc(sizeof(poly));
    int rc = ORC_OK;
    p_set_one(a_poly);
    for (size_t i = 0; i < m; ++i) p_mul_linear(a_poly, dom[sh[i].id]);
    a_der->len = a_poly->len > 1 ? a_poly->len - 1 : 0;
    for (int i = 1; i < a_poly->len; ++i) a_der->c[i - 1] = fr_mul(fr_from_u64((uint64_t)i), a_poly->c[i]);
    p_norm(a_der);
    out->len = 0;
    for (size_t i = 0; i < m && rc == ORC_OK; ++i) {
        fr x = dom[sh[i].id];
        fr denom = p_eval(a_der, x);
        if (fr_is_zero(&denom)) { rc = ORC_POLY_OP; break; }
        fr scalar = fr_mul(sh[i].y, fr_inv(denom));
        lin->len = 2; lin->c[0] = fr_neg(x); lin->c[1] = FR_ONE;
        if (!p_divmod(basis, rem, a_poly, lin) || rem->len != 0) { rc = ORC_POLY_OP; break; }
        p_scale(sc, basis, scalar);
        p_add(tmp, out, sc);
        *out = *tmp;
    }
    if (rc == ORC_OK) {
        size_t valid = 0;
        for (size_t i = 0; i < cnt; ++i) { fr e = p_eval(out, dom[sh[i].id]); if (fr_eq(&e, &sh[i].y)) valid++; }
        if (valid < degree + t + 1) rc = ORC_DECODING;
    }
    free(a_poly); free(a_der); free(basis); free(rem); free(lin); free(tmp); free(sc);
    return rc;
}

/* compute_g0_from_domain: robust_interpolate.rs:540-565 (memoised per n in the reference; recomputed here per call
 * unless the caller passes a cached copy) */
static void compute_g0(size_t n, poly *g0) {
    fr dom[MAXN]; domain_elements(n, dom);
    p_set_one(g0);
    for (size_t i = 0; i < n; ++i) p_mul_linear(g0, dom[i]);
}

/* gao_rs_decode: robust_interpolate.rs:456-538 */
static int gao_rs_decode(const fr *received /*[n]*/, size_t k, size_t n, const size_t *erasures, size_t ne, const poly *g0_full, poly *out) {
    if (k > n) return ORC_INVALID_INPUT;
    fr dom[MAXN]; domain_elements(n, dom);
    char erased[MAXN]; memset(erased, 0, sizeof erased);
    size_t s = 0;
    for (size_t i = 0; i < ne; ++i) if (!erased[erasures[i]]) { erased[erasures[i]] = 1; s++; }
    poly *s_poly = (poly *)malloc(sizeof(poly)), *g1 = (poly *)malloc(sizeof(poly)), *g0 = (poly *)malloc(sizeof(poly)), *rem = (poly *)malloc(sizeof(poly));
    poly *r0 = (poly *)malloc(sizeof(poly)), *r1 = (poly *)malloc(sizeof(poly)), *t0 = (poly *)malloc(sizeof(poly)), *t1 = (poly *)malloc(sizeof(poly));
    poly *q = (poly *)malloc(sizeof(poly)), *prod = (poly *)malloc(sizeof(poly)), *nr = (poly *)malloc(sizeof(poly)), *nt = (poly *)malloc(sizeof(poly));
    int rc = ORC_OK;
    p_set_one(s_poly);
    for (size_t i = 0; i < n; ++i) if (erased[i]) p_mul_linear(s_poly, dom[i]);
    fr xs[MAXN], ys[MAXN]; int kn = 0;
    for (size_t i = 0; i < n; ++i) if (!erased[i]) { xs[kn] = dom[i]; ys[kn] = received[i]; kn++; }
    rc = lagrange_interpolate(xs, ys, kn, g1);                                    /* :491 */
    if (rc == ORC_OK) {
        p_divmod(g0, rem, g0_full, s_poly);                                      /* :494-495 */
        size_t threshold = (n - s + k) / 2;                                      /* :498 */
        *r0 = *g0; *r1 = *g1; t0->len = 0; p_set_one(t1);
        while ((size_t)p_degree(r1) >= threshold) {                              /* :510-522 */
            p_divmod(q, rem, r0, r1);
            p_mul(prod, q, r1); p_sub(nr, r0, prod);
            p_mul(prod, q, t1); p_sub(nt, t0, prod);
            *r0 = *r1; *r1 = *nr; *t0 = *t1; *t1 = *nt;
        }
        /* f = g / v, accept iff remainder zero and deg f < k   (:524-537) */
        p_divmod(q, rem, r1, t1);
        p_mul(prod, q, t1); p_sub(nr, r1, prod);
        if (nr->len == 0 && (size_t)p_degree(q) < k) *out = *q; else rc = ORC_DECODING;
    }
    free(s_poly); free(g1); free(g0); free(rem); free(r0); free(r1); free(t0); free(t1); free(q); free(prod); free(nr); free(nt);
    return rc;
}

/* oec_decode: robust_interpolate.rs:579-628 */
static int oec_decode(size_t n, size_t t, const pt *sh, size_t cnt, size_t degree, poly *out, int *round) {
    fr dom[MAXN]; domain_elements(n, dom);
    poly *g0 = (poly *)malloc(sizeof(poly)), *cand = (poly *)malloc(sizeof(poly));
    compute_g0(n, g0);
    int rc = ORC_DECODING;
    for (size_t r = 1; r <= t; ++r) {
        size_t required = degree + t + 1 + r;
        if (cnt < required) break;
        fr received[MAXN]; size_t erasures[MAXN]; size_t ne = 0; char have[MAXN];
        memset(have, 0, sizeof have);
        for (size_t i = 0; i < n; ++i) received[i] = FR_ZERO;
        for (size_t i = 0; i < required; ++i) { received[sh[i].id] = sh[i].y; have[sh[i].id] = 1; }
        for (size_t i = 0; i < n; ++i) if (!have[i]) erasures[ne++] = i;
        if (gao_rs_decode(received, degree + 1, n, erasures, ne, g0, cand) == ORC_OK) {
            size_t matched = 0;
            for (size_t i = 0; i < required; ++i) { fr e = p_eval(cand, dom[sh[i].id]); if (fr_eq(&e, &sh[i].y)) matched++; }
            if (matched >= degree + t + 1) { *out = *cand; *round = (int)r; rc = ORC_OK; break; }
        }
    }
    free(g0); free(cand);
    return rc;
}

/* validation of RobustShare::recover_secret, robust_interpolate.rs:100-142 (degree equality is the caller's:
 * all shares of one call carry `degree`) */
static int validate_ids(size_t n, size_t t, size_t degree, size_t S, const size_t *ids) {
    if (n < 3 * t + 1) return ORC_INVALID_INPUT;
    if (S == 0) return ORC_INVALID_INPUT;
    char seen[MAXN]; memset(seen, 0, sizeof seen);
    for (size_t i = 0; i < S; ++i) { if (ids[i] < MAXN && seen[ids[i]]) return ORC_INVALID_INPUT; if (ids[i] < MAXN) seen[ids[i]] = 1; }
    for (size_t i = 0; i < S; ++i) if (ids[i] >= n) return ORC_INVALID_INPUT;
    if (S < degree + t + 1) return ORC_INVALID_INPUT;
    return ORC_OK;
}

static int recover_core(size_t n, size_t t, size_t degree, size_t S, const size_t *ids, const fr *vals, poly *out, int *path) {
    pt sh[MAXN];
    for (size_t i = 0; i < S; ++i) { sh[i].id = ids[i]; sh[i].y = vals[i]; }
    qsort(sh, S, sizeof(pt), cmp_pt);                                            /* :144-145 */
    if (interpolate_fnt(t, n, sh, degree + t + 1, degree, out) == ORC_OK) { *path = 0; return ORC_OK; }  /* :148 */
    return oec_decode(n, t, sh, S, degree, out, path);                          /* :152 */
}

/* RobustShare::recover_secret (robust_interpolate.rs:94-157), one codeword.
 * coeffs_out[degree+1] zero-padded; *coeff_len = trimmed length the reference returns; flags bit i (arrival order)
 * <=> share i disagrees with the decoded polynomial. */
int orc_robust_recover_secret(size_t n, size_t t, size_t degree, size_t S, const size_t *ids, const uint64_t *vals /*[S]*/,
                              uint64_t *coeffs_out, size_t *coeff_len, uint64_t *secret, int32_t *path, uint64_t *flags /*[ceil(S/64)]*/) {
    if (n > MAXN - 1 || !domain_size(n)) return ORC_NO_DOMAIN;
    int rc = validate_ids(n, t, degree, S, ids);
    if (rc) return rc;
    fr v[MAXN];
    for (size_t i = 0; i < S; ++i) if (!fr_from_canon(vals + 4 * i, &v[i])) return ORC_INVALID_INPUT;
    poly *p = (poly *)malloc(sizeof(poly)); int pth = 0;
    rc = recover_core(n, t, degree, S, ids, v, p, &pth);
    if (rc == ORC_OK) {
        fr dom[MAXN]; domain_elements(n, dom);
        for (size_t k = 0; k <= degree; ++k) fr_to_canon((int)k < p->len ? p->c[k] : FR_ZERO, coeffs_out + 4 * k);
        if (coeff_len) *coeff_len = (size_t)p->len;
        if (secret) fr_to_canon(p_eval(p, FR_ZERO), secret);
        if (path) *path = pth;
        if (flags) {
            for (size_t i = 0; i < (S + 63) / 64; ++i) flags[i] = 0;
            for (size_t i = 0; i < S; ++i) { fr e = p_eval(p, dom[ids[i]]); if (!fr_eq(&e, &v[i])) flags[i >> 6] |= 1ULL << (i & 63); }
        }
    } else if (path) *path = -rc;
    free(p);
    return rc;
}

/* Batched per-codeword robust interpolation: shares[B][S], shared ids (the layout of the C-ABI's
 * hbmpc_robust_interpolate_batch).  Per-item failures land in path[b] = -code; returns first failing code. */
typedef struct { size_t n, degree, t, S, fw; const size_t *ids; const uint64_t *shares; uint64_t *coeffs, *secrets; int32_t *path; uint64_t *flags;
                 int first; long first_b; pthread_mutex_t mu; } rb_ctx;
static void rb_range(void *vc, long lo, long hi) {
    rb_ctx *c = (rb_ctx *)vc;
    for (long b = lo; b < hi; ++b) {
        size_t cl; uint64_t sec[4]; int32_t pth = 0;
        uint64_t *co = c->coeffs + (size_t)b * (c->degree + 1) * 4;
        int rc = orc_robust_recover_secret(c->n, c->t, c->degree, c->S, c->ids, c->shares + (size_t)b * c->S * 4, co, &cl, sec, &pth,
                                           c->flags ? c->flags + (size_t)b * c->fw : NULL);
        if (rc) {
            memset(co, 0, 32 * (c->degree + 1)); memset(sec, 0, 32);
            if (c->flags) memset(c->flags + (size_t)b * c->fw, 0, 8 * c->fw);
            pthread_mutex_lock(&c->mu);
            if (c->first_b < 0 || b < c->first_b) { c->first_b = b; c->first = rc; }
            pthread_mutex_unlock(&c->mu);
        }
        if (c->secrets) memcpy(c->secrets + 4 * (size_t)b, sec, 32);
        if (c->path) c->path[b] = pth;
    }
}
int orc_robust_interpolate_batch(size_t n, size_t degree, size_t t, size_t S, const size_t *ids, size_t B, const uint64_t *shares,
                                 uint64_t *coeffs /*[B][degree+1]*/, uint64_t *secrets /*[B]*/, int32_t *path, uint64_t *flags, int threads) {
    if (n > MAXN - 1 || !domain_size(n)) return ORC_NO_DOMAIN;
    int rc0 = validate_ids(n, t, degree, S, ids);
    if (rc0) return rc0;
    rb_ctx c = {n, degree, t, S, (S + 63) / 64, ids, shares, coeffs, secrets, path, flags, 0, -1, PTHREAD_MUTEX_INITIALIZER};
    par_for((long)B, threads, 4, rb_range, &c);
    return c.first;
}

/* batch_recover_secret: robust_interpolate.rs:284-443.  evals[S][B] sender-major (arrival order), sender_ids[S].
 * coeffs[B][degree+1] zero-padded, coeff_len[B] = the length the reference returns for that chunk (degree+1 on the
 * optimistic path :419-428, trimmed on the fallback path :437-438), path[B] (0 optimistic / r / -code).
 * Return value: the reference aborts the whole call at the first failing chunk (`?` at :437): its code is returned,
 * but all chunks are still decoded so per-item outputs can be compared. */
typedef struct {
    size_t n, degree, t, S, B, needed, m, fw; const size_t *sender_ids; const size_t *order; const uint64_t *evals;
    const fr *dom; const poly *basis; const fr *verify;
    uint64_t *coeffs; size_t *coeff_len; int32_t *path; uint64_t *flags;
    int bad, first; long first_b; pthread_mutex_t mu;
} br_ctx;
static void br_range(void *vc, long lo, long hi) {
    br_ctx *x = (br_ctx *)vc;
    size_t m = x->m, S = x->S, B = x->B;
    poly *p = (poly *)malloc(sizeof(poly));
    for (long c = lo; c < hi; ++c) {
        fr y[MAXN]; /* id-sorted */
        int lbad = 0;
        for (size_t s = 0; s < S; ++s) if (!fr_from_canon(x->evals + (x->order[s] * B + (size_t)c) * 4, &y[s])) lbad = 1;
        if (lbad) { x->bad = 1; continue; }
        int ok = 1;
        for (size_t s = 0; s < x->needed && ok; ++s) {                          /* :404-415 */
            fr acc = FR_ZERO;
            for (size_t i = 0; i < m; ++i) acc = fr_add(acc, fr_mul(x->verify[s * m + i], y[i]));
            if (!fr_eq(&acc, &y[s])) ok = 0;
        }
        uint64_t *co = x->coeffs + (size_t)c * m * 4;
        if (ok) {                                                                /* :417-428 */
            p->len = (int)m;
            for (size_t k = 0; k < m; ++k) {
                fr acc = FR_ZERO;
                for (size_t i = 0; i < m; ++i) if ((int)k < x->basis[i].len) acc = fr_add(acc, fr_mul(x->basis[i].c[k], y[i]));
                p->c[k] = acc;
                fr_to_canon(acc, co + 4 * k);
            }
            if (x->coeff_len) x->coeff_len[c] = m;
            if (x->path) x->path[c] = 0;
            if (x->flags) {   /* decoded polynomial vs every supplied share (bit index = arrival order) */
                p_norm(p);
                uint64_t *fl = x->flags + (size_t)c * x->fw;
                for (size_t i = 0; i < x->fw; ++i) fl[i] = 0;
                for (size_t s = 0; s < S; ++s) {
                    fr e = p_eval(p, x->dom[x->sender_ids[x->order[s]]]);
                    if (!fr_eq(&e, &y[s])) fl[x->order[s] >> 6] |= 1ULL << (x->order[s] & 63);
                }
            }
        } else {                                                                 /* :429-439 */
            uint64_t v2[MAXN * 4];
            for (size_t s = 0; s < S; ++s) memcpy(v2 + 4 * s, x->evals + (s * B + (size_t)c) * 4, 32);
            size_t cl = 0; int32_t pth = 0;
            int rc = orc_robust_recover_secret(x->n, x->t, x->degree, S, x->sender_ids, v2, co, &cl, NULL, &pth, x->flags ? x->flags + (size_t)c * x->fw : NULL);
            if (rc) {
                memset(co, 0, 32 * m);
                if (x->flags) memset(x->flags + (size_t)c * x->fw, 0, 8 * x->fw);
                cl = 0;
                pthread_mutex_lock(&x->mu);
                if (x->first_b < 0 || c < x->first_b) { x->first_b = c; x->first = rc; }
                pthread_mutex_unlock(&x->mu);
            }
            if (x->coeff_len) x->coeff_len[c] = cl;
            if (x->path) x->path[c] = pth;
        }
    }
    free(p);
}
int orc_batch_recover_secret(size_t n, size_t degree, size_t t, size_t S, const size_t *sender_ids, size_t B, const uint64_t *evals,
                             uint64_t *coeffs, size_t *coeff_len, int32_t *path, uint64_t *flags, int threads) {
    if (n < 3 * t + 1) return ORC_INVALID_INPUT;                                 /* :290-296 */
    if (S == 0) return ORC_INVALID_INPUT;                                        /* :297-301 */
    if (B == 0) return ORC_INVALID_INPUT;                                        /* :303-305 */
    if (n > MAXN - 1 || !domain_size(n)) return ORC_NO_DOMAIN;
    if (S > MAXN) return ORC_INVALID_INPUT;
    size_t order[MAXN];                                                          /* sort by sender id :313-315 */
    for (size_t i = 0; i < S; ++i) order[i] = i;
    for (size_t i = 1; i < S; ++i) { size_t k = order[i]; size_t j = i; while (j > 0 && sender_ids[order[j - 1]] > sender_ids[k]) { order[j] = order[j - 1]; j--; } order[j] = k; }
    {   /* duplicate / range checks :317-330 */
        char seen[MAXN]; memset(seen, 0, sizeof seen);
        for (size_t i = 0; i < S; ++i) { size_t id = sender_ids[order[i]]; if (id < MAXN && seen[id]) return ORC_INVALID_INPUT; if (id >= n) return ORC_INVALID_INPUT; seen[id] = 1; }
    }
    size_t needed = degree + t + 1, m = degree + 1;
    if (S < needed) return ORC_INVALID_INPUT;                                    /* :332-341 */
    fr dom[MAXN]; domain_elements(n, dom);
    /* Lagrange basis built once (:348-376) */
    poly *a_poly = (poly *)malloc(sizeof(poly)), *a_der = (poly *)malloc(sizeof(poly)), *rem = (poly *)malloc(sizeof(poly)), *lin = (poly *)malloc(sizeof(poly)), *bp = (poly *)malloc(sizeof(poly));
    poly *basis = (poly *)malloc(sizeof(poly) * m);
    p_set_one(a_poly);
    for (size_t i = 0; i < m; ++i) p_mul_linear(a_poly, dom[sender_ids[order[i]]]);
    a_der->len = a_poly->len > 1 ? a_poly->len - 1 : 0;
    for (int i = 1; i < a_poly->len; ++i) a_der->c[i - 1] = fr_mul(fr_from_u64((uint64_t)i), a_poly->c[i]);
    p_norm(a_der);
    int rc_setup = ORC_OK;
    for (size_t i = 0; i < m; ++i) {
        fr xx = dom[sender_ids[order[i]]];
        fr denom = p_eval(a_der, xx);
        if (fr_is_zero(&denom)) { rc_setup = ORC_POLY_OP; break; }
        lin->len = 2; lin->c[0] = fr_neg(xx); lin->c[1] = FR_ONE;
        if (!p_divmod(bp, rem, a_poly, lin) || rem->len) { rc_setup = ORC_POLY_OP; break; }
        p_scale(&basis[i], bp, fr_inv(denom));
    }
    fr *verify = (fr *)malloc(sizeof(fr) * needed * m);                          /* :392-399 */
    br_ctx x = {n, degree, t, S, B, needed, m, (S + 63) / 64, sender_ids, order, evals, dom, basis, verify, coeffs, coeff_len, path, flags,
                0, 0, -1, PTHREAD_MUTEX_INITIALIZER};
    if (rc_setup == ORC_OK) {
        for (size_t s = 0; s < needed; ++s) for (size_t i = 0; i < m; ++i) verify[s * m + i] = p_eval(&basis[i], dom[sender_ids[order[s]]]);
        par_for((long)B, threads, 64, br_range, &x);
    }
    free(a_poly); free(a_der); free(rem); free(lin); free(bp); free(basis); free(verify);
    if (rc_setup) return rc_setup;
    if (x.bad) return ORC_INVALID_INPUT;
    return x.first;
}

/* NonRobustShare::recover_secret: shamir.rs:199-239 (naive Lagrange through ALL supplied points + degree check) */
int orc_nonrobust_recover_secret(size_t n, size_t deg, size_t S, const size_t *ids, const uint64_t *vals, uint64_t *coeffs_out /*[deg+1]*/,
                                 size_t *coeff_len, uint64_t *secret) {
    if (S == 0) return ORC_INVALID_INPUT;
    for (size_t i = 0; i < S; ++i) for (size_t j = i + 1; j < S; ++j) if (ids[i] == ids[j]) return ORC_INVALID_INPUT;
    if (S < deg + 1) return ORC_INSUFFICIENT;
    if (n > MAXN - 1 || !domain_size(n)) return ORC_NO_DOMAIN;
    for (size_t i = 0; i < S; ++i) if (ids[i] >= n) return ORC_INVALID_INPUT;
    fr dom[MAXN]; domain_elements(n, dom);
    fr xs[MAXN], ys[MAXN];
    for (size_t i = 0; i < S; ++i) { xs[i] = dom[ids[i]]; if (!fr_from_canon(vals + 4 * i, &ys[i])) return ORC_INVALID_INPUT; }
    poly *p = (poly *)malloc(sizeof(poly));
    int rc = lagrange_interpolate(xs, ys, (int)S, p);
    if (rc == ORC_OK && (size_t)p_degree(p) > deg) rc = ORC_DEGREE_MISMATCH;     /* :235-237 */
    if (rc == ORC_OK) {
        for (size_t k = 0; k <= deg; ++k) fr_to_canon((int)k < p->len ? p->c[k] : FR_ZERO, coeffs_out + 4 * k);
        if (coeff_len) *coeff_len = (size_t)p->len;
        /* reference returns result_poly[0] (panics on the zero polynomial); we return 0 there */
        if (secret) fr_to_canon(p->len ? p->c[0] : FR_ZERO, secret);
    }
    free(p);
    return rc;
}

/* Direct entry to gao_rs_decode for the reference's Gao unit tests (robust_interpolate.rs:683-756) */
int orc_gao_rs_decode(size_t n, size_t k, const uint64_t *received /*[n]*/, const size_t *erasures, size_t ne, uint64_t *coeffs_out /*[k]*/, size_t *coeff_len) {
    if (n > MAXN - 1 || !domain_size(n)) return ORC_NO_DOMAIN;
    fr rec[MAXN];
    for (size_t i = 0; i < n; ++i) if (!fr_from_canon(received + 4 * i, &rec[i])) return ORC_INVALID_INPUT;
    poly *g0 = (poly *)malloc(sizeof(poly)), *out = (poly *)malloc(sizeof(poly));
    compute_g0(n, g0);
    int rc = gao_rs_decode(rec, k, n, erasures, ne, g0, out);
    if (rc == ORC_OK) {
        for (size_t i = 0; i < k; ++i) fr_to_canon((int)i < out->len ? out->c[i] : FR_ZERO, coeffs_out + 4 * i);
        if (coeff_len) *coeff_len = (size_t)out->len;
    }
    free(g0); free(out);
    return rc;
}

/* lagrange_interpolate exposed for tests */
int orc_lagrange_interpolate(size_t k, const uint64_t *xs, const uint64_t *ys, uint64_t *coeffs_out /*[k]*/, size_t *coeff_len) {
    if (k > MAXN) return ORC_INVALID_INPUT;
    fr x[MAXN], y[MAXN];
    memset(x, 0, sizeof x);
    memset(y, 0, sizeof y);
    for (size_t i = 0; i < k; ++i) if (!fr_from_canon(xs + 4 * i, &x[i]) || !fr_from_canon(ys + 4 * i, &y[i])) return ORC_INVALID_INPUT;
    poly *p = (poly *)malloc(sizeof(poly));
    int rc = lagrange_interpolate(x, y, (int)k, p);
    if (rc == ORC_OK) {
        for (size_t i = 0; i < k; ++i) fr_to_canon((int)i < p->len ? p->c[i] : FR_ZERO, coeffs_out + 4 * i);
        if (coeff_len) *coeff_len = (size_t)p->len;
    }
    free(p);
    return rc;
}

/* element-wise share algebra (common/mod.rs:167-300): op 0 add, 1 sub, 2 mul (share_mul / Mul<F>) */
typedef struct { int op; const uint64_t *a, *b; uint64_t *out; int bad; } ew_ctx;
static void ew_range(void *vc, long lo, long hi) {
    ew_ctx *c = (ew_ctx *)vc;
    for (long i = lo; i < hi; ++i) {
        fr x, y;
        if (!fr_from_canon(c->a + 4 * i, &x) || !fr_from_canon(c->b + 4 * i, &y)) { c->bad = 1; continue; }
        fr z = c->op == 0 ? fr_add(x, y) : c->op == 1 ? fr_sub(x, y) : fr_mul(x, y);
        fr_to_canon(z, c->out + 4 * i);
    }
}
int orc_elementwise(int op, size_t B, const uint64_t *a, const uint64_t *b, uint64_t *out, int threads) {
    ew_ctx c = {op, a, b, out, 0};
    par_for((long)B, threads, 4096, ew_range, &c);
    return c.bad ? ORC_INVALID_INPUT : ORC_OK;
}

/* The multi-operator steps of the share algebra, operator by operator in the reference's order (what the product's
 * hbmpc_share_algebra_fused computes in one pass over the vectors):
 *   step 0  triple_generation.rs:332-340   out0 = share_mul(a, b) - r_2t                       in = {a, b, r_2t}
 *   step 1  multiplication.rs:417-426      out0 = a - x, out1 = b - y                          in = {a, x, b, y}
 *   step 2  multiplication.rs:79-97        mult_subs = da*db; share = c - mult_subs;
 *                                          share2 = share - y*da; out0 = share2 - x*db         in = {c, x, y, da, db}  */
int orc_share_algebra_step(int step, size_t B, const uint64_t *const *in, uint64_t *const *out) {
    const int nin = step == 0 ? 3 : step == 1 ? 4 : 5;
    if (step < 0 || step > 2) return ORC_INVALID_INPUT;
    for (size_t i = 0; i < B; ++i) {
        fr v[5];
        for (int k = 0; k < nin; ++k)
            if (!fr_from_canon(in[k] + 4 * i, &v[k])) return ORC_INVALID_INPUT;
        if (step == 0) {
            fr_to_canon(fr_sub(fr_mul(v[0], v[1]), v[2]), out[0] + 4 * i);
        } else if (step == 1) {
            fr_to_canon(fr_sub(v[0], v[1]), out[0] + 4 * i);
            fr_to_canon(fr_sub(v[2], v[3]), out[1] + 4 * i);
        } else {
            fr mult_subs = fr_mul(v[3], v[4]);
            fr mult_sub_a_y = fr_mul(v[2], v[3]);
            fr mult_sub_b_x = fr_mul(v[1], v[4]);
            fr share = fr_sub(v[0], mult_subs);
            fr share2 = fr_sub(share, mult_sub_a_y);
            fr_to_canon(fr_sub(share2, mult_sub_b_x), out[0] + 4 * i);
        }
    }
    return ORC_OK;
}

void orc_domain_element(size_t n, size_t j, uint64_t *out) {
    int N = domain_size(n); fr w = domain_gen(N);
    fr_to_canon(fr_pow_u64(w, j), out);
}

int orc_max_threads(void) {
    long nc = sysconf(_SC_NPROCESSORS_ONLN);
    return nc > 0 ? (int)nc : 1;
}
