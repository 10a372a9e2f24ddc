// robust.cuh -- K4: Reed-Solomon robust interpolation for the items that fail the optimistic check.
//
// Reference behaviour being reproduced (robust_interpolate.rs:94-157, 456-538, 579-628): after the optimistic
// attempt fails, OEC round r = 1..t looks at the lowest d+t+1+r ids ("prefix P_r"), runs Gao's decoder on it and
// accepts iff the decoded polynomial disagrees with at most r shares of that prefix.  Gao decodes up to
// floor((t+r)/2) >= r errors, and two distinct degree-<=d polynomials cannot both be within distance r <= t of the same
// prefix (they would differ on >= t+r+1 of its points), so the reference's outcome is exactly
//     "the first r for which SOME polynomial of degree <= d has <= r mismatches in P_r; that polynomial",
// independent of Gao's internals.  Any bounded-distance decoder with radius >= r therefore returns bit-identical
// (coefficients, path, flags).  This file uses syndromes + inversion-free Berlekamp-Massey + Chien + Forney:
//
//   attempt(P, nsyn, maxL):   S_j = sum_{i<P} u_i x_i^j y_i  (u_i = 1/prod_{l != i}(x_i - x_l), j < nsyn = P-(d+1));
//                             BM -> locator Lambda (degree L); accept iff L <= maxL and Lambda has L roots among
//                             the prefix points; Forney -> error values.
//   fast path  : ONE attempt on all S supplied shares (maxL = min(t, floor((S-d-1)/2))).  If it succeeds the true
//                polynomial and all e <= t error positions are known, and the reference's round is the smallest r with
//                #errors in P_r <= r (no other polynomial can be accepted earlier: it would need >= t+1 mismatches).
//   exact path : if the fast attempt fails (more than maxL errors overall), attempts r = 1..rmax on the prefixes
//                P_r with maxL = r -- the literal OEC loop.
//
// All evaluation points are N-th roots of unity, so the three "evaluate at every point" steps are size-N transforms:
//   syndromes  S_j = sum_i (u_i y_i) w^(id_i j)        = forward NTT of the weighted word (first nsyn outputs),
//   Chien      Lambda(w^(-id))                          = inverse-direction NTT of Lambda's coefficients,
//   Forney     Omega(w^(-id)), Lambda'(w^(-id))         = two more (when the locator is long enough to pay off),
// ~N/2*log2(N) products each instead of nsyn*P, P*L and 2*L^2.
// One thread decodes one codeword; per-thread polynomials live in a strided global workspace (coalesced across the
// warp).  The corrected coefficients are obtained by linearity from Lc * y[0..m), which the dense kernel wrote for the
// failing items:   coeffs(f) = Lc * y[0..m) - sum_{i in E, i<m} e_i * Lc[:, i].
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fr.cuh"

// The decoder's kernels keep one codeword per thread and their lanes diverge: the branch-free conditional subtraction (fr.cuh:
// cond_sub_mod_t<true>) measures 4-5 % faster here, while the lockstep transforms prefer the top-limb test.  This header is the last
// user of these names in the translation unit; they are restored at its end.
#define mont_mul mont_mul_t<true>
#define fr_add fr_add_t<true>
#define acc_reduce acc_reduce_t<true>

namespace hb {

#ifdef HB_ROBUST_PROF
__device__ unsigned long long g_robust_prof[8];  // cycles: syndromes, BM, omega, chien, forney, inversion+values, finish, items
#define HB_PROF_MARK(slot) do { long long t_ = clock64(); if ((threadIdx.x & 31) == 0) atomicAdd(&g_robust_prof[slot], (unsigned long long)(t_ - prof_t)); prof_t = t_; } while (0)
#else
#define HB_PROF_MARK(slot) do { } while (0)
#endif

struct RobustArgs {
    const uint4 *in;            // supplied shares: element (b, arrival j) at in[(b*in_sb + j*in_sc)*2 ..]
    long long in_sb, in_sc, B;
    const unsigned int *list;   // items that failed the optimistic check
    const unsigned int *count;
    const unsigned char *fail_scan;  // non-null: no list -- every thread scans fail_scan[b] of its own items (small batches)
    int need_lc;                // the optimistic stage did not leave Lc*y[lowest m] in `coeffs`: compute it here
    unsigned int *hist;         // optional [S]: hist[arrival j] += 1 for every error found at sender j (scout pass)
    unsigned int list_first;    // list mode: process list[list_first .. min(count, list_first + list_max))
    unsigned int list_max;      // 0: no limit
    unsigned char *clear_fail;  // optional: fail flag of every processed item is cleared (scout pass before a re-compaction)
    int S, m, t, needed, rmax;  // m = d+1, needed = d+t+1, rmax = min(t, S-needed)
    int fast;                   // attempt 0 (all S shares) is enabled
    const int *att_P;           // [1 + rmax] prefix size per attempt (index 0 = fast path)
    const int *att_nsyn;        // [1 + rmax]
    const int *att_maxL;        // [1 + rmax]
    const long long *att_uoff;  // [1 + rmax] offset of the attempt's u2 / uinv vectors [P]
    const uint4 *u2;            // u_i * R^2: canonical y_i times this is the Montgomery form of u_i*y_i
    const uint4 *uinv;          // prod_{l != i, l < P} (x_i - x_l), canonical (Montgomery c times this is canonical e)
    const int *sid;             // [S] domain index (share id) of sorted position i
    const uint4 *tw;            // [N/2] w^k   (Montgomery)
    const uint4 *itw;           // [N/2] w^-k  (Montgomery)
    int logn;                   // N = 1 << logn
    const uint4 *xs;            // [S] x_i (Montgomery), sorted by id
    const uint4 *xinv;          // [S] x_i^{-1} (Montgomery)
    const uint4 *Lc;            // [m][m] Lagrange coefficient matrix of the lowest m ids (Montgomery)
    const uint4 *Veval;         // [S][m] x_s^k (Montgomery)
    const int *order;           // [S] arrival index of sorted position i
    uint4 *coeffs;              // [B][mout] (already holds Lc*y from the optimistic kernel)
    int mout;                   // m, or 1 when only the secret is wanted
    int *path;
    unsigned long long *flags;  // may be nullptr
    int flag_words;
    unsigned int *fail_any;     // set to 1 when any item fails to decode
    uint4 *ws;                  // workspace: ws_elems Fr per thread, strided by total thread count
    int ws_elems;
    unsigned int *dense_fail_flag;  // scan mode: set to 1 when some warp finds at least half of its items failing
    unsigned int attack_min;        // list mode: dense_fail_flag is set when at least this many items are listed (0: never)
    unsigned int *count_out;        // list mode: the number of listed items is stored here (pinned host memory: the host reads it after
                                    // the call's synchronisation without a copy of its own)
    unsigned int skip_above;        // list mode: do nothing when at least this many items are listed (0: no limit) -- a launch enqueued
                                    // before the host knows the count; larger sets are decoded by the staged pipeline instead
    int hist_only;              // scout pass ahead of any other stage: only the per-sender error histogram is updated
    int skip_coeffs;            // staged decoder, all N points supplied: the coefficients are corrected by a transform afterwards
};

struct FrWs {
    uint4 *base;
    size_t stride;
    __device__ __forceinline__ void ld(uint32_t (&a)[8], int e) const {
        const uint4 *p = base + (size_t)e * stride * 2;
        load_fr(a, p[0], p[1]);
    }
    __device__ __forceinline__ void st(int e, const uint32_t (&a)[8]) const {
        uint4 *p = base + (size_t)e * stride * 2;
        p[0] = make_uint4(a[0], a[1], a[2], a[3]);
        p[1] = make_uint4(a[4], a[5], a[6], a[7]);
    }
};

__device__ __forceinline__ void ldg_fr(uint32_t (&a)[8], const uint4 *p) { load_fr(a, __ldg(p), __ldg(p + 1)); }
__device__ __forceinline__ void set_zero(uint32_t (&a)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0;
}
__device__ __forceinline__ void copy8(uint32_t (&d)[8], const uint32_t (&s)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = s[i];
}
__device__ __forceinline__ void fr_neg(uint32_t (&d)[8], const uint32_t (&a)[8]) {
    uint32_t z[8];
    set_zero(z);
    fr_sub(d, z, a);
}

// a^(r-2) in Montgomery form (Fermat).  ~255 squarings + ~128 products; used once per accepted attempt.
__device__ __noinline__ void fr_inv_mont(uint32_t (&out)[8], const uint32_t (&a)[8]) {
    const uint32_t e[8] = {0xffffffffu, 0xfffffffeu, HB_R2, HB_R3, HB_R4, HB_R5, HB_R6, HB_R7};
    uint32_t acc[8];
    one_mont_limbs(acc);
#pragma unroll 1
    for (int i = 254; i >= 0; --i) {
        uint32_t s[8];
        mont_mul(s, acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1u) mont_mul(acc, s, a);
        else copy8(acc, s);
    }
    copy8(out, acc);
}

// workspace slots (element offsets) -- sized by the host with the same formulae
struct WsLayout {
    int syn, lam, bp, om, num, den, pre, ev, total;
    __host__ __device__ WsLayout(int nsyn_max, int t, int N) {
        syn = 0;  // transform buffer of N elements; the syndromes are its first nsyn entries
        lam = syn + (nsyn_max > N ? nsyn_max : N);
        bp = lam + (t + 2);
        om = bp + (t + 2);
        num = om + (t + 1);
        den = num + (t + 1);
        pre = den + (t + 1);
        ev = pre + (t + 1);
        total = ev + (t + 1);
    }
};

// In-place radix-2 decimation-in-time transform of the N workspace elements at `slot0` (input in bit-reversed order,
// output in natural order), twiddles tw[k] = g^k in Montgomery form: out[j] = sum_p in_natural[p] * g^(p j).
__device__ __noinline__ void ntt_serial(const FrWs &ws, int slot0, int logn, const uint4 *tw) {
    const int N = 1 << logn;
#pragma unroll 1
    for (int s = 0; s < logn; ++s) {
        const int half = 1 << s, tws = logn - 1 - s;
#pragma unroll 1
        for (int i = 0; i < (N >> 1); ++i) {
            const int j = i & (half - 1);
            const int pos = ((i >> s) << (s + 1)) | j;
            uint32_t u[8], v[8], t[8], sm[8], df[8];
            ws.ld(u, slot0 + pos);
            ws.ld(v, slot0 + pos + half);
            if (j) {
                uint32_t w[8];
                ldg_fr(w, tw + ((size_t)(j << tws)) * 2);
                mont_mul(t, v, w);
            } else {
                copy8(t, v);
            }
            fr_add(sm, u, t);
            fr_sub(df, u, t);
            ws.st(slot0 + pos, sm);
            ws.st(slot0 + pos + half, df);
        }
    }
}
__device__ __forceinline__ int bitrev_n(int x, int logn) { return (int)(__brev((unsigned)x) >> (32 - logn)); }

// One bounded-distance decoding attempt.  Returns the number of errors L (positions in rootpos[0..L), ascending sorted
// position; canonical error values in ws[ev + q]) or -1.
__device__ __noinline__ int rs_attempt(const RobustArgs &a, int att, long long b, const FrWs &ws, const WsLayout &lay, int *rootpos) {
    const int P = a.att_P[att], nsyn = a.att_nsyn[att], maxL = a.att_maxL[att];
    const int N = 1 << a.logn;
    const uint4 *U2 = a.u2 + a.att_uoff[att] * 2;
    const uint4 *ybase = a.in + b * a.in_sb * 2;
    uint32_t zero[8];
    set_zero(zero);
#ifdef HB_ROBUST_PROF
    long long prof_t = clock64();
#endif

    // ---- syndromes (Montgomery form): forward transform of the weighted word, S_j = sum_i (u_i y_i) w^(id_i j)
#pragma unroll 1
    for (int p = 0; p < N; ++p) ws.st(lay.syn + p, zero);
#pragma unroll 1
    for (int i = 0; i < P; ++i) {
        uint32_t y[8], u[8], w[8];
        ldg_fr(y, ybase + (long long)a.order[i] * a.in_sc * 2);
        ldg_fr(u, U2 + i * 2);
        mont_mul(w, y, u);
        ws.st(lay.syn + bitrev_n(a.sid[i], a.logn), w);
    }
    ntt_serial(ws, lay.syn, a.logn, a.tw);
    HB_PROF_MARK(0);

    // ---- inversion-free Berlekamp-Massey:  Lambda <- bdis*Lambda - delta * z^shift * Bp
    uint32_t one[8], bdis[8];
    one_mont_limbs(one);
    copy8(bdis, one);
    ws.st(lay.lam, one);
    ws.st(lay.bp, one);
    int L = 0, lenB = 1, shift = 1;
#pragma unroll 1
    for (int j = 0; j < nsyn; ++j) {
        uint32_t delta[8];
        {
            acc_t A;
            acc_zero(A);
            const int lim = L < j ? L : j;
#pragma unroll 1
            for (int l = 0; l <= lim; ++l) {
                uint32_t x[8], s[8];
                ws.ld(x, lay.lam + l);
                ws.ld(s, lay.syn + j - l);
                acc_mac(A, x, s);
            }
            acc_reduce(A, delta);
        }
        if (fr_is_zero(delta)) { ++shift; continue; }
        uint32_t nd[8];
        fr_neg(nd, delta);
        const bool grow = 2 * L <= j;
        const int newL = grow ? j + 1 - L : L;
        if (newL > maxL) return -1;
#pragma unroll 1
        for (int l = newL; l >= 0; --l) {
            uint32_t lam[8], bl[8], res[8];
            if (l <= L) ws.ld(lam, lay.lam + l); else set_zero(lam);
            const int bi = l - shift;
            const bool hasb = bi >= 0 && bi < lenB;
            if (hasb) ws.ld(bl, lay.bp + bi); else set_zero(bl);
            acc_t A;
            acc_zero(A);
            acc_mac(A, lam, bdis);
            if (hasb) acc_mac(A, bl, nd);
            acc_reduce(A, res);
            ws.st(lay.lam + l, res);
            if (grow && l <= L) ws.st(lay.bp + l, lam);
        }
        if (grow) {
            lenB = L + 1;
            L = newL;
            copy8(bdis, delta);
            shift = 1;
        } else {
            ++shift;
        }
    }
    HB_PROF_MARK(1);
    if (L == 0) return 0;

    // ---- Forney numerator polynomial first (it needs the syndromes, which share the transform buffer):
    //      Omega = S*Lambda mod z^L
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
        acc_t A;
        acc_zero(A);
#pragma unroll 1
        for (int k = 0; k <= l; ++k) {
            uint32_t x[8], sy[8];
            ws.ld(x, lay.lam + k);
            ws.ld(sy, lay.syn + l - k);
            acc_mac(A, x, sy);
        }
        uint32_t o[8];
        acc_reduce(A, o);
        ws.st(lay.om + l, o);
    }

    HB_PROF_MARK(2);
    // ---- Chien search over the prefix points: Lambda(x_i^{-1}) == 0  <=>  position i is in error
    int nroots = 0;
    // warp-uniform choice (lanes hold different L): one transform for everybody beats paying for both code paths
    const int Lmax = __reduce_max_sync(__activemask(), L);
    const int nbfly = (N >> 1) * a.logn;          // products of one size-N transform
    const bool chien_ntt = Lmax * P > nbfly;      // direct: L Horner steps at each of the P points
    if (chien_ntt) {  // all N values Lambda(w^-j) with one inverse-direction transform
#pragma unroll 1
        for (int p = 0; p < N; ++p) ws.st(lay.syn + p, zero);
#pragma unroll 1
        for (int l = 0; l <= L; ++l) {
            uint32_t c[8];
            ws.ld(c, lay.lam + l);
            ws.st(lay.syn + bitrev_n(l, a.logn), c);
        }
        ntt_serial(ws, lay.syn, a.logn, a.itw);
    }
#pragma unroll 1
    for (int i = 0; i < P; ++i) {
        uint32_t v[8];
        if (chien_ntt) {
            ws.ld(v, lay.syn + a.sid[i]);
        } else {
            uint32_t z[8];
            ldg_fr(z, a.xinv + i * 2);
            ws.ld(v, lay.lam + L);
#pragma unroll 1
            for (int l = L - 1; l >= 0; --l) {
                uint32_t c[8], p[8];
                mont_mul(p, v, z);
                ws.ld(c, lay.lam + l);
                fr_add(v, p, c);
            }
        }
        if (fr_is_zero(v)) {
            if (nroots < L) rootpos[nroots] = i;
            ++nroots;
        }
    }
    HB_PROF_MARK(3);
    if (nroots != L) return -1;

    // ---- Forney:  c_i = -x_i Omega(x_i^{-1}) / Lambda'(x_i^{-1});  e_i = c_i * uinv_i
    uint32_t run[8];
    copy8(run, one);
    const bool forney_ntt = 2 * Lmax * Lmax + Lmax > 2 * nbfly;  // direct: 2 L^2 + L products for Omega and Lambda' at the L roots
    if (forney_ntt) {
        // numerators: Omega at every w^-j
#pragma unroll 1
        for (int p = 0; p < N; ++p) ws.st(lay.syn + p, zero);
#pragma unroll 1
        for (int l = 0; l < L; ++l) {
            uint32_t c[8];
            ws.ld(c, lay.om + l);
            ws.st(lay.syn + bitrev_n(l, a.logn), c);
        }
        ntt_serial(ws, lay.syn, a.logn, a.itw);
#pragma unroll 1
        for (int q = 0; q < L; ++q) {
            uint32_t v[8], x[8], numv[8];
            ws.ld(v, lay.syn + a.sid[rootpos[q]]);
            ldg_fr(x, a.xs + rootpos[q] * 2);
            mont_mul(numv, v, x);
            ws.st(lay.num + q, numv);
        }
        // denominators: Lambda'(z) = sum_{l>=1} l * Lambda_l z^(l-1) at every w^-j
#pragma unroll 1
        for (int p = 0; p < N; ++p) ws.st(lay.syn + p, zero);
        uint32_t lm[8];
        copy8(lm, one);
#pragma unroll 1
        for (int l = 1; l <= L; ++l) {
            uint32_t c[8], lc[8], nl[8];
            ws.ld(c, lay.lam + l);
            mont_mul(lc, c, lm);
            ws.st(lay.syn + bitrev_n(l - 1, a.logn), lc);
            fr_add(nl, lm, one);
            copy8(lm, nl);
        }
        ntt_serial(ws, lay.syn, a.logn, a.itw);
#pragma unroll 1
        for (int q = 0; q < L; ++q) {
            uint32_t dv[8], nr[8];
            ws.ld(dv, lay.syn + a.sid[rootpos[q]]);
            if (fr_is_zero(dv)) return -1;
            ws.st(lay.den + q, dv);
            ws.st(lay.pre + q, run);
            mont_mul(nr, run, dv);
            copy8(run, nr);
        }
    } else {
    // derivative coefficients dL_{l-1} = l * Lambda_l once (the B polynomial of Berlekamp-Massey is dead: reuse its slots)
    {
        uint32_t lm[8];
        copy8(lm, one);
#pragma unroll 1
        for (int l = 1; l <= L; ++l) {
            uint32_t c[8], lc[8], nl[8];
            ws.ld(c, lay.lam + l);
            mont_mul(lc, c, lm);
            ws.st(lay.bp + l - 1, lc);
            fr_add(nl, lm, one);
            copy8(lm, nl);
        }
    }
#pragma unroll 1
    for (int q = 0; q < L; ++q) {
        const int i = rootpos[q];
        uint32_t z[8], x[8], v[8], dv[8];
        ldg_fr(z, a.xinv + i * 2);
        ldg_fr(x, a.xs + i * 2);
        ws.ld(v, lay.om + L - 1);
        ws.ld(dv, lay.bp + L - 1);
#pragma unroll 1
        for (int l = L - 2; l >= 0; --l) {
            uint32_t c[8], p[8], c2[8], p2[8];
            mont_mul(p, v, z);
            ws.ld(c, lay.om + l);
            fr_add(v, p, c);
            mont_mul(p2, dv, z);
            ws.ld(c2, lay.bp + l);
            fr_add(dv, p2, c2);
        }
        uint32_t numv[8];
        mont_mul(numv, v, x);
        if (fr_is_zero(dv)) return -1;
        ws.st(lay.num + q, numv);
        ws.st(lay.den + q, dv);
        ws.st(lay.pre + q, run);
        uint32_t nr[8];
        mont_mul(nr, run, dv);
        copy8(run, nr);
    }
    }
    HB_PROF_MARK(4);
    uint32_t inv[8];
    fr_inv_mont(inv, run);
    const uint4 *U = a.uinv + a.att_uoff[att] * 2;
#pragma unroll 1
    for (int q = L - 1; q >= 0; --q) {
        uint32_t pre[8], den[8], dinv[8], numv[8], c[8], nc[8], u[8], e[8], ninv[8];
        ws.ld(pre, lay.pre + q);
        ws.ld(den, lay.den + q);
        ws.ld(numv, lay.num + q);
        mont_mul(dinv, inv, pre);
        mont_mul(ninv, inv, den);
        copy8(inv, ninv);
        mont_mul(c, numv, dinv);
        fr_neg(nc, c);
        ldg_fr(u, U + rootpos[q] * 2);
        mont_mul(e, nc, u);  // Montgomery c times canonical uinv -> canonical error value
        ws.st(lay.ev + q, e);
    }
    HB_PROF_MARK(5);
    return L;
}

// Writes the outcome of one decoded item: path < 0 -> DecodingError (zeroed outputs); otherwise the corrected coefficients by
// linearity, the flag bits (error positions + shares beyond the examined prefix that disagree with the decoded polynomial),
// the scout histogram and the path.  rootpos[0..L) = sorted positions of the errors, load_ev(e, q) = canonical error value q.
template <typename EvLoad>
__device__ __forceinline__ void robust_store_item(const RobustArgs &a, long long b, int L, const int *rootpos, int path, int Pused, EvLoad load_ev,
                                                  bool do_hist = true) {
    if (a.hist_only) {
        if (path >= 0)
            for (int q = 0; q < L; ++q) atomicAdd(&a.hist[a.order[rootpos[q]]], 1u);
        return;
    }
    uint4 *co = a.coeffs + b * a.mout * 2;
    unsigned long long *fl = a.flags ? a.flags + b * a.flag_words : nullptr;
    if (path < 0) {
        for (int k = 0; k < a.mout; ++k) { co[k * 2] = make_uint4(0, 0, 0, 0); co[k * 2 + 1] = make_uint4(0, 0, 0, 0); }
        if (fl) for (int w = 0; w < a.flag_words; ++w) fl[w] = 0ull;
        a.path[b] = -8;
        if (a.clear_fail) a.clear_fail[b] = 0;
        *(volatile unsigned int *)a.fail_any = 1u;
        return;
    }
    // corrected coefficients by linearity (Lc*y first when the optimistic stage did not provide it)
    for (int k = 0; k < (a.skip_coeffs ? 0 : a.mout); ++k) {
        if (a.need_lc) {
            acc_t A0;
            acc_zero(A0);
            const uint4 *yb = a.in + b * a.in_sb * 2;
            for (int i = 0; i < a.m; ++i) {
                uint32_t y[8], lc[8];
                ldg_fr(y, yb + (long long)a.order[i] * a.in_sc * 2);
                ldg_fr(lc, a.Lc + ((size_t)k * a.m + i) * 2);
                acc_mac(A0, y, lc);
            }
            uint32_t c0[8];
            acc_reduce(A0, c0);
            co[k * 2] = make_uint4(c0[0], c0[1], c0[2], c0[3]);
            co[k * 2 + 1] = make_uint4(c0[4], c0[5], c0[6], c0[7]);
        }
        acc_t A;
        acc_zero(A);
        bool any = false;
        for (int q = 0; q < L; ++q) {
            if (rootpos[q] >= a.m) break;
            uint32_t e[8], lc[8];
            load_ev(e, q);
            ldg_fr(lc, a.Lc + ((size_t)k * a.m + rootpos[q]) * 2);
            acc_mac(A, e, lc);
            any = true;
        }
        if (any) {
            uint32_t corr[8], c[8], r[8];
            acc_reduce(A, corr);
            load_fr(c, co[k * 2], co[k * 2 + 1]);
            fr_sub(r, c, corr);
            co[k * 2] = make_uint4(r[0], r[1], r[2], r[3]);
            co[k * 2 + 1] = make_uint4(r[4], r[5], r[6], r[7]);
        }
    }
    if (fl) {
        for (int w = 0; w < a.flag_words; ++w) fl[w] = 0ull;
        for (int q = 0; q < L; ++q) {
            const int j = a.order[rootpos[q]];
            fl[j >> 6] |= 1ull << (j & 63);
        }
        // shares beyond the examined prefix: evaluate the decoded polynomial (needs all m coefficients: mout == m)
        
        const uint4 *ybase = a.in + b * a.in_sb * 2;
        for (int s = Pused; s < a.S; ++s) {
            acc_t A;
            acc_zero(A);
            for (int k = 0; k < a.m; ++k) {
                uint32_t c[8], v[8];
                load_fr(c, co[k * 2], co[k * 2 + 1]);
                ldg_fr(v, a.Veval + ((size_t)s * a.m + k) * 2);
                acc_mac(A, c, v);
            }
            uint32_t fv[8], y[8];
            acc_reduce(A, fv);
            const int j = a.order[s];
            ldg_fr(y, ybase + (long long)j * a.in_sc * 2);
            if (!fr_eq(fv, y)) fl[j >> 6] |= 1ull << (j & 63);
        }
    }
    if (a.hist && do_hist)
        for (int q = 0; q < L; ++q) atomicAdd(&a.hist[a.order[rootpos[q]]], 1u);
    if (a.clear_fail) a.clear_fail[b] = 0;
    a.path[b] = path;
}

#ifndef HB_ROBUST_MAXT
#define HB_ROBUST_MAXT 85  // t < n/3, n <= 256
#endif

#ifndef HB_ROBUST_MINB
#define HB_ROBUST_MINB 8
#endif
__global__ void __launch_bounds__(128, HB_ROBUST_MINB) robust_kernel(const RobustArgs a) {
    fma_ballast(a.rmax < 0, a.fail_any);
    size_t cnt = a.fail_scan ? (size_t)a.B : (size_t)*a.count;
    if (!a.fail_scan && a.count_out && blockIdx.x == 0 && threadIdx.x == 0) *(volatile unsigned int *)a.count_out = (unsigned int)cnt;
    if (!a.fail_scan && a.skip_above && cnt >= (size_t)a.skip_above) return;
    if (!a.fail_scan && a.dense_fail_flag && a.attack_min && cnt >= (size_t)a.attack_min && blockIdx.x == 0 && threadIdx.x == 0)
        *(volatile unsigned int *)a.dense_fail_flag = 1u;   // list mode: a large failing set (asynchronous calls learn of the attack this way)
    if (!a.fail_scan && a.list_max && cnt > (size_t)a.list_first + a.list_max) cnt = (size_t)a.list_first + a.list_max;
    const size_t T = (size_t)gridDim.x * blockDim.x;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int nsyn_max = 0;
    for (int i = a.fast ? 0 : 1; i <= a.rmax; ++i) nsyn_max = max(nsyn_max, a.att_nsyn[i]);
    const WsLayout lay(nsyn_max, a.t, 1 << a.logn);
    FrWs ws{a.ws + g * 2, T};
    int rootpos[HB_ROBUST_MAXT];

    for (size_t idx = g + (a.fail_scan ? 0 : a.list_first); idx < cnt; idx += T) {
        if (a.fail_scan) {
            const bool mine = a.fail_scan[idx] != 0;
            // half of a warp's items failing in a batch that was not compacted: tell the host (mapped word, plain store), so that
            // the next call of this context counts its failing items and can take the staged decoder (attacks persist)
            if (a.dense_fail_flag && __popc(__ballot_sync(__activemask(), mine)) >= 16) *(volatile unsigned int *)a.dense_fail_flag = 1u;
            if (!mine) continue;
        }
        const long long b = a.fail_scan ? (long long)idx : (long long)a.list[idx];
        int L = -1, path = -8, used_att = -1;
        if (a.fast) {
            L = rs_attempt(a, 0, b, ws, lay, rootpos);
            if (L >= 0) {
                // all error positions known (L <= t): the reference accepts in the first round whose prefix holds <= r of them
                used_att = 0;
                int q = 0;
                for (int r = 1; r <= a.rmax; ++r) {
                    while (q < L && rootpos[q] < a.needed + r) ++q;
                    if (q <= r) { path = r; break; }
                }
            }
        }
        if (L < 0) {
            for (int r = 1; r <= a.rmax; ++r) {
                L = rs_attempt(a, r, b, ws, lay, rootpos);
                if (L >= 0) { path = r; used_att = r; break; }
            }
        }
        robust_store_item(a, b, L, rootpos, path, used_att < 0 ? 0 : a.att_P[used_att], [&](uint32_t (&e)[8], int q) { ws.ld(e, lay.ev + q); });
    }
}

// Speculative decode for persistent attackers (the same <= t senders wrong in every chunk).  The dense kernel has
// interpolated P from d+1 senders believed honest and set flag bit j for every other supplied sender whose share disagrees
// with P.  If at most t shares disagree, P is the polynomial the reference decodes (it is the only one within distance t of
// the prefixes it examines) and its round is the first r whose prefix holds at most r of the mismatches -- the same rule as
// the decoder's fast path.  Otherwise the item stays in the failing set for the full decoder.
struct SpecArgs {
    const unsigned int *list;      // failing items
    unsigned int first, count;     // process list[first .. count)
    const unsigned long long *sflags;  // [B][flag_words] mismatch bits in arrival order (scratch)
    unsigned long long *flags;     // user flags (may be nullptr)
    int flag_words;
    const int *pos_of;             // [S] sorted position of arrival index j
    int S, t, needed, rmax;
    int *path;
    unsigned char *fail;           // cleared for accepted items
    unsigned int *fail_any;
    uint4 *coeffs;
    int mout;
};
__global__ void spec_finalize_kernel(const SpecArgs a) {
    for (unsigned int idx = a.first + blockIdx.x * blockDim.x + threadIdx.x; idx < a.count; idx += gridDim.x * blockDim.x) {
        const long long b = a.list[idx];
        const unsigned long long *sf = a.sflags + b * a.flag_words;
        int e = 0;
        for (int w = 0; w < a.flag_words; ++w) e += __popcll(sf[w]);
        if (e > a.t) continue;  // not explained by <= t errors: left to the full decoder (fail[b] stays set)
        // prefix counts: cnt[p] = number of mismatches at sorted positions < p, evaluated lazily for p = needed + r
        int path = -8;
        for (int r = 1; r <= a.rmax; ++r) {
            const int P = a.needed + r;
            int c = 0;
            for (int j = 0; j < a.S; ++j)
                if (((sf[j >> 6] >> (j & 63)) & 1ull) && a.pos_of[j] < P) ++c;
            if (c <= r) { path = r; break; }
        }
        a.fail[b] = 0;
        a.path[b] = path;
        if (path < 0) {  // the reference's OEC loop runs out of rounds: DecodingError
            uint4 *co = a.coeffs + b * a.mout * 2;
            for (int k = 0; k < 2 * a.mout; ++k) co[k] = make_uint4(0, 0, 0, 0);
            *(volatile unsigned int *)a.fail_any = 1u;
        }
        if (a.flags)
            for (int w = 0; w < a.flag_words; ++w) a.flags[b * a.flag_words + w] = path < 0 ? 0ull : sf[w];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Staged decoder for large failing sets (the throughput path of K4).  robust_kernel above keeps one codeword per thread
// from start to finish: its lanes diverge with the error count of their codewords (Berlekamp-Massey trip counts) and its
// transforms run serially through a global workspace.  Here the fast attempt (all S shares, see the header) is cut into
// uniform stages over a WAVE of failing items ("slots"):
//   1. syndromes        ntt_kernel<LOGN,2> on the weighted word (warp-cooperative transform, item_list indirection)
//   2. Berlekamp-Massey bm_segment_kernel, one thread per slot, in segments of a few iterations.  Between segments the
//                       slots are re-sorted by their current locator degree L (counting sort) and their state is moved to
//                       the new position (permute_kernel), so the lanes of a warp run the same trip counts: codewords
//                       with few errors reveal themselves early (L stops growing) and no longer ride along with the
//                       longest one.  BM state lives in a group-interleaved layout (32 consecutive positions interleaved
//                       at 16-byte granularity), so every per-lane access of a warp is one coalesced 512-byte request.
//   3. Omega, Lambda'   omega_kernel; leaves Lambda, Lambda', Omega item-major per slot for the transforms
//   4. Chien search     ntt_kernel<LOGN,3> on Lambda -> bit mask of the roots among the supplied ids
//   5. Forney values    ntt_kernel<LOGN,4> on Omega and Lambda' -> values at the roots only
//   6. finish           staged_finish_kernel (sorted by L): error values, path rule, corrected coefficients, flags
// Items whose fast attempt fails (more than maxL errors overall, a locator without enough roots) are appended to a second
// list and go through robust_kernel's exact path (fast = 0): outcomes are bit-identical by construction, whatever the route.
struct StagedArgs {
    // Berlekamp-Massey state by POSITION (group-interleaved), ping-pong: [0] current, [1] destination of the next permute
    uint4 *synG[2], *lamG[2], *bpG[2];   // [W/32][ld][2][32]
    uint4 *bdisP[2];                     // [W/32][1][2][32]
    int4 *stateP[2];                     // [W] (L, lenB, shift, dead)
    unsigned int *originP[2];            // [W] slot of the item at this position
    unsigned char *keyP;                 // [W] sort key of the position: maxL - L (descending L), 255 = dead
    // per SLOT (item-major): inputs / outputs of the Chien / Forney transforms
    uint4 *lam, *bp, *om;       // [W][tp]     Lambda, Lambda', Omega (zero padded)
    uint4 *num, *den;           // [W][tp]     Omega / Lambda' at the roots (later: canonical error values in num)
    int4 *state;                // [W]         final (L, -, -, dead)
    unsigned int *rootmask;     // [W][8]
    unsigned char *key;         // [W]         final sort key
    const unsigned int *perm;   // [W]         permute: old position of new position q; finish: slots in sorted order
    unsigned int W;
    int syn_ld, tp, nsyn, maxL, j0, j1;
    // finish
    unsigned int list_first;    // slot s decodes item list[list_first + s]
    const int *pos_of_dom;      // [N] sorted position of the share with domain index k (-1: not supplied)
    const uint4 *uinv0;         // [S] attempt-0 uinv
    unsigned int *list2, *count2;  // items left to the exact path
    unsigned int hist_slots;    // the per-sender error histogram (RobustArgs::hist) samples the slots below this number
    uint4 *runs;                // [W] by sorted index: product of the slot's Forney denominators, then its inverse
    unsigned char *okf;         // [W] by sorted index: the fast attempt has produced a consistent locator
    int direct;                 // the items come straight from the all-points NTT check (no dense check has looked at the examined
                                // prefix): no error inside the prefix means the reference's optimistic attempt succeeds (path 0)
    // device-count mode (asynchronous calls: no host decision): W is an upper bound, the slots s >= *cnt_dev - list_first are not
    // items at all (dead from the start, never handed to the exact path); attack_flag is raised when the count is large
    const unsigned int *cnt_dev;
    unsigned int *attack_flag;
    unsigned int attack_min;
};
__device__ __forceinline__ unsigned int staged_live(const StagedArgs &a) {
    if (!a.cnt_dev) return a.W;
    const unsigned int c = *a.cnt_dev;
    const unsigned int live = c > a.list_first ? c - a.list_first : 0u;
    return live < a.W ? live : a.W;
}

__device__ __forceinline__ void ld_fr2(uint32_t (&a)[8], const uint4 *p) { load_fr(a, p[0], p[1]); }
__device__ __forceinline__ void st_fr2(uint4 *p, const uint32_t (&a)[8]) {
    p[0] = make_uint4(a[0], a[1], a[2], a[3]);
    p[1] = make_uint4(a[4], a[5], a[6], a[7]);
}
// element e of the thread's position in a group-interleaved array with `ld` elements per position
struct GView {
    uint4 *base;  // array + (group*ld*2)*32 + lane
    __device__ __forceinline__ GView(uint4 *arr, unsigned int pos, int ld) : base(arr + ((size_t)(pos >> 5) * ld * 2) * 32 + (pos & 31)) {}
    __device__ __forceinline__ void ld(uint32_t (&a)[8], int e) const { load_fr(a, base[(size_t)e * 64], base[(size_t)e * 64 + 32]); }
    __device__ __forceinline__ void st(int e, const uint32_t (&a)[8]) const {
        base[(size_t)e * 64] = make_uint4(a[0], a[1], a[2], a[3]);
        base[(size_t)e * 64 + 32] = make_uint4(a[4], a[5], a[6], a[7]);
    }
};

// Berlekamp-Massey iterations j0 <= j < j1 of every live position (same recurrences as rs_attempt)
#ifndef HB_BM_MINB
#define HB_BM_MINB 5
#endif
__global__ void __launch_bounds__(128, HB_BM_MINB) bm_segment_kernel(const StagedArgs a) {
    fma_ballast(a.syn_ld < 0, a.rootmask);
    const unsigned int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= a.W) return;
    if (a.cnt_dev && a.j0 == 0) {
        const unsigned int live = staged_live(a);
        if (pos == 0 && a.attack_flag && live >= a.attack_min) *(volatile unsigned int *)a.attack_flag = 1u;
        if (pos >= live) {   // not an item: dead from the start
            a.stateP[0][pos] = make_int4(0, 0, 0, 1);
            a.originP[0][pos] = pos;
            a.keyP[pos] = (unsigned char)255;
            return;
        }
    }
    const GView lam(a.lamG[0], pos, a.tp), bp(a.bpG[0], pos, a.tp), syn(a.synG[0], pos, a.syn_ld), bd(a.bdisP[0], pos, 1);
    uint32_t one[8], bdis[8];
    one_mont_limbs(one);
    int L, lenB, shift;
    if (a.j0 == 0) {
        copy8(bdis, one);
        lam.st(0, one);
        bp.st(0, one);
        L = 0; lenB = 1; shift = 1;
        a.originP[0][pos] = pos;
    } else {
        const int4 st = a.stateP[0][pos];
        if (st.w) {  // dead: the fast attempt cannot succeed
            a.keyP[pos] = (unsigned char)255;   // (the key array is not permuted: keep dead positions sorted to the end)
            return;
        }
        L = st.x; lenB = st.y; shift = st.z;
        bd.ld(bdis, 0);
    }
    bool dead = false;
#pragma unroll 1
    for (int j = a.j0; j < a.j1; ++j) {
        uint32_t delta[8];
        {
            acc_t A;
            acc_zero(A);
            const int lim = L < j ? L : j;
            // software pipelining with two operand sets (no register moves): the next operands are in flight during a product
            uint32_t x0[8], s0[8], x1[8], s1[8];
            lam.ld(x0, 0);
            syn.ld(s0, j);
#pragma unroll 1
            for (int l = 0; l <= lim; l += 2) {
                const int l1 = l + 1 <= lim ? l + 1 : lim, l2 = l + 2 <= lim ? l + 2 : lim;
                lam.ld(x1, l1);
                syn.ld(s1, j - l1);
                acc_mac(A, x0, s0);
                lam.ld(x0, l2);
                syn.ld(s0, j - l2);
                if (l + 1 <= lim) acc_mac(A, x1, s1);
            }
            acc_reduce(A, delta);
        }
        if (fr_is_zero(delta)) { ++shift; continue; }
        uint32_t nd[8];
        fr_neg(nd, delta);
        const bool grow = 2 * L <= j;
        const int newL = grow ? j + 1 - L : L;
        if (newL > a.maxL) { dead = true; break; }
        uint32_t lmn[8], bln[8];  // operands of the next coefficient, loaded one step ahead
        if (newL <= L) lam.ld(lmn, newL); else set_zero(lmn);
        if (newL - shift >= 0 && newL - shift < lenB) bp.ld(bln, newL - shift); else set_zero(bln);
#pragma unroll 1
        for (int l = newL; l >= 0; --l) {
            uint32_t lm[8], bl[8], res[8];
            copy8(lm, lmn);
            copy8(bl, bln);
            const int bi = l - shift;
            const bool hasb = bi >= 0 && bi < lenB;
            if (l >= 1) {
                if (l - 1 <= L) lam.ld(lmn, l - 1); else set_zero(lmn);
                if (bi - 1 >= 0 && bi - 1 < lenB) bp.ld(bln, bi - 1); else set_zero(bln);
            }
            // two interleaved-row products and a modular addition: fewer multiply-pipe instructions than two lazy products
            // followed by a reduction of the 512-bit sum
            mont_mul(res, lm, bdis);
            if (hasb) {
                uint32_t p2[8], sm[8];
                mont_mul(p2, bl, nd);
                fr_add(sm, res, p2);
                copy8(res, sm);
            }
            lam.st(l, res);
            if (grow && l <= L) bp.st(l, lm);
        }
        if (grow) {
            lenB = L + 1;
            L = newL;
            copy8(bdis, delta);
            shift = 1;
        } else {
            ++shift;
        }
    }
    a.stateP[0][pos] = make_int4(L, lenB, shift, dead ? 1 : 0);
    bd.st(0, bdis);
    a.keyP[pos] = dead ? (unsigned char)255 : (unsigned char)(a.maxL - L);  // longest locators first: their CTAs must not start last
}

// moves the live state of old position perm[q] to new position q (buffers [0] -> [1]); the host swaps the buffers afterwards
__global__ void __launch_bounds__(256) permute_kernel(const StagedArgs a) {
    const unsigned int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.W) return;
    const unsigned int src = a.perm[q];
    const int4 st = a.stateP[0][src];
    a.stateP[1][q] = st;
    a.originP[1][q] = a.originP[0][src];
    if (st.w) return;
    uint32_t v[8];
    {
        const GView s0(a.bdisP[0], src, 1), d0(a.bdisP[1], q, 1);
        s0.ld(v, 0);
        d0.st(0, v);
    }
    {
        const GView s0(a.lamG[0], src, a.tp), d0(a.lamG[1], q, a.tp);
        for (int e = 0; e <= st.x; ++e) { s0.ld(v, e); d0.st(e, v); }
    }
    {
        const GView s0(a.bpG[0], src, a.tp), d0(a.bpG[1], q, a.tp);
        for (int e = 0; e < st.y; ++e) { s0.ld(v, e); d0.st(e, v); }
    }
    {
        const GView s0(a.synG[0], src, a.syn_ld), d0(a.synG[1], q, a.syn_ld);
        for (int e = 0; e < a.nsyn; ++e) { s0.ld(v, e); d0.st(e, v); }
    }
}

// counting sort of the slots by key (any order inside a bin): hist -> exclusive scan -> scatter
__global__ void sort_hist_kernel(const unsigned char *key, unsigned int W, unsigned int *hist) {
    __shared__ unsigned int h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < W; i += gridDim.x * blockDim.x) atomicAdd(&h[key[i]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[i], h[i]);
}
__global__ void sort_scan_kernel(unsigned int *hist) {  // <<<1, 32>>>: hist -> exclusive prefix sums
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int i = 0; i < 256; ++i) { const unsigned int c = hist[i]; hist[i] = run; run += c; }
    }
}
__global__ void sort_scatter_kernel(const unsigned char *key, unsigned int W, unsigned int *offs, unsigned int *perm) {
    __shared__ unsigned int h[256], base[256];
    const unsigned int per = (W + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = min(W, lo + per);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (unsigned int i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&h[key[i]], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        base[i] = h[i] ? atomicAdd(&offs[i], h[i]) : 0u;
        h[i] = 0;
    }
    __syncthreads();
    for (unsigned int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const unsigned int k = key[i];
        perm[base[k] + atomicAdd(&h[k], 1u)] = i;
    }
}

// Omega = S*Lambda mod z^L and the Lambda' coefficients l*Lambda_l, read by position, written per slot (item-major, zero
// padded) for the Chien / Forney transforms together with the slot's final state and sort key
__global__ void __launch_bounds__(128, 4) omega_kernel(const StagedArgs a) {
    fma_ballast(a.syn_ld < 0, a.rootmask);
    const unsigned int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= a.W) return;
    const unsigned int slot = a.originP[0][pos];
    const int4 st = a.stateP[0][pos];
    a.state[slot] = st;
    a.key[slot] = st.w ? (unsigned char)255 : (unsigned char)(a.maxL - st.x);
    if (st.w) return;
    const int L = st.x;
    const GView lamg(a.lamG[0], pos, a.tp), syn(a.synG[0], pos, a.syn_ld);
    uint4 *lam = a.lam + (size_t)slot * a.tp * 2, *bp = a.bp + (size_t)slot * a.tp * 2, *om = a.om + (size_t)slot * a.tp * 2;
    uint32_t zero[8], one[8], lmul[8];
    set_zero(zero);
    one_mont_limbs(one);
    copy8(lmul, one);
#pragma unroll 1
    for (int l = 0; l < L; ++l) {
        acc_t A;
        acc_zero(A);
#pragma unroll 1
        for (int k = 0; k <= l; ++k) {
            uint32_t x[8], sy[8];
            lamg.ld(x, k);
            syn.ld(sy, l - k);
            acc_mac(A, x, sy);
        }
        uint32_t o[8];
        acc_reduce(A, o);
        st_fr2(om + (size_t)l * 2, o);
    }
    {
        uint32_t c[8];
        lamg.ld(c, 0);
        st_fr2(lam, c);
    }
#pragma unroll 1
    for (int l = 1; l <= L; ++l) {
        uint32_t c[8], lc[8], nl[8];
        lamg.ld(c, l);
        st_fr2(lam + (size_t)l * 2, c);
        mont_mul(lc, c, lmul);
        st_fr2(bp + (size_t)(l - 1) * 2, lc);
        fr_add(nl, lmul, one);
        copy8(lmul, nl);
    }
    for (int l = L; l < a.tp; ++l) { st_fr2(om + (size_t)l * 2, zero); st_fr2(bp + (size_t)l * 2, zero); }
    for (int l = L + 1; l < a.tp; ++l) st_fr2(lam + (size_t)l * 2, zero);
}

// error positions of a slot from its root mask; returns the number of roots (rootpos filled up to L entries)
__device__ __forceinline__ int staged_roots(const StagedArgs &s, unsigned int slot, int N, int L, int *rootpos) {
    const unsigned int *mk = s.rootmask + (size_t)slot * 8;
    int nroots = 0;
    for (int w = 0; w < (N + 31) / 32; ++w) {
        unsigned int bits = mk[w];
        while (bits) {
            const int k = w * 32 + __ffs(bits) - 1;
            bits &= bits - 1;
            if (nroots < L) rootpos[nroots] = s.pos_of_dom[k];
            ++nroots;
        }
    }
    return nroots;
}

// 6a. per slot (sorted order idx): prefix products of the Forney denominators Lambda'(x_q^-1) into `om`, their product into
// runs[idx] (1 for slots whose fast attempt has failed, so that the batched inversion is not poisoned), verdict into okf[idx]
__global__ void __launch_bounds__(128, 4) staged_prefix_kernel(const RobustArgs a, const StagedArgs s) {
    fma_ballast(s.syn_ld < 0, s.rootmask);
    const unsigned int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= s.W) return;
    const unsigned int slot = s.perm ? s.perm[idx] : idx;
    const int4 st = s.state[slot];
    const int L = st.x;
    bool ok = !st.w && L <= s.maxL && !(s.direct && L == 0);
    if (ok) {
        const unsigned int *mk = s.rootmask + (size_t)slot * 8;
        int nroots = 0;
        for (int w = 0; w < 8; ++w) nroots += __popc(mk[w]);
        ok = nroots == L;
    }
    uint32_t one[8], run[8];
    one_mont_limbs(one);
    copy8(run, one);
    if (ok) {
        const uint4 *den = s.den + (size_t)slot * s.tp * 2;
        uint4 *pre = s.om + (size_t)slot * s.tp * 2;
#pragma unroll 1
        for (int q = 0; q < L; ++q) {
            uint32_t dv[8], nr[8];
            ld_fr2(dv, den + (size_t)q * 2);
            if (fr_is_zero(dv)) { ok = false; break; }
            st_fr2(pre + (size_t)q * 2, run);
            mont_mul(nr, run, dv);
            copy8(run, nr);
        }
    }
    if (!ok) copy8(run, one);
    st_fr2(s.runs + (size_t)idx * 2, run);
    s.okf[idx] = ok ? 1 : 0;
}

// 6b. runs[idx] <- runs[idx]^-1, HB_INV_BATCH consecutive values per thread with one Fermat inversion (Montgomery's trick)
#define HB_INV_BATCH 8
__global__ void __launch_bounds__(128) staged_invert_kernel(const StagedArgs s) {
    fma_ballast(s.syn_ld < 0, s.rootmask);
    const unsigned int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * HB_INV_BATCH;
    if (i0 >= s.W) return;
    const int cnt = (int)min((unsigned int)HB_INV_BATCH, s.W - i0);
    uint32_t pre[HB_INV_BATCH][8], run[8], inv[8];
    one_mont_limbs(run);
#pragma unroll
    for (int k = 0; k < HB_INV_BATCH; ++k) {
        copy8(pre[k], run);
        if (k < cnt) {
            uint32_t v[8], nr[8];
            ld_fr2(v, s.runs + (size_t)(i0 + k) * 2);
            mont_mul(nr, run, v);
            copy8(run, nr);
        }
    }
    fr_inv_mont(inv, run);
#pragma unroll
    for (int k = HB_INV_BATCH - 1; k >= 0; --k) {
        if (k < cnt) {
            uint32_t v[8], iv[8], ninv[8];
            ld_fr2(v, s.runs + (size_t)(i0 + k) * 2);
            mont_mul(iv, inv, pre[k]);
            mont_mul(ninv, inv, v);
            copy8(inv, ninv);
            st_fr2(s.runs + (size_t)(i0 + k) * 2, iv);
        }
    }
}

// 6c. error values, path rule, outputs
__global__ void __launch_bounds__(128, 4) staged_finish_kernel(const RobustArgs a, const StagedArgs s) {
    fma_ballast(s.syn_ld < 0, s.rootmask);
    const unsigned int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= s.W) return;
    const unsigned int slot = s.perm ? s.perm[idx] : idx;
    if (s.cnt_dev && slot >= staged_live(s)) return;   // device-count mode: not an item
    const unsigned int item = a.list[s.list_first + slot];
    const long long b = (long long)item;
    const int L = s.state[slot].x;
    if (!s.okf[idx]) {  // the fast attempt failed: this item takes the exact path
        s.list2[atomicAdd(s.count2, 1u)] = item;
        if (s.direct) for (int w = 0; w < 8; ++w) s.rootmask[(size_t)slot * 8 + w] = 0u;
        return;
    }
    int rootpos[HB_ROBUST_MAXT];
    staged_roots(s, slot, 1 << a.logn, L, rootpos);
    uint4 *num = s.num + (size_t)slot * s.tp * 2;
    if (L > 0) {
        // c_q = -x_q Omega(x_q^-1) / Lambda'(x_q^-1), e_q = c_q * uinv_q
        const uint4 *den = s.den + (size_t)slot * s.tp * 2, *pre = s.om + (size_t)slot * s.tp * 2;
        uint32_t inv[8];
        ld_fr2(inv, s.runs + (size_t)idx * 2);
#pragma unroll 1
        for (int q = L - 1; q >= 0; --q) {
            uint32_t pr[8], dv[8], dinv[8], nv[8], xq[8], numv[8], c[8], nc[8], u[8], e[8], ninv[8];
            ld_fr2(pr, pre + (size_t)q * 2);
            ld_fr2(dv, den + (size_t)q * 2);
            ld_fr2(nv, num + (size_t)q * 2);
            ldg_fr(xq, a.xs + rootpos[q] * 2);
            mont_mul(numv, nv, xq);
            mont_mul(dinv, inv, pr);
            mont_mul(ninv, inv, dv);
            copy8(inv, ninv);
            mont_mul(c, numv, dinv);
            fr_neg(nc, c);
            ldg_fr(u, s.uinv0 + rootpos[q] * 2);
            mont_mul(e, nc, u);  // Montgomery c times canonical uinv -> canonical error value
            st_fr2(num + (size_t)q * 2, e);
        }
    }
    int path = -8;
    if (s.direct && rootpos[0] >= a.needed) path = 0;
    else {
        int q = 0;
        for (int r = 1; r <= a.rmax; ++r) {
            while (q < L && rootpos[q] < a.needed + r) ++q;
            if (q <= r) { path = r; break; }
        }
    }
    robust_store_item(a, b, L, rootpos, path, a.S, [&](uint32_t (&e)[8], int q) { ld_fr2(e, num + (size_t)q * 2); }, slot < s.hist_slots);
    if (s.direct && path < 0) for (int w = 0; w < 8; ++w) s.rootmask[(size_t)slot * 8 + w] = 0u;  // zeroed outputs stay zero
}

// items with fail[b] != 0 -> list (any order)
__global__ void compact_kernel(const unsigned char *fail, long long B, unsigned int *list, unsigned int *count) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x)
        if (fail[b]) list[atomicAdd(count, 1u)] = (unsigned int)b;
}

}  // namespace hb

#undef mont_mul
#undef fr_add
#undef acc_reduce
