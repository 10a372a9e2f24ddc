// ntt.cuh -- K1/K2 on the evaluation domain: out[b][j] = sum_{k<cols} in[b][k] * w_N^(jk), j < n <= N = next_pow2(n).
//
// The reference evaluates the share polynomial with `domain.fft(&poly)` and keeps the first n values
// (robust_interpolate.rs:68-79, shamir.rs:181-193), and its Vandermonde matrix has entries element(j)^k = w_N^(jk)
// (common/share/mod.rs:31-45): both are the first n outputs of a zero-padded size-N number-theoretic transform.
// Residues are unique, so a radix-2 NTT returns bit-identical shares with ~N/2*log2(N) modular products per item instead
// of n*cols (n=64, d=21: ~130 products against 1408 multiply-accumulates + 64 reductions).
//
// ntt_kernel<LOGN, MODE> -- decimation in time.  N/4 threads (N/8 for N = 256; all in one warp) own one item; each thread holds
// 4 (8) elements in registers and runs 2 (3) butterfly stages per pass; passes exchange through a padded shared-memory buffer
// (conflict-free 128-bit accesses) with __syncwarp only.  Pass 0 takes the input in bit-reversed order from per-thread staging slots
// that cp.async fills one tile ahead (zero padding beyond `cols` costs no loads), the last pass writes natural-order outputs j < n.
// Data stays canonical, twiddles are in Montgomery form (fr.cuh), so there is no conversion pass.
//   MODE 0  forward transform (K1 share generation, K2 Vandermonde apply)
//   MODE 1  inverse transform of all N supplied shares + "top coefficients vanish" check (K3 fast path, a10 degree check)
//   MODE 2  inverse/forward transform of a weighted word with missing ids (K3 erasure check; syndromes of the staged decoder)
//   MODE 3/4/5  transforms of the staged robust decoder (robust.cuh): Chien search into a root mask, Forney values gathered at the
//           roots, sparse inverse transform subtracted from the coefficients
// ntt64_cta_kernel<MODE> (MODE 0/1, N = 64, large batches) regroups a tile's work across the warps of the CTA so that trivial
// twiddles and known-zero operands are skipped by whole warps (see its header below).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fr.cuh"
#include "matvec.cuh"  // ldg_stream / stg_stream

// ntt_kernel / ntt64_cta_kernel (small batches, and the transforms of the staged decoder): the branch-free conditional subtraction
// measures faster here (fr.cuh: cond_sub_mod_t<true>; K4 n = 128 leg 8.8 -> 9.2 M codewords/s together with robust.cuh), unlike in
// ntt16x_kernel, whose routines call the plain names.  Restored at the end of this header.
#define mont_mul mont_mul_t<true>
#define fr_add fr_add_t<true>

namespace hb {

struct NttArgs {
    const uint4 *in;   // element (b, k) at in[(b*in_sb + k*in_sc)*2 ..]
    uint4 *out;        // element (b, j) at out[(b*out_sb + j*out_sr)*2 ..]
    const uint4 *tw;   // [N/2][2] w_N^k in Montgomery form
    long long B, in_sb, in_sc, out_sb, out_sr;
    int cols, n;
    unsigned int *err;
    // MODE 1 (inverse transform + degree check, the all-shares-present fast path of K3):
    const int *in_map;     // domain index j -> record index of the share with id j (nullptr: identity)
    const uint4 *scale;    // N^{-1} in Montgomery form (unused by the kernels since the scaling is a shift: fr_div_pow2)
    int m, mout;           // coefficients k < mout are stored (scaled), coefficients k >= m must vanish
    unsigned char *fail;   // fail[b] = 1 when some coefficient k >= m is non-zero ...
    unsigned int *fail_list, *fail_count;   // ... and b is appended to fail_list[(*fail_count)++] (optional)
    // MODE 2 (inverse transform of an erasure-weighted word, the general optimistic check of K3): the share with id k is
    // multiplied by wt[k] = Zc(w^k) (Zc = product over the ids outside the examined set; Montgomery form) while it is
    // loaded, ids outside the examined set contribute zero (in_map[k] < 0); outputs are left unscaled.
    const uint4 *wt;       // [N][2]
    int *path;             // MODE 1/2: path[b] = 0 is written here (the decoder overwrites it for items that fail)
    // staged robust decoder (robust.cuh): the input item of output slot b is item_list[b] (MODE 2: syndromes of the failing
    // items); MODE 3 (Chien search) sets bit `pos` of rootmask[b][8] for every transformed value that is zero at a domain
    // position of the supplied id set (idset[pos] >= 0); MODE 4 stores the value at the q-th set bit of rootmask[b] to
    // out[b][q] (Forney numerators / denominators at the error positions only)
    const unsigned int *item_list;
    unsigned int *rootmask;
    const int *idset;
    // MODE 5 (correction of the staged decoder, all N points supplied): inverse transform of the sparse error word (value q of
    // in[b] at the q-th set bit of rootmask[b]), scaled by N^{-1}, SUBTRACTED from out[item_list[b]][k], k < mout
    int out_group32;       // MODE 2: output slot b, element pos at out[(((b>>5)*out_sb + pos)*2 + half)*32 + (b&31)] (groups of 32
                           // slots interleaved at 16-byte granularity: one thread per slot reads it coalesced)
    // MODE 2, two-sided recovery (hi_cnt > 0): besides the coefficients k < mout, the hi_cnt coefficients hi_top, hi_top-1, ... are
    // stored, at out positions mout, mout+1, ... (the top of Q = P*Zc in reversed order: P's upper half is recovered from it)
    int hi_top, hi_cnt;
    // ntt_kernel: optional device-side bound of the batch (staged decoder without a host count): items b >= *b_dev - b_first are skipped
    const unsigned int *b_dev;
    unsigned int b_first;
    // ntt16x_kernel: dynamic tile queue (nullptr: static round-robin).  work[0] = tiles handed out beyond the first one of every
    // warp, work[1] = warps that have finished; the last warp to finish zeroes both, so the words are ready for the next launch
    // on the same stream.
    unsigned long long *work;
};

// SKIP_ONE: test for the trivial twiddle (only where the index is warp-uniform -- pass 0 -- so the test folds away or
// never diverges); elsewhere always multiply (tw[0] is the Montgomery form of 1).
template <bool SKIP_ONE>
__device__ __forceinline__ void ntt_butterfly(uint32_t (&u)[8], uint32_t (&v)[8], const uint4 *tw, int twidx) {
    uint32_t t[8], s[8], d[8];
    if (!SKIP_ONE || twidx != 0) {
        uint32_t w[8];
        load_fr(w, tw[twidx * 2], tw[twidx * 2 + 1]);
        mont_mul(t, v, w);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = v[i];
    }
    fr_add(s, u, t);
    fr_sub(d, u, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) { u[i] = s[i]; v[i] = d[i]; }
}

// 16-byte asynchronous global->shared copy (LDGSTS, L2 only): the next tile's inputs land in shared memory while the
// current tile is being transformed
__device__ __forceinline__ void cp_async16(uint4 *smem_dst, const uint4 *gmem_src) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// GG butterfly stages on the 2^GG register-resident elements x[e] that sit at positions base + e*h0
template <int GG, int E, bool SKIP_ONE>
__device__ __forceinline__ void ntt_stages(uint32_t (&x)[E][8], const uint4 *tw, int low, int h0, int tw_shift0) {
#pragma unroll
    for (int q = 0; q < GG; ++q) {
        // half = h0 << q; twiddle exponent of position p: (p mod half) * N/(2*half)
        const int tws = tw_shift0 - q;  // log2(N / (2*half))
#pragma unroll
        for (int e = 0; e < (1 << GG); ++e) {
            if (e & (1 << q)) continue;
            const int twidx = (low + (e & ((1 << q) - 1)) * h0) << tws;
            ntt_butterfly<SKIP_ONE>(x[e], x[e | (1 << q)], tw, twidx);
        }
    }
}

// butterfly stages per pass (2^G register-resident elements per thread).  G = 2 keeps the kernel at <= 80 registers so
// three CTAs (24 warps) stay resident per SM; N = 256 needs G = 3 to keep one item inside one warp.
#ifndef HB_NTT_G
#define HB_NTT_G 2
#endif
#ifndef HB_NTT_MINB
#define HB_NTT_MINB 3
#endif
#ifndef HB_NTT_BLOCK
#define HB_NTT_BLOCK 256
#endif
template <int LOGN>
__host__ __device__ constexpr int ntt_g() { return LOGN < HB_NTT_G ? LOGN : (LOGN >= 8 ? 3 : HB_NTT_G); }

// c = v * 2^-LOGN mod r without a product: r = 1 mod 2^32, so k = -v mod 2^LOGN makes v + k*r divisible by 2^LOGN, and
// (v + k*r) >> LOGN < r is the canonical quotient (8 narrow multiplies and a funnel shift instead of a 106-multiply CIOS product
// by the Montgomery form of 1/N; same value bit for bit).
template <int LOGN>
__device__ __forceinline__ void fr_div_pow2(uint32_t (&c)[8], const uint32_t (&v)[8]) {
    if (LOGN == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = v[i];
        return;
    }
    const uint32_t k = (0u - v[0]) & ((1u << LOGN) - 1u);
    const uint32_t rl[8] = {HB_R0, HB_R1, HB_R2, HB_R3, HB_R4, HB_R5, HB_R6, HB_R7};
    uint32_t t[9];
    unsigned long long acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc += (unsigned long long)k * rl[i] + v[i];
        t[i] = (uint32_t)acc;
        acc >>= 32;
    }
    t[8] = (uint32_t)acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = __funnelshift_r(t[i], t[i + 1], LOGN);
}

// one transformed value at natural-order position `pos` of item b
template <int MODE, int LOGN>
__device__ __forceinline__ void ntt_emit(const NttArgs &a, long long b, int pos, const uint32_t (&v)[8]) {
    if (MODE == 5) {
        if (pos < a.mout) {
            uint32_t c[8], cur[8], r[8];
            fr_div_pow2<LOGN>(c, v);
            const long long ob = (long long)a.item_list[b];
            uint4 *o = a.out + (ob * a.out_sb + (long long)pos * a.out_sr) * 2;
            load_fr(cur, o[0], o[1]);
            fr_sub(r, cur, c);
            o[0] = make_uint4(r[0], r[1], r[2], r[3]);
            o[1] = make_uint4(r[4], r[5], r[6], r[7]);
        }
    } else if (MODE == 3) {
        if (fr_is_zero(v) && a.idset[pos] >= 0) atomicOr(a.rootmask + b * 8 + (pos >> 5), 1u << (pos & 31));
    } else if (MODE == 4) {
        const unsigned int *mk = a.rootmask + b * 8;
        const unsigned int w = mk[pos >> 5];
        if ((w >> (pos & 31)) & 1u) {
            int q = __popc(w & ((1u << (pos & 31)) - 1u));
            for (int i = 0; i < (pos >> 5); ++i) q += __popc(mk[i]);
            if (q < a.cols) {  // a locator has at most cols - 1 roots; anything else is a slot the decoder has given up on
                uint4 *o = a.out + (b * a.out_sb + (long long)q * a.out_sr) * 2;
                o[0] = make_uint4(v[0], v[1], v[2], v[3]);
                o[1] = make_uint4(v[4], v[5], v[6], v[7]);
            }
        }
    } else if (MODE == 0) {
        if (pos < a.n) {
            uint4 *o = a.out + (b * a.out_sb + (long long)pos * a.out_sr) * 2;
            stg_stream(o, make_uint4(v[0], v[1], v[2], v[3]));
            stg_stream(o + 1, make_uint4(v[4], v[5], v[6], v[7]));
        }
    } else {
        if (pos < a.mout) {
            uint32_t c[8];
            if (MODE == 1) fr_div_pow2<LOGN>(c, v);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) c[i] = v[i];
            }
            if (MODE == 2 && a.out_group32) {
                uint4 *o = a.out + (((b >> 5) * a.out_sb + (long long)pos) * 2) * 32 + (b & 31);
                o[0] = make_uint4(c[0], c[1], c[2], c[3]);
                o[32] = make_uint4(c[4], c[5], c[6], c[7]);
            } else {
                uint4 *o = a.out + (b * a.out_sb + (long long)pos * a.out_sr) * 2;
                stg_stream(o, make_uint4(c[0], c[1], c[2], c[3]));
                stg_stream(o + 1, make_uint4(c[4], c[5], c[6], c[7]));
            }
        } else if (pos >= a.m) {
            if (!fr_is_zero(v)) mark_fail(a.fail, b, a.fail_list, a.fail_count);
        } else if (MODE == 2 && pos <= a.hi_top && pos > a.hi_top - a.hi_cnt) {
            uint4 *o = a.out + (b * a.out_sb + (long long)(a.mout + a.hi_top - pos) * a.out_sr) * 2;
            stg_stream(o, make_uint4(v[0], v[1], v[2], v[3]));
            stg_stream(o + 1, make_uint4(v[4], v[5], v[6], v[7]));
        }
    }
}

template <int LOGN, int MODE>
__global__ void __launch_bounds__(HB_NTT_BLOCK, HB_NTT_MINB) ntt_kernel(const NttArgs a) {
    constexpr int N = 1 << LOGN;
    constexpr int G = ntt_g<LOGN>();
    constexpr int E = 1 << G;
    constexpr int TPI = N / E;                  // threads per item (<= 32)
    constexpr int IPC = HB_NTT_BLOCK / TPI;     // items per CTA tile
    constexpr int NP = (LOGN + G - 1) / G;      // passes
    constexpr int GL = LOGN - G * (NP - 1);     // stages in the last pass
    constexpr int PADN = N + N / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *sTw = reinterpret_cast<uint4 *>(smem_raw);            // [N/2][2]
    fma_ballast(a.B < 0, a.err);
    uint4 *sD = sTw + (N > 1 ? N : 2);                           // [IPC][2][PADN]   (NP > 1 only)
    uint4 *sIn = sD + (NP > 1 ? (size_t)IPC * 2 * PADN : 0);     // [E][2][BLOCK] per-thread input staging (prefetch)
    for (int i = threadIdx.x; i < N; i += blockDim.x) sTw[i] = a.tw[i];  // N/2 entries * 2 halves
    __syncthreads();

    const int tid_i = threadIdx.x % TPI, item_l = threadIdx.x / TPI;
    uint4 *myD = sD + (size_t)item_l * 2 * PADN;
    // A tile is the IPW items of ONE warp (the warps of a CTA only share the twiddle table): warps fetch their tiles from a global
    // queue (a.work, see ntt16x.cuh: the scheduler's priorities make statically assigned warps finish at very different times), one
    // tile ahead of the prefetch, i.e. two ahead of the transform; a.work == nullptr: static round-robin.
    constexpr int IPW = 32 / TPI, WPC = HB_NTT_BLOCK / 32;
    const int lane = threadIdx.x & 31, item_w = lane / TPI;
    long long Beff = a.B;
    if (a.b_dev) {
        const unsigned int c = *a.b_dev;
        const long long live = c > a.b_first ? (long long)(c - a.b_first) : 0;
        Beff = live < Beff ? live : Beff;
    }
    const long long ntiles = (Beff + IPW - 1) / IPW;
    const long long nwarps = (long long)gridDim.x * WPC;
    unsigned bad = 0;

    // Which record feeds each of this thread's E positions is the same for every tile: position pos holds the input with
    // natural index k = bitrev(pos) (zero beyond `cols` / outside the examined id set).
    int rec_of[E], k_of[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int pos = tid_i * E + e;
        const int k = (int)(__brev((unsigned)pos) >> (32 - LOGN));
        k_of[e] = k;
        rec_of[e] = (MODE != 5 && k < a.cols) ? (((MODE == 1 || MODE == 2) && a.in_map) ? a.in_map[k] : k) : -1;
    }
    uint4 *myIn = sIn + threadIdx.x;
    // `gate` carries a data dependence on the values just read from the staging slots, so the asynchronous copies that
    // overwrite those slots cannot be issued before the reads have completed
    const unsigned int never = (unsigned int)a.n + 0x7fff0000u;
    auto prefetch = [&](long long t, unsigned int gate) {
        const long long bb = t * IPW + item_w;
        if (t < ntiles && bb < Beff && gate != never) {
            const long long src = (MODE == 2 && a.item_list) ? (long long)a.item_list[bb] : bb;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (rec_of[e] < 0) continue;
                const uint4 *p = a.in + (src * a.in_sb + (long long)rec_of[e] * a.in_sc) * 2;
                cp_async16(myIn + (e * 2) * HB_NTT_BLOCK, p);
                cp_async16(myIn + (e * 2 + 1) * HB_NTT_BLOCK, p + 1);
            }
        }
        cp_async_commit();
    };
    long long tile = (long long)blockIdx.x * WPC + (threadIdx.x >> 5);
    long long next = tile + nwarps;
    if (a.work) {
        unsigned long long q = 0;
        if (lane == 0) q = atomicAdd(a.work, 1ull);
        next = nwarps + (long long)__shfl_sync(0xffffffffu, q, 0);
    }
    prefetch(tile, 0u);

    while (tile < ntiles) {
        unsigned long long q2 = 0;
        if (a.work && lane == 0) q2 = atomicAdd(a.work, 1ull);   // the tile after `next`; consumed at the end of this iteration
        const long long b = tile * IPW + item_w;
        const bool active = b < Beff;
        uint32_t x[E][8];
        // ---- pass 0: inputs from the prefetch staging (thread-private slots: no barrier), stages with half = 1, 2, 4
        cp_async_wait_all();
        unsigned dep = 0;  // MODE 3/4: data dependence of the next prefetch on the staged values (see `gate`)
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if (active && rec_of[e] >= 0) {
                const int k = k_of[e];
                load_fr(x[e], myIn[(e * 2) * HB_NTT_BLOCK], myIn[(e * 2 + 1) * HB_NTT_BLOCK]);
                if (MODE <= 2) bad |= geq_mod(x[e]) ? 1u : 0u;  // MODE 3/4 read decoder state, not user data
                else dep |= (x[e][0] ^ x[e][7]) >> 31;
                if (MODE == 2) {
                    uint32_t w[8], y[8];
                    load_fr(w, __ldg(a.wt + k * 2), __ldg(a.wt + k * 2 + 1));
                    mont_mul(y, x[e], w);
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[e][i] = y[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[e][i] = 0;
                if (MODE == 5 && active) {  // sparse input: the q-th set bit of rootmask[b] carries in[b][q] at its domain position
                    const int k = k_of[e];
                    const unsigned int *mk = a.rootmask + b * 8;
                    const unsigned int w = mk[k >> 5];
                    if ((w >> (k & 31)) & 1u) {
                        int q = __popc(w & ((1u << (k & 31)) - 1u));
                        for (int i = 0; i < (k >> 5); ++i) q += __popc(mk[i]);
                        const uint4 *p = a.in + (b * a.in_sb + (long long)q * a.in_sc) * 2;
                        load_fr(x[e], p[0], p[1]);
                    }
                }
            }
        }
        if ((MODE == 1 || MODE == 2) && a.path && active && tid_i == 0) a.path[b] = 0;
        // the staged values are in registers (and were examined by geq_mod): the slots can take the next tile's inputs
        prefetch(next, bad | dep);
        ntt_stages<G, E, true>(x, sTw, 0, 1, LOGN - 1);
        if constexpr (NP == 1) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int pos = tid_i * E + e;
                if (active) ntt_emit<MODE, LOGN>(a, b, pos, x[e]);
            }
        } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int pos = tid_i * E + e, idx = pos + (pos >> 3);
            myD[idx] = make_uint4(x[e][0], x[e][1], x[e][2], x[e][3]);
            myD[PADN + idx] = make_uint4(x[e][4], x[e][5], x[e][6], x[e][7]);
        }
        __syncwarp();
        // ---- middle passes (G stages each)
#pragma unroll
        for (int p = 1; p < NP - 1; ++p) {
            const int sh = G * p, h0 = 1 << sh;
            const int low = tid_i & (h0 - 1), base = ((tid_i >> sh) << (sh + G)) | low;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int pos = base + e * h0, idx = pos + (pos >> 3);
                load_fr(x[e], myD[idx], myD[PADN + idx]);
            }
            ntt_stages<G, E, false>(x, sTw, low, h0, LOGN - 1 - sh);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int pos = base + e * h0, idx = pos + (pos >> 3);
                myD[idx] = make_uint4(x[e][0], x[e][1], x[e][2], x[e][3]);
                myD[PADN + idx] = make_uint4(x[e][4], x[e][5], x[e][6], x[e][7]);
            }
            __syncwarp();
        }
        // ---- last pass (GL stages): N >> GL units per item, natural-order outputs to global
        {
            constexpr int sh = G * (NP - 1), h0 = 1 << sh, EL = 1 << GL;
#pragma unroll 1
            for (int u = tid_i; u < (N >> GL); u += TPI) {
                const int low = u & (h0 - 1), base = ((u >> sh) << (sh + GL)) | low;
                uint32_t y[EL][8];
#pragma unroll
                for (int e = 0; e < EL; ++e) {
                    const int pos = base + e * h0, idx = pos + (pos >> 3);
                    load_fr(y[e], myD[idx], myD[PADN + idx]);
                }
                ntt_stages<GL, EL, false>(y, sTw, low, h0, LOGN - 1 - sh);
#pragma unroll
                for (int e = 0; e < EL; ++e) {
                    const int pos = base + e * h0;
                    if (active) ntt_emit<MODE, LOGN>(a, b, pos, y[e]);
                }
            }
            __syncwarp();  // the tile's buffer is reused by the next tile's pass 0 writes
        }
        }  // NP > 1
        tile = next;
        next = a.work ? nwarps + (long long)__shfl_sync(0xffffffffu, q2, 0) : next + nwarps;
    }
    if (a.work && lane == 0) {   // the last warp to leave resets the queue for the next launch on this stream
        if (atomicAdd(a.work + 1, 1ull) == (unsigned long long)(nwarps - 1)) {
            a.work[0] = 0ull;
            a.work[1] = 0ull;
        }
    }
    if (bad) *(volatile unsigned int *)a.err = 1u;  // mapped host memory: plain store, every writer stores 1
}

// ------------------------------------------------------------------------------------------------------------------
// ntt64_cta_kernel<MODE>: the 64-point transform (the headline shape: n = 64) with the work of a 16-item tile regrouped
// ACROSS the warps of the CTA, so that twiddle factors that are trivial -- or operands that are known to be zero -- are so
// for every lane of a warp and the product is skipped by the whole warp (ntt_kernel keeps one item inside one warp: there a
// trivial twiddle only idles a quarter of the lanes and saves nothing).
//   pass 0 (stages 0,1): thread = (item, r): the 16 thread indices are ordered so that those whose only product (w^16 * x3) has a
//                        possibly non-zero operand come first: with 22 coefficients of 64 only 3 of the 8 warps multiply
//   pass 1 (stages 2,3): thread = (item, hi, low) with `low` (the twiddle index) constant per warp: the two warps with low = 0
//                        run one product instead of four
//   pass 2 (stages 4,5): thread = (item, tid) as in ntt_kernel (lane-dependent twiddles, coalesced natural-order outputs)
// Passes exchange through shared memory with __syncthreads (3 per tile); items are `ISTR` = odd number of uint4 apart so that
// the 8 items of a quarter warp fall into different banks.  MODE 0 / 1 as in ntt_kernel; bit-identical results.
// Executed products per item: 122 (MODE 0, 22 coefficients) / 132 (all 64 inputs) instead of 144.
#ifndef HB_NTT64_ROT
#define HB_NTT64_ROT 1   // rotate the warps' roles from tile to tile (0: fixed roles, the first version)
#endif
#ifndef HB_NTT64_IPC
#define HB_NTT64_IPC 8   // items per CTA tile (8: four warps per barrier, six CTAs per SM; 16: eight warps, three CTAs)
#endif
template <int MODE>
__global__ void __launch_bounds__(HB_NTT64_IPC * 16, 768 / (HB_NTT64_IPC * 16)) ntt64_cta_kernel(const NttArgs a) {
    constexpr int LOGN = 6, N = 64, E = 4, IPC = HB_NTT64_IPC, PADN = N + N / 8, ISTR = 2 * PADN + 1, BLOCK = IPC * 16;
    constexpr int LI = IPC == 16 ? 4 : 3;  // log2(IPC)
    fma_ballast(a.B < 0, a.err);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *sTw = reinterpret_cast<uint4 *>(smem_raw);   // [N/2][2]
    uint4 *sD = sTw + N;                                 // [IPC][ISTR]
    uint4 *sIn = sD + (size_t)IPC * ISTR;                // [E][2][BLOCK] per-thread input staging (prefetch)
    const int t = threadIdx.x;
    for (int i = t; i < N; i += BLOCK) sTw[i] = a.tw[i];

    // ---- roles.  A thread's share of the tile's work depends on its LOGICAL index tl = lw*32 + lane, where the logical warp
    // lw = (warp + tile iteration) & 3 rotates from tile to tile: the light roles (the warp whose pass-1 twiddle is trivial, the
    // warps whose pass-0 product has zero operands) would otherwise always sit on the same SM sub-partition (warp w of every
    // CTA is scheduled on sub-partition w & 3) and leave its multiplier idle while the other three are the bottleneck.
    // pass 0: position pos = tid*4 + e holds the input with natural index bitrev(pos); the product of pass 0 is w^16 * x3 where
    // x3 = (input e=2) - (input e=3): thread indices whose inputs 2 and 3 do not exist are ranked last (sRank / sRec tables).
    int *sRank = reinterpret_cast<int *>(sIn + (size_t)E * 2 * BLOCK);   // [16] rank -> pass-0 thread index
    int *sRec = sRank + 16;                                             // [16][4] pass-0 thread index -> record of element e (-1: zero)
    if (t < 16) {
        auto rec_for = [&](int tid, int e) -> int {
            const int k = (int)(__brev((unsigned)(tid * E + e)) >> (32 - LOGN));
            return (k < a.cols) ? ((MODE == 1 && a.in_map) ? a.in_map[k] : k) : -1;
        };
        int rank = 0, found = -1;
        for (int pass = 0; pass < 2 && found < 0; ++pass)      // first the indices that need the product, then the others
            for (int c = 0; c < 16 && found < 0; ++c) {
                const bool need = rec_for(c, 2) >= 0 || rec_for(c, 3) >= 0;
                if (need == (pass == 0)) {
                    if (rank == t) found = c;
                    ++rank;
                }
            }
        sRank[t] = found;
#pragma unroll
        for (int e = 0; e < E; ++e) sRec[t * E + e] = rec_for(t, e);
    }
    const int lane = t & 31, warp = t >> 5;
    constexpr int NW = BLOCK / 32;
    const unsigned FULL = 0xffffffffu;
    // ---- pass-2 role (every warp does the same amount of work there: no rotation)
    const int item2 = t >> 4, tid2 = t & 15;        // positions tid2 + e*16
    __syncthreads();

    const long long ntiles = (a.B + IPC - 1) / IPC;
    unsigned bad = 0;
    uint4 *myIn = sIn + t;
    const unsigned int never = (unsigned int)a.n + 0x7fff0000u;
    // role of this thread in tile iteration `it`: logical thread index
    auto logical = [&](unsigned it) -> int { return HB_NTT64_ROT ? ((((warp + (int)it) & (NW - 1)) << 5) | lane) : t; };
    auto prefetch = [&](long long tl, unsigned it, unsigned int gate) {
        const int tlg = logical(it);
        const long long bb = tl * IPC + (tlg & (IPC - 1));
        if (tl < ntiles && bb < a.B && gate != never) {
            const int c0 = sRank[tlg >> LI];
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int rec = sRec[c0 * E + e];
                if (rec < 0) continue;
                const uint4 *p = a.in + (bb * a.in_sb + (long long)rec * a.in_sc) * 2;
                cp_async16(myIn + (e * 2) * BLOCK, p);
                cp_async16(myIn + (e * 2 + 1) * BLOCK, p + 1);
            }
        }
        cp_async_commit();
    };
    auto bfly = [&](uint32_t (&u)[8], uint32_t (&v)[8], int twidx, bool mul) {  // mul == false: twiddle 1, or v known to be zero
        uint32_t tt[8], sm[8], df[8];
        if (mul) {
            uint32_t w[8];
            load_fr(w, sTw[twidx * 2], sTw[twidx * 2 + 1]);
            mont_mul(tt, v, w);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) tt[i] = v[i];
        }
        fr_add(sm, u, tt);
        fr_sub(df, u, tt);
#pragma unroll
        for (int i = 0; i < 8; ++i) { u[i] = sm[i]; v[i] = df[i]; }
    };
    auto idx_of = [](int pos) { return pos + (pos >> 3); };
    prefetch(blockIdx.x, 0u, 0u);

    unsigned it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        uint32_t x[E][8];
        const int tlg = logical(it);
        const int item0 = tlg & (IPC - 1), tid0 = sRank[tlg >> LI];
        int rec_of[E];
#pragma unroll
        for (int e = 0; e < E; ++e) rec_of[e] = sRec[tid0 * E + e];
        const bool has1 = __any_sync(FULL, rec_of[1] >= 0), has2 = __any_sync(FULL, rec_of[2] >= 0), has3 = __any_sync(FULL, rec_of[3] >= 0);
        const int item1 = item0, low1 = tlg >> (LI + 2), tid1 = (((tlg >> LI) & 3) << 2) | low1;
        const int base1 = ((tid1 >> 2) << 4) | low1;   // pass-1 positions base1 + e*4; low1 (the twiddle index) is constant per warp
        // ---- pass 0
        {
            const long long b = tile * IPC + item0;
            const bool active = b < a.B;
            cp_async_wait_all();
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (active && rec_of[e] >= 0) {
                    load_fr(x[e], myIn[(e * 2) * BLOCK], myIn[(e * 2 + 1) * BLOCK]);
                    bad |= geq_mod(x[e]) ? 1u : 0u;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[e][i] = 0;
                }
            }
            prefetch(tile + gridDim.x, it + 1, bad);
            // stage 0: (x0,x1), (x2,x3) with twiddle 1; stage 1: (x0,x2) with 1, (x1,x3) with w^16
            if (has1) bfly(x[0], x[1], 0, false);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[1][i] = x[0][i];
            }
            if (has3) bfly(x[2], x[3], 0, false);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[3][i] = x[2][i];
            }
            if (has2 || has3) {
                bfly(x[0], x[2], 0, false);
                bfly(x[1], x[3], N / 4, true);
            } else {  // x2 == x3 == 0 in every lane of the warp
#pragma unroll
                for (int i = 0; i < 8; ++i) { x[2][i] = x[0][i]; x[3][i] = x[1][i]; }
            }
            uint4 *d = sD + (size_t)item0 * ISTR;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int ix = idx_of(tid0 * E + e);
                d[ix] = make_uint4(x[e][0], x[e][1], x[e][2], x[e][3]);
                d[PADN + ix] = make_uint4(x[e][4], x[e][5], x[e][6], x[e][7]);
            }
        }
        __syncthreads();
        // ---- pass 1: half = 4, 8; twiddle exponent (pos mod half) * N/(2*half)
        {
            uint4 *d = sD + (size_t)item1 * ISTR;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int ix = idx_of(base1 + e * 4);
                load_fr(x[e], d[ix], d[PADN + ix]);
            }
            const bool nz = low1 != 0;  // warp-uniform
            bfly(x[0], x[1], low1 << 3, nz);
            bfly(x[2], x[3], low1 << 3, nz);
            bfly(x[0], x[2], low1 << 2, nz);
            bfly(x[1], x[3], (low1 + 4) << 2, true);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int ix = idx_of(base1 + e * 4);
                d[ix] = make_uint4(x[e][0], x[e][1], x[e][2], x[e][3]);
                d[PADN + ix] = make_uint4(x[e][4], x[e][5], x[e][6], x[e][7]);
            }
        }
        __syncthreads();
        // ---- pass 2: half = 16, 32; natural-order outputs
        {
            const long long b = tile * IPC + item2;
            const bool active = b < a.B;
            const uint4 *d = sD + (size_t)item2 * ISTR;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int ix = idx_of(tid2 + e * 16);
                load_fr(x[e], d[ix], d[PADN + ix]);
            }
            if (MODE == 1 && a.path && active && tid2 == 0) a.path[b] = 0;
            bfly(x[0], x[1], tid2 << 1, true);
            bfly(x[2], x[3], tid2 << 1, true);
            bfly(x[0], x[2], tid2, true);
            bfly(x[1], x[3], tid2 + 16, true);
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (active) ntt_emit<MODE, LOGN>(a, b, tid2 + e * 16, x[e]);
        }
        __syncthreads();  // the tile's buffer is rewritten by the next tile's pass 0
    }
    if (bad) *(volatile unsigned int *)a.err = 1u;
}
inline size_t ntt64_cta_smem_bytes() { return (size_t)(64 + HB_NTT64_IPC * (2 * 72 + 1) + 4 * 2 * HB_NTT64_IPC * 16) * 16 + 16 + 5 * 16 * 4; }

template <int LOGN>
inline size_t ntt_smem_bytes() {
    constexpr int N = 1 << LOGN;
    constexpr int G = ntt_g<LOGN>();
    constexpr int TPI = N / (1 << G);
    constexpr int IPC = HB_NTT_BLOCK / TPI;
    constexpr int NP = (LOGN + G - 1) / G;
    size_t tw = (size_t)(N > 1 ? N : 2) * 16;
    return tw + (NP > 1 ? (size_t)IPC * 2 * (N + N / 8) * 16 : 0) + (size_t)(1 << G) * 2 * HB_NTT_BLOCK * 16 + 16;
}
template <int LOGN>
inline int ntt_items_per_cta() {
    return HB_NTT_BLOCK / ((1 << LOGN) / (1 << ntt_g<LOGN>()));
}

}  // namespace hb

#undef mont_mul
#undef fr_add
