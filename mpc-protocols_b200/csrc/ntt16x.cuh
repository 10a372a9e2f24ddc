// ntt16x.cuh -- the domain transforms of K1 / K2 / K3 for N = 16 .. 128: no CTA barrier, warp-uniform twiddles, compact code.
//
// Same mathematics as ntt.cuh (radix-2 decimation in time over the reference's evaluation domain, robust_interpolate.rs:68-79,
// common/share/mod.rs:31-45; data canonical, twiddles in Montgomery form), different mapping.  N = S*L with S = 8 (16 for N = 128)
// rows.  L lanes of a warp own one item (32/L items per warp); the item's S x L working set lives in the warp's tile of shared
// memory as S rows x 32 lane-columns of 32-byte elements:
//   P0  lane j computes, alone and in its own column, the S-point transform of the inputs k = L*i + bitrev(j).  Every lane of the
//       warp runs the same butterfly at the same time, so twiddle indices -- and which operands are structurally zero (share
//       generation feeds d+1 coefficients into N points) -- are warp-uniform: trivial twiddles and zero operands cost nothing,
//       for every lane, without divergence.  No synchronisation: the column is private.
//   P1  lane j then owns the rows q = L*e + j of its item and runs the remaining log2(L) stages in place along each row (the L
//       columns of the item), emitting natural-order outputs from the last stage.  One __syncwarp separates P0 and P1.
// What bounds these kernels (profiles/r02_*): the Montgomery product alone runs at the multiplier pipe's limit with two warps per
// sub-partition (73 G products/s, tools/probe.py), but a transform issues ~2.4 other instructions per wide multiply-add -- carry
// chains of the butterfly's add / subtract / conditional subtractions, shared-memory traffic, index arithmetic -- and the issue
// slot, not the pipe, runs out.  So the schedule is resolved at COMPILE time (fully unrolled loops over stages and butterflies: no
// index arithmetic, no flag tests at run time) while the arithmetic lives in a handful of out-of-line butterfly routines
// (`bf_*`, __noinline__: one copy of the product in the instruction cache, ~25 KB of code against the 55 KB of the fully inlined
// kernels of ntt.cuh) that take shared-memory addresses in registers.
// Last stage of the inverse transform (MODE 1/2): a butterfly whose two outputs both lie in the must-vanish range needs no
// product -- u + w*v = u - w*v = 0  <=>  u = v = 0 -- and outputs that are neither stored nor checked are not computed.
// Executed products per item, N = 64: 121 (share generation, 22 coefficients), 119 (inverse + degree check, m = 22), 129 (full).
// Row p of the tile stores column c at c ^ (p & (L-1)): conflict-free 128-bit accesses in P0 (whole rows) and in P1 (the lanes
// of a quarter warp read different bank groups).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fr.cuh"
#include "ntt.cuh"

namespace hb {

#ifndef HB_NTT16X_LOGS
#define HB_NTT16X_LOGS 3     // log2(rows) for N = 32, 64
#endif
#ifndef HB_NTT16X_MINB8
#define HB_NTT16X_MINB8 6    // resident CTAs (of 4 warps) per SM asked of ptxas for the 8-row tiles: bounds the registers per thread
#endif
#ifndef HB_NTT16X_LOGS16
#define HB_NTT16X_LOGS16 3   // log2(rows) for N = 16
#endif
#ifndef HB_NTT16X_LOGS128
#define HB_NTT16X_LOGS128 3  // log2(rows) for N = 128.  3: 8 rows x 16 lanes per item, two lanes share a row in P1 (8 KB of tile per
#endif                       // warp, 6 CTAs per SM); 4: 16 rows x 8 lanes (16 KB per warp: shared memory allows 3 CTAs per SM)
constexpr int NTT16X_WARPS = 4;

template <int LOGN>
__host__ __device__ constexpr int ntt16x_logs() { return LOGN == 7 ? HB_NTT16X_LOGS128 : (LOGN == 4 ? HB_NTT16X_LOGS16 : HB_NTT16X_LOGS); }
template <int LOGN>
__host__ __device__ constexpr int ntt16x_minb() { return ntt16x_logs<LOGN>() == 3 ? HB_NTT16X_MINB8 : 3; }

template <int LOGN, int MODE>
inline size_t ntt16x_smem_bytes() {
    constexpr int N = 1 << LOGN, S = 1 << ntt16x_logs<LOGN>(), L = N / S;
    return (size_t)((N / 2) * 2 + (L - 1) * S * 2 + (MODE == 2 ? N * 2 : 0) + NTT16X_WARPS * S * 2 * 32) * 16 + 16;
}

// ---- shared-memory accesses by 32-bit window address (the butterfly routines are real calls: generic pointers would turn
// every access into a generic LD / ST)
__device__ __forceinline__ void lds_fr(uint32_t (&x)[8], unsigned addr) {   // halves 512 bytes apart ([row][half][32 lanes] uint4)
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]) : "r"(addr));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+512];" : "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]) : "r"(addr));
}
__device__ __forceinline__ void sts_fr(unsigned addr, const uint32_t (&x)[8]) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]) : "memory");
    asm volatile("st.shared.v4.u32 [%0+512], {%1,%2,%3,%4};" ::"r"(addr), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]) : "memory");
}
__device__ __forceinline__ void lds_tw(uint32_t (&w)[8], unsigned addr) {   // table entries: two consecutive uint4
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(addr));
}

// ---- out-of-line butterflies on tile addresses.  Return value (bf_add / bf_copy / bf_scale): 1 when a non-canonical (>= r) input was
// seen (chk != 0).  Values are validated where they are first touched: stage 0, whose twiddles are all trivial, so the product
// routine itself never checks and returns nothing.
// (u, v) <- (u + w*v, u - w*v)
__device__ __noinline__ void bf_mul(unsigned pu, unsigned pv, unsigned pw) {
    uint32_t u[8], v[8], w[8], t[8], sm[8], df[8];
    lds_fr(v, pv);
    lds_tw(w, pw);
    lds_fr(u, pu);
    mont_mul(t, v, w);
    fr_add(sm, u, t);
    fr_sub(df, u, t);
    sts_fr(pu, sm);
    sts_fr(pv, df);
}
// (u, v) <- (u + v, u - v)   (trivial twiddle)
__device__ __noinline__ unsigned bf_add(unsigned pu, unsigned pv, unsigned chk) {
    uint32_t u[8], v[8], sm[8], df[8];
    lds_fr(v, pv);
    lds_fr(u, pu);
    unsigned bad = 0;
    if (chk) bad = (geq_mod(u) || geq_mod(v)) ? 1u : 0u;
    fr_add(sm, u, v);
    fr_sub(df, u, v);
    sts_fr(pu, sm);
    sts_fr(pv, df);
    return bad;
}
// v <- u   (v structurally zero: the butterfly with any twiddle is a copy)
__device__ __forceinline__ unsigned bf_copy(unsigned pu, unsigned pv, unsigned chk) {
    uint32_t u[8];
    lds_fr(u, pu);
    sts_fr(pv, u);
    return chk ? (geq_mod(u) ? 1u : 0u) : 0u;
}
// v <- w*v   (MODE 2: erasure weights)
__device__ __noinline__ unsigned bf_scale(unsigned pv, unsigned pw) {
    uint32_t v[8], w[8], t[8];
    lds_fr(v, pv);
    lds_tw(w, pw);
    const unsigned bad = geq_mod(v) ? 1u : 0u;
    mont_mul(t, v, w);
    sts_fr(pv, t);
    return bad;
}
// last stage: the two results leave the tile.  MODE 0: out[pos] for pos < lim0 = n.  MODE 1/2: pos < lim0 = mout stored (MODE 1:
// scaled by 1/N), pos >= lim1 = m must vanish (bit 1 of the return value set otherwise), anything between is not wanted.
// mul == 0: trivial twiddle.
template <int LOGN, int MODE>
__device__ __noinline__ unsigned bf_emit(unsigned pu, unsigned pv, unsigned pw, int mul, uint4 *out, long long out_sr, int posU, int posV, int lim0p, int lim1) {
    // MODE 2: lim0p packs (mout, hi_cnt << 8, hi_top << 16): positions hi_top - hi_cnt < pos <= hi_top are stored too, at out index
    // mout + hi_top - pos (NttArgs::hi_top / hi_cnt: the reversed top of the product polynomial)
    const int lim0 = MODE == 2 ? (lim0p & 0xff) : lim0p;
    const int hi_cnt = MODE == 2 ? ((lim0p >> 8) & 0xff) : 0, hi_top = MODE == 2 ? (lim0p >> 16) : -1;
    uint32_t u[8], v[8], t[8], sm[8], df[8];
    unsigned rc = 0;
    if (MODE != 0) {
        const bool hiU = MODE == 2 && posU <= hi_top && posU > hi_top - hi_cnt, hiV = MODE == 2 && posV <= hi_top && posV > hi_top - hi_cnt;
        const bool needU = posU < lim0 || posU >= lim1 || hiU, needV = posV < lim0 || posV >= lim1 || hiV;
        if (!needU && !needV) return 0u;
        if (posU >= lim1) {   // both must vanish <=> u == 0 and v == 0: no product
            lds_fr(u, pu);
            lds_fr(v, pv);
            return (fr_is_zero(u) && fr_is_zero(v)) ? 0u : 2u;
        }
    } else if (posU >= lim0 && posV >= lim0) return 0u;
    lds_fr(v, pv);
    lds_fr(u, pu);
    if (mul) {
        uint32_t w[8];
        lds_tw(w, pw);
        mont_mul(t, v, w);
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = v[i];
    }
    fr_add(sm, u, t);
    fr_sub(df, u, t);
#pragma unroll 1
    for (int z = 0; z < 2; ++z) {
        const int pos = z ? posV : posU;
        uint32_t val[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) val[i] = z ? df[i] : sm[i];
        const bool hi = MODE == 2 && pos <= hi_top && pos > hi_top - hi_cnt && pos >= lim0 && pos < lim1;
        if (pos < lim0 || hi) {
            uint32_t c[8];
            if (MODE == 1) fr_div_pow2<LOGN>(c, val);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) c[i] = val[i];
            }
            uint4 *o = out + (long long)(hi ? lim0 + hi_top - pos : pos) * out_sr * 2;
            stg_stream(o, make_uint4(c[0], c[1], c[2], c[3]));
            stg_stream(o + 1, make_uint4(c[4], c[5], c[6], c[7]));
        } else if (MODE != 0 && pos >= lim1) {
            if (!fr_is_zero(val)) rc |= 2u;
        }
    }
    return rc;
}

template <int LOGN, int MODE>
__global__ void __launch_bounds__(NTT16X_WARPS * 32, ntt16x_minb<LOGN>()) ntt16x_kernel(const NttArgs a) {
    static_assert(LOGN >= 4 && LOGN <= 7, "N = 16 .. 128");
    static_assert(MODE >= 0 && MODE <= 2, "forward, inverse + degree check, weighted inverse");
    fma_ballast(a.B < 0, a.err);
    constexpr int LOGS = ntt16x_logs<LOGN>(), S = 1 << LOGS;
    constexpr int N = 1 << LOGN, LL = LOGN - LOGS, L = 1 << LL, IPW = 32 / L;
    static_assert(L <= 32 && (L <= S || L % S == 0), "P1: a lane owns S/L rows, or L/S lanes share one row");
    constexpr int HALVES = L > S ? L / S : 1;                         // P1: lanes sharing one row (each takes 1/HALVES of a stage's butterflies)
    constexpr int ROWS = L > S ? 1 : S / L;                           // P1: rows per lane
    constexpr int BFL = (L / 2) / HALVES;                             // P1: butterflies per lane and stage
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *sTw = reinterpret_cast<uint4 *>(smem_raw);                 // [N/2][2]  w^k (w^-k for the inverse), Montgomery form
    // P1's twiddles depend on the lane's row q: in sTw the entries of one (stage, lo) for q = 0 .. S-1 lie 32 << (LL-1-s2) bytes
    // apart -- the same bank group for every lane in the early stages (ncu: 16.7 wavefronts per LDS.128, 2/3 of the kernel's
    // excess shared-memory wavefronts).  sTw1 holds them contiguous in q: block (2^s2 - 1 + lo) of S entries.
    uint4 *sTw1 = sTw + N;                                            // [(L-1)*S][2]
    uint4 *sWt = sTw1 + (L - 1) * S * 2;                              // [N][2]    MODE 2: weights by domain index
    uint4 *sD = sWt + (MODE == 2 ? 2 * N : 0);                        // [warps][S rows][2 halves][32 columns]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < N; i += blockDim.x) sTw[i] = a.tw[i];
    for (int i = t; i < (L - 1) * S * 2; i += blockDim.x) {
        const int half = i & 1, e = i >> 1, q = e & (S - 1), blk = e >> LOGS;          // blk = 2^s2 - 1 + lo
        const int s2 = 31 - __clz(blk + 1), lo = blk + 1 - (1 << s2);
        sTw1[i] = a.tw[(((S * lo + q) << (LL - 1 - s2)) << 1) + half];
    }
    if (MODE == 2)
        for (int i = t; i < 2 * N; i += blockDim.x) sWt[i] = a.wt[i];
    __syncthreads();
    uint4 *D = sD + (size_t)warp * (S * 2 * 32);
    const unsigned dD = (unsigned)__cvta_generic_to_shared(D), dTw = (unsigned)__cvta_generic_to_shared(sTw),
                   dTw1 = (unsigned)__cvta_generic_to_shared(sTw1), dWt = (unsigned)__cvta_generic_to_shared(sWt);
    const int j = lane & (L - 1), item = lane >> LL;
    const int cj = (LL == 0) ? 0 : (int)(__brev((unsigned)j) >> (32 - (LL == 0 ? 1 : LL)));   // residue class of this lane's inputs
    const int cols = a.cols < N ? a.cols : N;
    const int cnt = min(S, (cols + L - 1) / L);                      // inputs per column that may be non-zero (warp-uniform)
    unsigned mask0 = 0;                                               // rows that hold a (possibly) non-zero value after loading
    for (int i = 0; i < cnt; ++i) mask0 |= 1u << (__brev((unsigned)i) >> (32 - LOGS));
    // window address of element (row p, column c); half 1 at +512
    auto slot = [&](int p, int c) -> unsigned { return dD + (unsigned)(((p * 2) * 32 + (c ^ (p & (L - 1)))) * 16); };
    const int lim0 = MODE == 0 ? a.n : (MODE == 2 ? (a.mout | (a.hi_cnt << 8) | (a.hi_top << 16)) : a.mout), lim1 = a.m;

    const long long ntiles = (a.B + IPW - 1) / IPW;
    unsigned bad = 0;
    // Tiles are handed out dynamically: the warp scheduler favours some warps of a sub-partition over others (measured: with a
    // static round-robin the favoured warps run out of tiles at ~60 % of the kernel's duration and the sub-partition finishes with
    // 3-4 of its 6 warps, 4.6 on average), so every warp fetches its next tile from a global counter while it works on the
    // current one (one 64-bit atomic per tile, its latency hidden behind the transform).
    const long long nwarps = (long long)gridDim.x * NTT16X_WARPS;
    long long tile = (long long)blockIdx.x * NTT16X_WARPS + warp;
    while (tile < ntiles) {
        unsigned long long nxt = 0;
        if (a.work && lane == 0) nxt = atomicAdd(a.work, 1ull);
        const long long b = tile * IPW + item;
        const int active = b < a.B ? 1 : 0;
        // ---- inputs: lane j takes k = L*i + cj to row bitrev(i) of its own column (asynchronous copies, no registers); rows that
        // receive nothing are zeroed, so a butterfly may always read both operands
#pragma unroll
        for (int i = 0; i < S; ++i) {
            const int k = L * i + cj, p = (int)(__brev((unsigned)i) >> (32 - LOGS));
            int rec = -1;
            if (active && k < cols) rec = (MODE != 0 && a.in_map) ? a.in_map[k] : k;
            uint4 *d = D + (p * 2) * 32 + (lane ^ (p & (L - 1)));
            if (rec >= 0) {
                const uint4 *src = a.in + (b * a.in_sb + (long long)rec * a.in_sc) * 2;
                cp_async16(d, src);
                cp_async16(d + 32, src + 1);
            } else {
                d[0] = make_uint4(0, 0, 0, 0);
                d[32] = make_uint4(0, 0, 0, 0);
            }
        }
        cp_async_commit();
        if (MODE != 0 && a.path && active && j == 0) a.path[b] = 0;
        cp_async_wait_all();
        unsigned mask = mask0, rcs = 0;
        uint4 *outb = a.out + (active ? b : 0) * a.out_sb * 2;
        const int l0 = active ? lim0 : 0, l1 = active ? lim1 : 0x7fffffff;   // lanes beyond the batch store and check nothing

        if (MODE == 2) {   // weighting: x_k *= wt[k] (the erasure weights Zc(w^k); ids outside the examined set hold zero)
#pragma unroll
            for (int i = 0; i < S; ++i) {
                const int p = (int)(__brev((unsigned)i) >> (32 - LOGS));
                if ((mask >> p) & 1u) rcs |= bf_scale(slot(p, lane), dWt + (unsigned)((L * i + cj) * 32));
            }
        }
        // ---- P0: the S-point transform of this column.  Indices, twiddles and "is the twiddle trivial" are compile-time constants;
        // only the structural-zero pattern (how many coefficients the caller supplied) is a run-time, warp-uniform, test.
#pragma unroll
        for (int s = 0; s < LOGS; ++s) {
#pragma unroll
            for (int bf = 0; bf < S / 2; ++bf) {
                const int h = 1 << s, lo = bf & (h - 1), p = ((bf >> s) << (s + 1)) | lo;
                const int twi = (lo << (LOGS - 1 - s)) * L;
                const unsigned nu = (mask >> p) & 1u, nv = (mask >> (p + h)) & 1u;
                const unsigned chk = (MODE != 2 && s == 0) ? 1u : 0u;   // stage 0 sees every loaded value once: canonical-form check (MODE 2: at weighting)
                if (!(nu | nv)) continue;
                const unsigned pu = slot(p, lane), pv = slot(p + h, lane);
                if (LL == 0 && s == LOGS - 1) {   // one-pass transform (N = S): this is the last stage
                    const unsigned r = bf_emit<LOGN, MODE>(pu, pv, dTw + twi * 32, twi != 0, outb, a.out_sr, p, p + h, l0, l1);
                    if (active) rcs |= r;
                } else if (!nv) {
                    rcs |= bf_copy(pu, pv, chk);
                } else if (twi != 0) {
                    bf_mul(pu, pv, dTw + twi * 32);   // (never in stage 0: nothing to validate here)
                } else {
                    rcs |= bf_add(pu, pv, chk);
                }
            }
            // positions touched by this stage now hold (possibly) non-zero values
#pragma unroll
            for (int bf = 0; bf < S / 2; ++bf) {
                const int h = 1 << s, lo = bf & (h - 1), p = ((bf >> s) << (s + 1)) | lo;
                if (((mask >> p) | (mask >> (p + h))) & 1u) mask |= (1u << p) | (1u << (p + h));
            }
        }
        // ---- P1: this lane owns row q of its item; stage s2 pairs the item's columns bu and bu + 2^s2 (twiddle depends on the lane)
        if constexpr (LL > 0) {
            __syncwarp();
#pragma unroll
            for (int e = 0; e < ROWS; ++e) {
                const int q = HALVES > 1 ? (j & (S - 1)) : L * e + j;
                const int hf = HALVES > 1 ? (j >> LOGS) : 0;
#pragma unroll
                for (int s2 = 0; s2 < LL; ++s2) {
#pragma unroll
                    for (int bfi = 0; bfi < BFL; ++bfi) {
                        const int bf = hf * BFL + bfi;
                        const int hb2 = 1 << s2, lo = bf & (hb2 - 1), bu = ((bf >> s2) << (s2 + 1)) | lo;
                        const unsigned pu = slot(q, item * L + bu), pv = slot(q, item * L + bu + hb2);
                        // twiddle w_N^((pos mod h) * N/(2h)), h = S*2^s2: index (S*lo + q) << (LL-1-s2) of sTw = entry q of block 2^s2 - 1 + lo of sTw1
                        const unsigned ptw = dTw1 + (unsigned)((((hb2 - 1 + lo) << LOGS) + q) * 32);
                        if (s2 == LL - 1) {
                            const unsigned r = bf_emit<LOGN, MODE>(pu, pv, ptw, 1, outb, a.out_sr, S * bu + q, S * bu + q + S * hb2, l0, l1);
                            if (active) rcs |= r;
                        } else {
                            bf_mul(pu, pv, ptw);   // tw[0] is the Montgomery form of 1: the q = 0 row multiplies like the others
                        }
                    }
                    if (HALVES > 1 && s2 + 1 < LL) __syncwarp();   // the next stage pairs columns written by the row's other lane(s)
                }
            }
        }
        bad |= rcs & 1u;
        if (MODE != 0) {   // one lane per failing item raises the flag and appends the item to the list
            const unsigned fm = __ballot_sync(0xffffffffu, (rcs & 2u) && active);
            const unsigned grp = (L == 32 ? 0xffffffffu : ((1u << L) - 1u)) << (item * L);
            if ((fm & grp) && j == 0) mark_fail(a.fail, b, a.fail_list, a.fail_count);
        }
        __syncwarp();   // the rows are reloaded by other lanes' columns in the next tile
        tile = a.work ? nwarps + (long long)__shfl_sync(0xffffffffu, nxt, 0) : tile + nwarps;
    }
    if (a.work && lane == 0) {   // the last warp to leave resets the queue for the next launch on this stream
        if (atomicAdd(a.work + 1, 1ull) == (unsigned long long)(nwarps - 1)) {
            a.work[0] = 0ull;
            a.work[1] = 0ull;
        }
    }
    if (bad) *(volatile unsigned int *)a.err = 1u;
}

}  // namespace hb
