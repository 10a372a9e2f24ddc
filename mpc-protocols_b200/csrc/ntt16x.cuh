// ntt16x.cuh -- the domain transforms of K1 / K2 / K3 for N = 16 .. 128 with ONE multiplication site and no CTA barrier.
//
// Same mathematics as ntt.cuh (radix-2 decimation in time over the reference's evaluation domain, robust_interpolate.rs:68-79,
// common/share/mod.rs:31-45; data canonical, twiddles in Montgomery form), different mapping.  N = 16*L.  L lanes of a warp own
// one item (32/L items per warp); the item's 16 x L working set lives in the warp's 16 KB of shared memory as 16 rows x 32
// lane-columns of 32-byte elements:
//   P0  lane j computes, alone and in its own column, the 16-point transform of the inputs k = L*i + bitrev(j) (4 stages, 32
//       butterflies).  Every lane of the warp runs the same butterfly at the same time, so twiddle indices -- and which operands
//       are structurally zero (share generation feeds d+1 coefficients into N points) -- are warp-uniform: trivial twiddles and
//       zero operands cost nothing, for every lane, without any divergence.  No synchronisation: the column is private.
//   P1  lane j then owns the rows q = L*e + j of its item and runs the remaining log2(L) stages in place along each row (the L
//       columns of the item), emitting natural-order outputs from the last stage.  One __syncwarp separates P0 and P1.
// The butterflies of P0 and P1 (and the weighting of MODE 2) are iterations of one loop around a single inlined Montgomery
// product: the kernel is ~1.5k instructions (24 KB, inside the 32 KB L1.5 instruction cache) where the fully unrolled kernels of
// ntt.cuh are 55 KB, and its warps never wait for each other (ntt64_cta_kernel: three CTA barriers per tile).
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

#ifndef HB_NTT16X_MINB
#define HB_NTT16X_MINB 3
#endif
constexpr int NTT16X_WARPS = 4;

template <int LOGN, int MODE>
inline size_t ntt16x_smem_bytes() {
    constexpr int N = 1 << LOGN;
    return (size_t)((N / 2) * 2 + (MODE == 2 ? N * 2 : 0) + NTT16X_WARPS * 16 * 2 * 32) * 16 + 16;
}

template <int LOGN, int MODE>
__global__ void __launch_bounds__(NTT16X_WARPS * 32, HB_NTT16X_MINB) ntt16x_kernel(const NttArgs a) {
    static_assert(LOGN >= 4 && LOGN <= 7, "N = 16 .. 128");
    static_assert(MODE >= 0 && MODE <= 2, "forward, inverse + degree check, weighted inverse");
    constexpr int N = 1 << LOGN, LL = LOGN - 4, L = 1 << LL, IPW = 32 / L;
    constexpr int P1_PER_ROW = (L / 2) * LL, ROWS = 16 / L;          // P1: butterflies per row, rows per lane
    constexpr int OPS_SCALE = (MODE == 2) ? 16 : 0, OPS_P0 = 32, OPS_P1 = ROWS * P1_PER_ROW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *sTw = reinterpret_cast<uint4 *>(smem_raw);                 // [N/2][2]  w^k (w^-k for the inverse), Montgomery form
    uint4 *sWt = sTw + N;                                             // [N][2]    MODE 2: weights by domain index
    uint4 *sD = sWt + (MODE == 2 ? 2 * N : 0);                        // [warps][16 rows][2 halves][32 columns]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < N; i += blockDim.x) sTw[i] = a.tw[i];
    if (MODE == 2)
        for (int i = t; i < 2 * N; i += blockDim.x) sWt[i] = a.wt[i];
    __syncthreads();
    uint4 *D = sD + (size_t)warp * (16 * 2 * 32);
    const int j = lane & (L - 1), item = lane >> LL;
    const int cj = (LL == 0) ? 0 : (int)(__brev((unsigned)j) >> (32 - (LL == 0 ? 1 : LL)));   // residue class of this lane's inputs
    const int cols = a.cols < N ? a.cols : N;
    const int cnt = min(16, (cols + L - 1) / L);                      // inputs per column that may be non-zero (warp-uniform)
    unsigned mask0 = 0;                                               // rows that hold a (possibly) non-zero value after loading
    for (int i = 0; i < cnt; ++i) mask0 |= 1u << (__brev((unsigned)i) >> 28);
    // element (row p, half hf) of column c of this warp's tile
    auto slot = [&](int p, int c) -> uint4 * { return D + (p * 2) * 32 + (c ^ (p & (L - 1))); };

    const long long ntiles = (a.B + IPW - 1) / IPW;
    unsigned bad = 0;
    for (long long tile = (long long)blockIdx.x * NTT16X_WARPS + warp; tile < ntiles; tile += (long long)gridDim.x * NTT16X_WARPS) {
        const long long b = tile * IPW + item;
        const bool active = b < a.B;
        // ---- inputs: lane j takes k = L*i + cj to row bitrev4(i) of its own column (asynchronous copies, no registers)
        for (int i = 0; i < cnt; ++i) {
            const int k = L * i + cj, p = (int)(__brev((unsigned)i) >> 28);
            int rec = -1;
            if (active && k < cols) rec = (MODE != 0 && a.in_map) ? a.in_map[k] : k;
            uint4 *d = slot(p, lane);
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
        unsigned mask = mask0;
        unsigned failed = 0;

#pragma unroll 1
        for (int op = -OPS_SCALE; op < OPS_P0 + OPS_P1; ++op) {
            uint4 *pu, *pv;               // operands (half 0; half 1 at +32)
            const uint4 *pw;              // multiplier
            bool mul, has_u = true, copy = false, scale = false, last, check0 = false;
            int posU = 0, posV = 0;       // natural-order output positions (last stage only)
            if (MODE == 2 && op < 0) {
                // weighting: x_k *= wt[k] (the erasure weights Zc(w^k); ids outside the examined set hold zero)
                const int i = op + OPS_SCALE, p = (int)(__brev((unsigned)i) >> 28);
                if (!((mask >> p) & 1u)) continue;
                pv = slot(p, lane);
                pu = pv;
                pw = sWt + (L * i + cj) * 2;
                mul = true; scale = true; last = false;
            } else if (op < OPS_P0) {
                // P0: stage s of the 16-point transform of this column; everything below is warp-uniform
                const int s = op >> 3, bf = op & 7, h = 1 << s, lo = bf & (h - 1);
                const int p = ((bf >> s) << (s + 1)) | lo;
                const unsigned nu = (mask >> p) & 1u, nv = (mask >> (p + h)) & 1u;
                if (!(nu | nv)) continue;
                mask |= (1u << p) | (1u << (p + h));
                pu = slot(p, lane);
                pv = slot(p + h, lane);
                has_u = nu != 0;
                copy = nv == 0;
                const int twi = (lo << (3 - s)) * L;
                pw = sTw + twi * 2;
                mul = twi != 0;
                last = (LL == 0) && s == 3;
                posU = p; posV = p + h;
                check0 = (MODE != 2) && s == 0;        // stage 0 sees every loaded value exactly once: canonical-form check (MODE 2: at weighting)
            } else if constexpr (LL > 0) {
                // P1: this lane owns row q of its item; stage s2 pairs the item's columns bu and bu + 2^s2
                if (op == OPS_P0) __syncwarp();
                const int o = op - OPS_P0, e = o / P1_PER_ROW, r1 = o - e * P1_PER_ROW;
                const int s2 = r1 / (L / 2), bf = r1 - s2 * (L / 2), hb2 = 1 << s2, lo = bf & (hb2 - 1);
                const int bu = ((bf >> s2) << (s2 + 1)) | lo, q = L * e + j;
                pu = slot(q, item * L + bu);
                pv = slot(q, item * L + bu + hb2);
                const int twi = (16 * lo + q) << (LL - 1 - s2);   // w_N^((pos mod h) * N/(2h)), h = 16*2^s2
                pw = sTw + twi * 2;
                mul = twi != 0;
                last = s2 == LL - 1;
                posU = 16 * bu + q; posV = posU + 16 * hb2;
            } else {
                continue;
            }
            uint32_t u[8], v[8], tt[8];
            if (last) {
                // which outputs are wanted: MODE 0: j < n;  MODE 1/2: stored below mout, checked (must vanish) from m on
                bool needU, needV;
                if (MODE == 0) { needU = posU < a.n; needV = posV < a.n; }
                else { needU = posU < a.mout || posU >= a.m; needV = posV < a.mout || posV >= a.m; }
                if (!needU && !needV) continue;
                if (MODE != 0 && posU >= a.m) {   // both must vanish <=> u == 0 and v == 0: no product
                    load_fr(u, pu[0], pu[32]);
                    load_fr(v, pv[0], pv[32]);
                    if (!fr_is_zero(u) || !fr_is_zero(v)) failed = 1;
                    continue;
                }
            }
            if (copy && !last) {   // v is structurally zero: (u, v) <- (u, u)
                load_fr(u, pu[0], pu[32]);
                if (check0) bad |= geq_mod(u) ? 1u : 0u;
                pv[0] = make_uint4(u[0], u[1], u[2], u[3]);
                pv[32] = make_uint4(u[4], u[5], u[6], u[7]);
                continue;
            }
            if (copy) {            // (last stage of a one-pass transform: fall through and emit u twice)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0;
                mul = false;
            } else {
                load_fr(v, pv[0], pv[32]);
                if (check0 || scale) bad |= geq_mod(v) ? 1u : 0u;
            }
            if (mul) {
                uint32_t w[8];
                load_fr(w, pw[0], pw[1]);
                mont_mul(tt, v, w);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) tt[i] = v[i];
            }
            if (scale) {
                pv[0] = make_uint4(tt[0], tt[1], tt[2], tt[3]);
                pv[32] = make_uint4(tt[4], tt[5], tt[6], tt[7]);
                continue;
            }
            if (has_u) {
                load_fr(u, pu[0], pu[32]);
                if (check0) bad |= geq_mod(u) ? 1u : 0u;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) u[i] = 0;
            }
            uint32_t sm[8], df[8];
            fr_add(sm, u, tt);
            fr_sub(df, u, tt);
            if (!last) {
                pu[0] = make_uint4(sm[0], sm[1], sm[2], sm[3]);
                pu[32] = make_uint4(sm[4], sm[5], sm[6], sm[7]);
                pv[0] = make_uint4(df[0], df[1], df[2], df[3]);
                pv[32] = make_uint4(df[4], df[5], df[6], df[7]);
            } else if (active) {
#pragma unroll 1
                for (int z = 0; z < 2; ++z) {
                    const int pos = z ? posV : posU;
                    uint32_t val[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) val[i] = z ? df[i] : sm[i];
                    if (MODE == 0) {
                        if (pos < a.n) {
                            uint4 *o = a.out + (b * a.out_sb + (long long)pos * a.out_sr) * 2;
                            stg_stream(o, make_uint4(val[0], val[1], val[2], val[3]));
                            stg_stream(o + 1, make_uint4(val[4], val[5], val[6], val[7]));
                        }
                    } else if (pos < a.mout) {
                        uint32_t c[8];
                        if (MODE == 1) fr_div_pow2<LOGN>(c, val);
                        else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) c[i] = val[i];
                        }
                        uint4 *o = a.out + (b * a.out_sb + (long long)pos * a.out_sr) * 2;
                        stg_stream(o, make_uint4(c[0], c[1], c[2], c[3]));
                        stg_stream(o + 1, make_uint4(c[4], c[5], c[6], c[7]));
                    } else if (pos >= a.m) {
                        if (!fr_is_zero(val)) failed = 1;
                    }
                }
            }
        }
        if (MODE != 0 && failed && active) a.fail[b] = 1;
        __syncwarp();   // the rows are reloaded by other lanes' columns in the next tile
    }
    if (bad) *(volatile unsigned int *)a.err = 1u;
}

}  // namespace hb
