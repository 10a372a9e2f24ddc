// matvec.cuh -- the dense kernel: a small constant Fr matrix applied to a huge batch of Fr vectors.
//
//   out[b][r] = sum_c  M[r][c] * in[b][col_map[c]]   (mod r),   b < B, r < R, c < C (or c < row_len[r])
//
// It serves every dense site of the hot path that the NTT kernels (ntt.cuh) do not take:
//   K3/K4  optimistic check with flags          M = [check rows L_i(x_s); coefficient rows L_i[k]]
//          (robust_interpolate.rs:392-428: verify_matrix + basis_coeffs, identity rows dropped), also restricted to the
//          items an NTT check rejected (item_list) and as plain Lc*y for the items handed to the robust decoder
//   K3     triangular coefficient recovery P = Q*(N*Zc)^-1 mod x^(d+1) after the erasure-weighted inverse NTT
//   a10    NonRobustShare::recover_secret on id subsets: coefficient rows above `deg` are zero-check rows (shamir.rs:234-237)
//   K1/K2  Vandermonde with more columns than domain points, caller-supplied matrices (apply_vandermonde, share/mod.rs:50-76)
//
// Rows [0, n_chk) are "check rows": the result is compared with the supplied share in[b][chk_map[r]] (or with zero when
// chk_map[r] < 0) instead of being written; a mismatch in rows [0, n_gate) marks the item as failing (-> robust decode /
// DegreeMismatch), any mismatch sets the item's flag bit chk_map[r].
//
// Mapping: M lives in shared memory in Montgomery form (warp-uniform broadcast reads), a tile of TBT*32 batch items is
// staged in shared memory as [c][half][lane] uint4 (conflict-free 128-bit reads); work items (row, 32-lane sub-tile)
// are dealt round-robin to the CTA's warps; one thread accumulates one output lazily (fr.cuh) and reduces once.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fr.cuh"

namespace hb {

struct MatvecArgs {
    const uint4 *M;   // [R][C][2] uint4 (Montgomery form)
    const uint4 *in;  // element (b, j) at in[(b*in_sb + j*in_sc)*2 .. +1]
    uint4 *out;       // element (b, row) at out[(b*out_sb + row*out_sr)*2 .. +1]
    int R, C;
    long long B;
    long long in_sb, in_sc, out_sb, out_sr;  // strides in 32-byte elements
    const int *col_map;                     // [C] input index of column c, or nullptr (identity)
    int n_chk, n_gate;
    const int *chk_map;        // [n_chk] input index checked by check row r (< 0: compare with zero)
    unsigned char *fail;       // [B] set to 1 when a gate row mismatches (may be nullptr when n_gate == 0) ...
    unsigned int *fail_list, *fail_count;   // ... and the item is appended to fail_list[(*fail_count)++] (optional)
    unsigned long long *flags; // [B][flag_words] or nullptr
    int flag_words;
    unsigned int *err;         // device word: bit0 set when a non-canonical (>= r) input is seen
    int rows_per_slice;        // rows handled by one blockIdx.y slice
    int in_chunk_major;        // 1: in_sc == 1 (tile is contiguous), 0: in_sb == 1 (sender-major)
    const unsigned int *item_list;   // optional indirection: process items item_list[0 .. *item_count) instead of 0 .. B
    const unsigned int *item_count;
    const int *row_len;        // optional [R]: row r uses only row_len[r] columns (triangular matrices) ...
    const int *row_start;      // ... starting at column row_start[r] (optional, with row_len; nullptr: column 0)
    int M_ld;                  // leading dimension of M in elements (0: C) -- column blocks of a wider matrix
    int col0;                  // first column of the block: column c reads input col_map[col0 + c] (or col0 + c)
};

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(uint4 *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int TBT>
__global__ void __launch_bounds__(512) matvec_kernel(const MatvecArgs a) {
    fma_ballast(a.B < 0, a.err);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int C = a.C;
    const int r0 = blockIdx.y * a.rows_per_slice;
    const int nrows = min(a.rows_per_slice, a.R - r0);
    uint4 *sM = reinterpret_cast<uint4 *>(smem_raw);                 // [nrows][C][2]
    uint4 *sD = sM + (size_t)a.rows_per_slice * C * 2;               // [TBT][C][2][32]
    unsigned int *sFail = reinterpret_cast<unsigned int *>(sD + (size_t)TBT * C * 64);  // [TBT*32]
    unsigned long long *sFlags = reinterpret_cast<unsigned long long *>(sFail + TBT * 32);  // [TBT*32][flag_words]

    // constant matrix slice -> shared memory (once per CTA; the CTA is persistent over batch tiles)
    {
        const int ld = a.M_ld ? a.M_ld : C;
        for (int i = tid; i < nrows * C * 2; i += blockDim.x) sM[i] = a.M[((size_t)(r0 + i / (C * 2)) * ld) * 2 + i % (C * 2)];
    }

    // staging (below): this thread's first element of a chunk-major tile and the step of one pass of the CTA, as (chunk, 2*column + half)
    constexpr int STAGE_U = 8;
    const int stage_bl0 = tid / (2 * C), stage_c20 = tid - stage_bl0 * (2 * C);
    const int stage_qd = (int)blockDim.x / (2 * C), stage_rd = (int)blockDim.x - stage_qd * (2 * C);

    const long long TILE = TBT * 32;
    const long long NB = a.item_list ? (long long)*a.item_count : a.B;  // items to process
    const long long ntiles = (NB + TILE - 1) / TILE;
    const int nitems = nrows * TBT;
    const bool checks_here = r0 < a.n_chk;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long b0 = tile * TILE;
        __syncthreads();  // previous tile fully consumed (and sM visible on the first pass)
        // ---- stage the data tile: element (b, c) -> sD[sub][c][half][lane]
        // STAGE_U loads are issued before the first of their stores (a load followed at once by its dependent store exposes one
        // DRAM round trip per element: with few rows per tile -- the triangular recovery has 132 terms per chunk -- staging was a
        // third of the kernel's time); chunk-major tiles advance (chunk, column) incrementally instead of dividing by C per element
        const int nhalves = TBT * 32 * C * 2;
        int bl_i = stage_bl0, c2_i = stage_c20;
        for (int i0 = tid; i0 < nhalves; i0 += (int)blockDim.x * STAGE_U) {
            uint4 v[STAGE_U];
            int dst[STAGE_U];
#pragma unroll
            for (int u = 0; u < STAGE_U; ++u) {
                const int i = i0 + u * (int)blockDim.x;
                v[u] = make_uint4(0, 0, 0, 0);
                dst[u] = -1;
                if (i < nhalves) {
                    int half, bl, c;
                    if (a.in_chunk_major) {   // i = (bl*C + c)*2 + half
                        half = c2_i & 1; c = c2_i >> 1; bl = bl_i;
                        c2_i += stage_rd; bl_i += stage_qd;
                        if (c2_i >= 2 * C) { c2_i -= 2 * C; ++bl_i; }
                    } else { half = i & 1; bl = (i >> 1) % (TBT * 32); c = (i >> 1) / (TBT * 32); }
                    long long b = b0 + bl;
                    if (b < NB) {
                        if (a.item_list) b = a.item_list[b];
                        int j = a.col_map ? a.col_map[a.col0 + c] : a.col0 + c;
                        v[u] = ldg_stream(a.in + (b * a.in_sb + (long long)j * a.in_sc) * 2 + half);
                    }
                    dst[u] = (((bl >> 5) * C + c) * 2 + half) * 32 + (bl & 31);
                }
            }
#pragma unroll
            for (int u = 0; u < STAGE_U; ++u)
                if (dst[u] >= 0) sD[dst[u]] = v[u];
        }
        if (checks_here) {
            for (int i = tid; i < TBT * 32; i += blockDim.x) sFail[i] = 0;
            if (a.flags) for (int i = tid; i < TBT * 32 * a.flag_words; i += blockDim.x) sFlags[i] = 0ull;
        }
        __syncthreads();

        for (int item = warp; item < nitems; item += W) {
            const int rl = item / TBT, sub = item - rl * TBT;
            const int r = r0 + rl;
            long long b = b0 + sub * 32 + lane;
            const bool live = b < NB;
            if (live && a.item_list) b = a.item_list[b];
            const uint4 *Mrow = sM + (size_t)rl * C * 2;
            const uint4 *Dsub = sD + (size_t)sub * C * 64 + lane;   // (both advanced to the row's first column below)
            const bool is_chk = r < a.n_chk;
            // the supplied share a check row is compared with (prefetched; hidden behind the dot product)
            uint4 y_lo = make_uint4(0, 0, 0, 0), y_hi = y_lo;
            int chk_j = 0;
            if (is_chk) {
                chk_j = a.chk_map[r];  // < 0: the row must evaluate to zero (degree check)
                if (chk_j >= 0 && live) {
                    const uint4 *p = a.in + (b * a.in_sb + (long long)chk_j * a.in_sc) * 2;
                    y_lo = ldg_stream(p);
                    y_hi = ldg_stream(p + 1);
                }
            }
            acc_t A;
            acc_zero(A);
            unsigned bad = 0;
            const int Cr = a.row_len ? a.row_len[r] : C;  // terms of this row
            if (a.row_len && a.row_start) {                // ... starting at column row_start[r]
                const int cs = a.row_start[r];
                Mrow += cs * 2;
                Dsub += cs * 64;
            }
            const bool validate = (rl == 0);  // one row per slice validates the tile's inputs (each input exactly once per slice)
            // two register sets, loads issued one term ahead, no register moves (loop unrolled by two)
            uint4 a0 = Dsub[0], a1 = Dsub[32], b0 = Mrow[0], b1 = Mrow[1];
            uint4 c0 = a0, c1 = a1, d0 = b0, d1 = b1;
#pragma unroll 1
            for (int c = 0; c < Cr; c += 2) {
                if (c + 1 < Cr) {
                    c0 = Dsub[(c + 1) * 64];
                    c1 = Dsub[(c + 1) * 64 + 32];
                    d0 = Mrow[(c + 1) * 2];
                    d1 = Mrow[(c + 1) * 2 + 1];
                }
                {
                    const uint32_t x[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const uint32_t m[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    if (validate) bad |= geq_mod(x) ? 1u : 0u;
                    acc_mac(A, x, m);
                }
                if (c + 2 < Cr) {
                    a0 = Dsub[(c + 2) * 64];
                    a1 = Dsub[(c + 2) * 64 + 32];
                    b0 = Mrow[(c + 2) * 2];
                    b1 = Mrow[(c + 2) * 2 + 1];
                }
                if (c + 1 < Cr) {
                    const uint32_t x[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                    const uint32_t m[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                    if (validate) bad |= geq_mod(x) ? 1u : 0u;
                    acc_mac(A, x, m);
                }
            }
            uint32_t res[8];
            acc_reduce(A, res);
            if (is_chk) {
                uint32_t y[8];
                load_fr(y, y_lo, y_hi);
                bad |= geq_mod(y) ? 1u : 0u;
                if (live && !fr_eq(res, y)) {
                    if (r < a.n_gate) sFail[sub * 32 + lane] = 1u;
                    if (a.flags && chk_j >= 0) atomicOr(&sFlags[(size_t)(sub * 32 + lane) * a.flag_words + (chk_j >> 6)], 1ull << (chk_j & 63));
                }
            } else if (live) {
                uint4 *o = a.out + (b * a.out_sb + (long long)(r - a.n_chk) * a.out_sr) * 2;
                stg_stream(o, make_uint4(res[0], res[1], res[2], res[3]));
                stg_stream(o + 1, make_uint4(res[4], res[5], res[6], res[7]));
            }
            if (bad) *(volatile unsigned int *)a.err = 1u;  // mapped host memory: plain store, every writer stores 1
        }
        if (checks_here) {
            __syncthreads();
            for (int i = tid; i < TBT * 32; i += blockDim.x) {
                long long b = b0 + i;
                if (b < NB) {
                    if (a.item_list) b = a.item_list[b];
                    if (a.fail && sFail[i]) mark_fail(a.fail, b, a.fail_list, a.fail_count);  // slices only ever raise the flag (buffer pre-zeroed by the host side)
                    if (a.flags)
                        for (int w = 0; w < a.flag_words; ++w) {
                            unsigned long long f = sFlags[(size_t)i * a.flag_words + w];
                            if (f) atomicOr(&a.flags[b * a.flag_words + w], f);
                        }
                }
            }
        }
    }
}

// Column-split fallback for matrices too wide for one shared-memory tile (C > ~200): the two column blocks are applied by
// two plain launches into T1 / T2 [B][R]; this kernel adds them and applies the row semantics (check rows / outputs).
__global__ void matvec_combine_kernel(const MatvecArgs a, const uint4 *T1, const uint4 *T2) {
    const long long NB = a.item_list ? (long long)*a.item_count : a.B;
    const long long total = NB * a.R;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long b = i / a.R;
        const int r = (int)(i - b * a.R);
        if (a.item_list) b = a.item_list[b];
        uint32_t x[8], y[8], v[8];
        load_fr(x, T1[(b * a.R + r) * 2], T1[(b * a.R + r) * 2 + 1]);
        load_fr(y, T2[(b * a.R + r) * 2], T2[(b * a.R + r) * 2 + 1]);
        fr_add(v, x, y);
        if (r < a.n_chk) {
            const int chk_j = a.chk_map[r];
            uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (chk_j >= 0) {
                const uint4 *p = a.in + (b * a.in_sb + (long long)chk_j * a.in_sc) * 2;
                load_fr(w, ldg_stream(p), ldg_stream(p + 1));
            }
            if (!fr_eq(v, w)) {
                if (r < a.n_gate && a.fail) mark_fail(a.fail, b, a.fail_list, a.fail_count);
                if (a.flags && chk_j >= 0) atomicOr(&a.flags[b * a.flag_words + (chk_j >> 6)], 1ull << (chk_j & 63));
            }
        } else {
            uint4 *o = a.out + (b * a.out_sb + (long long)(r - a.n_chk) * a.out_sr) * 2;
            stg_stream(o, make_uint4(v[0], v[1], v[2], v[3]));
            stg_stream(o + 1, make_uint4(v[4], v[5], v[6], v[7]));
        }
    }
}

// ---- launch-shape chooser ---------------------------------------------------------------------
struct MatvecPlan {
    int tbt;             // 32-lane sub-tiles per CTA tile (1, 2 or 4)
    int warps;           // warps per CTA
    int rows_per_slice;  // rows of M resident per CTA
    int slices;          // gridDim.y
    int ctas_per_sm;
    size_t smem;
};

inline size_t matvec_smem_bytes(int rows, int C, int tbt, int flag_words) {
    return (size_t)rows * C * 32 + (size_t)tbt * C * 1024 + (size_t)tbt * 32 * 4 + (size_t)tbt * 32 * 8 * (flag_words > 0 ? flag_words : 0) + 16;
}

// Pick the shape with the most resident warps per SM (register file allows ~24 warps at <=84 regs) and the
// best work balance (items per warp integral), keeping all of M resident when it fits in 227 KB.
inline MatvecPlan matvec_plan(int R, int C, int flag_words, int regs_per_thread, long long B = -1, int num_sms = 148) {
    const size_t SMEM_CTA = 227 * 1024, SMEM_SM = 228 * 1024;
    MatvecPlan best{};
    double best_score = -1.0;
    const int max_warps_regs = (65536 / (regs_per_thread > 0 ? regs_per_thread : 96)) / 32;
    for (int slices = 1; slices <= R; ++slices) {
        int rps = (R + slices - 1) / slices;
        if ((rps * (slices - 1)) >= R && slices > 1) continue;  // empty last slice
        bool any = false;
        for (int tbt : {4, 2, 1}) {
            // session-sized batches: prefer tiles small enough to give every SM one (latency, not throughput, matters there)
            if (B > 0 && tbt > 1 && (B + tbt * 32 - 1) / (tbt * 32) < num_sms) continue;
            for (int warps : {16, 8, 4}) {
                size_t smem = matvec_smem_bytes(rps, C, tbt, flag_words);
                if (smem > SMEM_CTA) continue;
                int ctas = (int)(SMEM_SM / (smem + 1024));
                if (ctas < 1) continue;
                int by_regs = max_warps_regs / warps;
                if (by_regs < 1) continue;
                if (ctas > by_regs) ctas = by_regs;
                if (ctas > 32) ctas = 32;
                int resident = ctas * warps;
                int items = rps * tbt;
                int per_warp = (items + warps - 1) / warps;
                double balance = (double)items / ((double)per_warp * warps);
                double occ = resident >= 16 ? 1.0 : (double)resident / 16.0;
                // staging overhead relative to MAC work shrinks with rows per slice; prefer fewer slices strongly
                double score = balance * occ - 0.02 * (slices - 1) + 0.001 * resident;
                if (score > best_score) { best_score = score; best = MatvecPlan{tbt, warps, rps, slices, ctas, smem}; }
                any = true;
            }
        }
        if (any && slices >= 1 && best_score > 0.85) break;  // good enough without further slicing
    }
    return best;
}

}  // namespace hb
