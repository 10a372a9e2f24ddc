// sampler.cuh -- the randomness of a reference sharing, drawn on the device.
//
// The reference draws the polynomial INSIDE compute_shares (robust_interpolate.rs:68-69: DensePolynomial::rand(degree, rng), coefficient
// 0 overwritten by the secret; share_gen.rs:250 draws the secret with F::rand first) from rand 0.8's StdRng.  Neither generator nor
// sampler is in the reference tree (rand_chacha 0.3, ark-ff 0.5: crates.io dependencies); their published algorithms are restated here
// (the CPU restatement that the tests check this file against is tests-side: chacha_fr.py):
//   StdRng = ChaCha12: key = 32-byte seed, 64-bit block counter (state words 12-13) from 0, stream id 0; next_u64 = two consecutive
//   output words, low word first.   Fp::rand = four next_u64 limbs, top bit of the last limb cleared, redrawn while >= r; the accepted
//   limbs are the Montgomery representation, so the value is limbs * 2^-256 mod r (one Montgomery product with 1).
// Candidate c of the stream is words 8c .. 8c+7 (half of block c/2).  Which candidates are accepted is data dependent, so the rank of
// a candidate among the accepted ones is a prefix sum: count per CTA -> scan -> emit (ChaCha blocks are recomputed, not stored).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fr.cuh"

namespace hb {

struct SampleArgs {
    uint32_t key[8];
    unsigned long long ncand;      // candidates examined: 0 .. ncand-1
    unsigned int *counts;          // [nblk] accepted candidates per CTA; after the scan: exclusive prefix
    unsigned long long *total;     // accepted candidates among the ncand
    uint4 *out;                    // canonical elements
    unsigned long long want;       // accepted elements 0 .. want-1 are written
    int mode, per;                 // 0: out[rank]; 1: sharing b = rank / (per+2): draw 0 -> coeffs[b][0], draw 1 dropped, draw j -> coeffs[b][j-1];
                                   // 2: sharing b = rank / (per+1): draw 0 dropped (the caller supplies the secret), draw j -> coeffs[b][j]   (per = degree)
};

#define HB_CC_QR(a, b, c, d)                                  \
    a += b; d ^= a; d = __funnelshift_l(d, d, 16);            \
    c += d; b ^= c; b = __funnelshift_l(b, b, 12);            \
    a += b; d ^= a; d = __funnelshift_l(d, d, 8);             \
    c += d; b ^= c; b = __funnelshift_l(b, b, 7);

__device__ __forceinline__ void chacha12_block(const uint32_t (&key)[8], unsigned long long counter, uint32_t (&o)[16]) {
    const uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                             (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t x0 = in[0], x1 = in[1], x2 = in[2], x3 = in[3], x4 = in[4], x5 = in[5], x6 = in[6], x7 = in[7], x8 = in[8], x9 = in[9], x10 = in[10],
             x11 = in[11], x12 = in[12], x13 = in[13], x14 = in[14], x15 = in[15];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        HB_CC_QR(x0, x4, x8, x12) HB_CC_QR(x1, x5, x9, x13) HB_CC_QR(x2, x6, x10, x14) HB_CC_QR(x3, x7, x11, x15)
        HB_CC_QR(x0, x5, x10, x15) HB_CC_QR(x1, x6, x11, x12) HB_CC_QR(x2, x7, x8, x13) HB_CC_QR(x3, x4, x9, x14)
    }
    o[0] = x0 + in[0]; o[1] = x1 + in[1]; o[2] = x2 + in[2]; o[3] = x3 + in[3]; o[4] = x4 + in[4]; o[5] = x5 + in[5]; o[6] = x6 + in[6];
    o[7] = x7 + in[7]; o[8] = x8 + in[8]; o[9] = x9 + in[9]; o[10] = x10 + in[10]; o[11] = x11 + in[11]; o[12] = x12 + in[12];
    o[13] = x13 + in[13]; o[14] = x14 + in[14]; o[15] = x15 + in[15];
}

// candidate c -> limbs (top bit cleared) and whether Fp::rand accepts them
__device__ __forceinline__ bool sample_candidate(const uint32_t (&key)[8], unsigned long long c, uint32_t (&limbs)[8]) {
    uint32_t o[16];
    chacha12_block(key, c >> 1, o);
    const int h = (int)(c & 1ull) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) limbs[i] = (c & 1ull) ? o[8 + i] : o[i];
    (void)h;
    limbs[7] &= 0x7fffffffu;
    return !geq_mod(limbs);
}

constexpr int SAMPLE_THREADS = 256;

__global__ void __launch_bounds__(SAMPLE_THREADS) sample_count_kernel(const SampleArgs a) {
    const unsigned long long c = (unsigned long long)blockIdx.x * SAMPLE_THREADS + threadIdx.x;
    uint32_t limbs[8];
    const bool acc = c < a.ncand && sample_candidate(a.key, c, limbs);
    const int n = __syncthreads_count(acc ? 1 : 0);
    if (threadIdx.x == 0) a.counts[blockIdx.x] = (unsigned int)n;
}
// exclusive scan of the per-CTA counts (one CTA; the counts of 10^8 candidates are 4*10^5 words)
__global__ void __launch_bounds__(1024) sample_scan_kernel(unsigned int *counts, unsigned int nblk, unsigned long long *total) {
    __shared__ unsigned long long part[1024];
    const unsigned int per = (nblk + 1023) / 1024, lo = threadIdx.x * per, hi = min(nblk, lo + per);
    unsigned long long s = 0;
    for (unsigned int i = lo; i < hi; ++i) s += counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; ++i) { const unsigned long long v = part[i]; part[i] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (unsigned int i = lo; i < hi; ++i) { const unsigned int v = counts[i]; counts[i] = (unsigned int)run; run += v; }   // (prefixes fit 32 bits: < 2^32 candidates per call)
}
__global__ void __launch_bounds__(SAMPLE_THREADS) sample_emit_kernel(const SampleArgs a) {
    __shared__ unsigned int warp_base[SAMPLE_THREADS / 32];
    const unsigned long long c = (unsigned long long)blockIdx.x * SAMPLE_THREADS + threadIdx.x;
    uint32_t limbs[8];
    const bool acc = c < a.ncand && sample_candidate(a.key, c, limbs);
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int bal = __ballot_sync(0xffffffffu, acc);
    if (lane == 0) warp_base[warp] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int run = 0;
        for (int w = 0; w < SAMPLE_THREADS / 32; ++w) { const unsigned int v = warp_base[w]; warp_base[w] = run; run += v; }
    }
    __syncthreads();
    if (!acc) return;
    const unsigned long long rank = (unsigned long long)a.counts[blockIdx.x] + warp_base[warp] + __popc(bal & ((1u << lane) - 1u));
    if (rank >= a.want) return;
    unsigned long long dst;
    if (a.mode == 0) dst = rank;
    else if (a.mode == 1) {
        const unsigned long long b = rank / (unsigned long long)(a.per + 2);
        const int j = (int)(rank - b * (unsigned long long)(a.per + 2));
        if (j == 1) return;   // the coefficient DensePolynomial::rand drew for position 0 is overwritten by the secret
        dst = b * (unsigned long long)(a.per + 1) + (j == 0 ? 0 : j - 1);
    } else {
        const unsigned long long b = rank / (unsigned long long)(a.per + 1);
        const int j = (int)(rank - b * (unsigned long long)(a.per + 1));
        if (j == 0) return;
        dst = b * (unsigned long long)(a.per + 1) + j;
    }
    uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0}, v[8];
    mont_mul(v, limbs, one);   // Montgomery representation -> canonical value
    a.out[dst * 2] = make_uint4(v[0], v[1], v[2], v[3]);
    a.out[dst * 2 + 1] = make_uint4(v[4], v[5], v[6], v[7]);
}

}  // namespace hb
