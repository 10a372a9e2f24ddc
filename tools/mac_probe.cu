// mac_probe.cu -- register-only microbenchmarks of candidate Fr multiply-accumulate inner loops on sm_100a.
//   A: radix 2^32, 8 limbs, even/odd IMAD.WIDE.U32.X carry chains (fr.cuh acc_mac)
//   B: radix 2^29, 9 limbs, 81 carry-free IMAD.WIDE.U32 into 17 64-bit columns, normalised every 7 terms
//   C: B without normalisation (upper bound)
//   D: single instruction rates: IMAD.WIDE.U32 (no carry) / .cc only / .X
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../mpc-protocols_b200/csrc/fr.cuh"
using namespace hb;

__global__ void __launch_bounds__(256) probeA(uint32_t *sink, uint32_t seed, int iters) {
    uint32_t a[8], b[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed * (i + 3) + threadIdx.x; b[i] = seed * (i + 11) ^ threadIdx.x; }
    __shared__ uint32_t sb[64][8];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sb[i >> 3][i & 7] = seed * (i + 17);
    __syncthreads();
    acc_t A; acc_zero(A);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { b[i] = sb[it & 63][i]; a[i] += b[(i + 1) & 7]; }
        acc_mac(A, a, b);
    }
    uint32_t r[8]; acc_reduce(A, r);
    uint32_t s = 0; for (int i = 0; i < 8; ++i) s ^= r[i];
    if (s == 0x1234567u) sink[0] = s;
}

template <int NORM>
__global__ void __launch_bounds__(256) probeB(uint32_t *sink, uint32_t seed, int iters) {
    uint32_t a[9], b[9];
    for (int i = 0; i < 9; ++i) { a[i] = (seed * (i + 3) + threadIdx.x) & 0x1fffffffu; b[i] = (seed * (i + 11) ^ threadIdx.x) & 0x1fffffffu; }
    __shared__ uint32_t sb[64][9];
    for (int i = threadIdx.x; i < 576; i += blockDim.x) sb[i / 9][i % 9] = (seed * (i + 17)) & 0x1fffffffu;
    __syncthreads();
    unsigned long long col[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) col[i] = 0;
#pragma unroll 1
    for (int it = 0; it < iters; it += 7) {
#pragma unroll
        for (int u = 0; u < 7; ++u) {
#pragma unroll
            for (int i = 0; i < 9; ++i) { b[i] = sb[(it + u) & 63][i]; a[i] = (a[i] + b[(i + 1) % 9]) & 0x1fffffffu; }
#pragma unroll
            for (int i = 0; i < 9; ++i)
#pragma unroll
                for (int j = 0; j < 9; ++j)
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(col[i + j]) : "r"(a[i]), "r"(b[j]));
        }
        if (NORM) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { col[i + 1] += col[i] >> 29; col[i] &= 0x1fffffffull; }
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 17; ++i) s ^= col[i];
    if (s == 0x1234567ull) sink[0] = (uint32_t)s;
}

// D: single-instruction-kind chains whose multiplicand is the accumulator's own low word (defeats hoisting)
template <int KIND>
__global__ void __launch_bounds__(256) probeD(uint32_t *sink, uint32_t seed, int iters) {
    uint32_t lo[8], hi[8], k = 0, m1 = seed * 77u + 5;
    for (int i = 0; i < 8; ++i) { lo[i] = threadIdx.x * 2654435761u + i; hi[i] = seed + i; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (KIND == 0) {  // 64-bit MAC, no carry in/out: IMAD.WIDE.U32 Rd, Ra, Rb, Rc
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("mad.lo.cc.u32 %0, %0, %2, %0;\n\tmadc.hi.u32 %1, %0, %2, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(m1));
            } else if (KIND == 1) {  // carry-out captured by addc
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint32_t a = lo[i];
                    asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;" : "+r"(lo[i]), "+r"(hi[i]), "+r"(k) : "r"(a), "r"(m1));
                }
            } else if (KIND == 2) {  // 4-lane carry chain (the product kernels' chain4)
                chain4(lo[0], hi[0], lo[1], hi[1], lo[2], hi[2], lo[3], hi[3], k, lo[4], lo[5], lo[6], lo[7], m1 + u);
                chain4(lo[4], hi[4], lo[5], hi[5], lo[6], hi[6], lo[7], hi[7], k, lo[0], lo[1], lo[2], lo[3], m1 - u);
            } else if (KIND == 3) {  // lo and hi as separate 32-bit IMADs
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint32_t a = lo[i];
                    asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[i]) : "r"(a), "r"(m1));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(a), "r"(m1));
                }
            } else {  // product-only wide multiply + 64-bit add done separately
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    unsigned long long p = (unsigned long long)lo[i] * m1;
                    unsigned long long x = ((unsigned long long)hi[i] << 32 | lo[i]) + p;
                    lo[i] = (uint32_t)x; hi[i] = (uint32_t)(x >> 32);
                }
            }
        }
    }
    uint32_t s = k;
    for (int i = 0; i < 8; ++i) s ^= lo[i] ^ hi[i];
    if (s == 0x1234567u) sink[0] = s;
}

__global__ void __launch_bounds__(256) probeF(double *sink, double seed, int iters) {
    double x[8], m = 1.0 + seed * 1e-9, c = seed * 1e-7;
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __fma_rz(x[i], m, c);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 1.2345) sink[0] = s;
}

template <typename F>
static float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    return best;
}

int main() {
    uint32_t *sink; cudaMalloc(&sink, 64);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 7 * 600;
    const double T = (double)blocks * threads;
    float ms;
    ms = timeit([&] { probeA<<<blocks, threads>>>(sink, 12345, iters); });
    printf("A radix32 chain MAC   : %.3f ms  %.1f GMAC/s  (%.2f T wide-mults/s)\n", ms, T * iters / ms / 1e6, T * iters * 64 / ms / 1e9);
    ms = timeit([&] { probeB<1><<<blocks, threads>>>(sink, 12345, iters); });
    printf("B radix29 MAC + norm  : %.3f ms  %.1f GMAC/s  (%.2f T wide-mults/s)\n", ms, T * iters / ms / 1e6, T * iters * 81 / ms / 1e9);
    ms = timeit([&] { probeB<0><<<blocks, threads>>>(sink, 12345, iters); });
    printf("C radix29 MAC no norm : %.3f ms  %.1f GMAC/s  (%.2f T wide-mults/s)\n", ms, T * iters / ms / 1e6, T * iters * 81 / ms / 1e9);
    const int it2 = 4096;
    ms = timeit([&] { probeD<0><<<blocks, threads>>>(sink, 12345, it2); });
    printf("D0 wide MAC no carry  : %.3f ms  %.2f T wide/s\n", ms, T * it2 * 64 / ms / 1e9);
    ms = timeit([&] { probeD<1><<<blocks, threads>>>(sink, 12345, it2); });
    printf("D1 wide MAC carry-out : %.3f ms  %.2f T wide/s\n", ms, T * it2 * 64 / ms / 1e9);
    ms = timeit([&] { probeD<2><<<blocks, threads>>>(sink, 12345, it2); });
    printf("D2 4-lane carry chain : %.3f ms  %.2f T wide/s\n", ms, T * it2 * 64 / ms / 1e9);
    ms = timeit([&] { probeD<3><<<blocks, threads>>>(sink, 12345, it2); });
    printf("D3 lo + hi separate   : %.3f ms  %.2f T (lo+hi pairs)/s\n", ms, T * it2 * 64 / ms / 1e9);
    ms = timeit([&] { probeD<4><<<blocks, threads>>>(sink, 12345, it2); });
    printf("D4 wide mul + add64   : %.3f ms  %.2f T wide/s\n", ms, T * it2 * 64 / ms / 1e9);
    ms = timeit([&] { probeF<<<blocks, threads>>>((double *)sink, 3.0, it2); });
    printf("F  DFMA chains        : %.3f ms  %.2f T dfma/s\n", ms, T * it2 * 64 / ms / 1e9);
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
