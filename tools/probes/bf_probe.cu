// bf_probe.cu -- where does an NTT butterfly lose multiplier-pipe time?  Register-only / shared-memory / out-of-line variants of
// (u, v) <- (u + w*v, u - w*v) over fr.cuh's mont_mul, timed at a given number of resident warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mpc-protocols_b200/csrc -o build/bf_probe tools/probes/bf_probe.cu
//   build/bf_probe [warps_per_smsp=6] [iters=2000]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "fr.cuh"
using namespace hb;

__device__ __forceinline__ void lds8(uint32_t (&x)[8], const uint4 *p) { uint4 a = p[0], b = p[32]; load_fr(x, a, b); }
__device__ __forceinline__ void sts8(uint4 *p, const uint32_t (&x)[8]) { p[0] = make_uint4(x[0], x[1], x[2], x[3]); p[32] = make_uint4(x[4], x[5], x[6], x[7]); }

__device__ __noinline__ void bf_call(uint4 *pu, uint4 *pv, const uint4 *pw) {
    uint32_t u[8], v[8], w[8], t[8], s[8], d[8];
    lds8(v, pv); lds8(w, pw); lds8(u, pu);
    mont_mul(t, v, w);
    fr_add(s, u, t); fr_sub(d, u, t);
    sts8(pu, s); sts8(pv, d);
}

template <int V>
__global__ void __launch_bounds__(128) probe(unsigned int *sink, unsigned int seed, int iters) {
    fma_ballast(iters < 0, sink);
    __shared__ uint4 sm[4][8 * 2 * 32];   // per warp: 8 rows x 2 halves x 32 lanes
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 *D = sm[warp];
    uint32_t u[8], v[8], w[8], u2[8], v2[8];
    for (int i = 0; i < 8; ++i) { u[i] = seed * (i + 1) + threadIdx.x; v[i] = seed * (i + 3) ^ blockIdx.x; w[i] = seed + 7 * i; u2[i] = u[i] ^ 0x55; v2[i] = v[i] + 77; }
    u[7] &= 0x3fffffff; v[7] &= 0x3fffffff; w[7] &= 0x3fffffff; u2[7] &= 0x3fffffff; v2[7] &= 0x3fffffff;
    for (int r = 0; r < 8; ++r) { sts8(D + (r * 2) * 32 + lane, r & 1 ? v : u); }
    __syncwarp();
    if (V == 0) {
        for (int it = 0; it < iters; ++it) { uint32_t t[8]; mont_mul(t, v, w); for (int i = 0; i < 8; ++i) v[i] = t[i]; }
    } else if (V == 1) {
        for (int it = 0; it < iters; ++it) {
            uint32_t t[8], s[8], d[8];
            mont_mul(t, v, w); fr_add(s, u, t); fr_sub(d, u, t);
            for (int i = 0; i < 8; ++i) { u[i] = s[i]; v[i] = d[i]; }
        }
    } else if (V == 2) {
        for (int it = 0; it < iters; ++it) {
            const int r = (it & 3) * 2;
            uint32_t t[8], s[8], d[8];
            lds8(v, D + ((r + 1) * 2) * 32 + lane); lds8(w, D + ((r ^ 2) * 2) * 32 + lane); lds8(u, D + (r * 2) * 32 + lane);
            mont_mul(t, v, w); fr_add(s, u, t); fr_sub(d, u, t);
            sts8(D + (r * 2) * 32 + lane, s); sts8(D + ((r + 1) * 2) * 32 + lane, d);
        }
    } else if (V == 3) {
        for (int it = 0; it < iters; ++it) {
            const int r = (it & 3) * 2;
            bf_call(D + (r * 2) * 32 + lane, D + ((r + 1) * 2) * 32 + lane, D + ((r ^ 2) * 2) * 32 + lane);
        }
    } else if (V == 5) {   // two independent butterflies per iteration (register-only)
        for (int it = 0; it < iters; it += 2) {
            uint32_t t[8], s[8], d[8], t2[8], s2[8], d2[8];
            mont_mul(t, v, w); mont_mul(t2, v2, w);
            fr_add(s, u, t); fr_sub(d, u, t); fr_add(s2, u2, t2); fr_sub(d2, u2, t2);
            for (int i = 0; i < 8; ++i) { u[i] = s[i]; v[i] = d[i]; u2[i] = s2[i]; v2[i] = d2[i]; }
        }
    } else if (V == 7) {   // register-blocked: 8 elements of a row in registers, 3 stages (12 butterflies), one shared-memory round trip per 12 products
        for (int it = 0; it < iters; it += 12) {
            uint32_t x[8][8];
#pragma unroll
            for (int e = 0; e < 8; ++e) lds8(x[e], D + (e * 2) * 32 + lane);
#pragma unroll
            for (int s2 = 0; s2 < 3; ++s2) {
#pragma unroll
                for (int bf = 0; bf < 4; ++bf) {
                    const int h = 1 << s2, lo = bf & (h - 1), bu = ((bf >> s2) << (s2 + 1)) | lo;
                    uint32_t t[8], sm_[8], d[8], tw[8];
                    lds8(tw, D + (((lo + s2 + lane) & 7) * 2) * 32 + (lane ^ 1));
                    tw[7] &= 0x3fffffff;
                    mont_mul(t, x[bu + h], tw);
                    fr_add(sm_, x[bu], t); fr_sub(d, x[bu], t);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { x[bu][i] = sm_[i]; x[bu + h][i] = d[i]; }
                }
            }
            __syncwarp();
#pragma unroll
            for (int e = 0; e < 8; ++e) sts8(D + (e * 2) * 32 + lane, x[e]);
            __syncwarp();
        }
    } else if (V == 6) {   // software pipelined: the add/sub of butterfly k overlaps the product of butterfly k+1 (independent data)
        uint32_t t[8];
        mont_mul(t, v, w);
        for (int it = 0; it < iters; it += 2) {
            uint32_t t2[8], s[8], d[8];
            mont_mul(t2, v2, w); fr_add(s, u, t); fr_sub(d, u, t);
            for (int i = 0; i < 8; ++i) { u[i] = s[i]; v[i] = d[i]; }
            mont_mul(t, v, w); fr_add(s, u2, t2); fr_sub(d, u2, t2);
            for (int i = 0; i < 8; ++i) { u2[i] = s[i]; v2[i] = d[i]; }
        }
    }
    unsigned acc = 0;
    for (int i = 0; i < 8; ++i) acc ^= u[i] ^ v[i] ^ u2[i] ^ v2[i];
    uint32_t z[8]; lds8(z, D + lane); acc ^= z[0] ^ z[5];
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int V>
static double run(int wps, int iters, unsigned int *sink) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ctas = sms * wps;   // 4 warps per CTA -> wps CTAs per SM = wps warps per sub-partition
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<V><<<ctas, 128>>>(sink, 12345u, 64);
    cudaEventRecord(e0);
    probe<V><<<ctas, 128>>>(sink, 12345u, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return (double)ctas * 128 * iters / (ms * 1e-3) / 1e9;
}

int main(int argc, char **argv) {
    const int wps = argc > 1 ? atoi(argv[1]) : 6, iters = argc > 2 ? atoi(argv[2]) : 2000;
    unsigned int *sink; cudaMalloc(&sink, 256); cudaMemset(sink, 0, 256);
    printf("{\"warps_per_smsp\": %d, \"iters\": %d, \"gproducts_per_s\": {", wps, iters);
    printf("\"v0_mont_mul_chain\": %.2f, ", run<0>(wps, iters, sink));
    printf("\"v1_butterfly_registers\": %.2f, ", run<1>(wps, iters, sink));
    printf("\"v2_butterfly_smem_inline\": %.2f, ", run<2>(wps, iters, sink));
    printf("\"v3_butterfly_smem_noinline_call\": %.2f, ", run<3>(wps, iters, sink));
    printf("\"v5_two_butterflies_registers\": %.2f, ", run<5>(wps, iters, sink));
    printf("\"v6_software_pipelined\": %.2f, ", run<6>(wps, iters, sink));
    printf("\"v7_register_blocked_8x3\": %.2f}}\n", run<7>(wps, iters, sink));
    return 0;
}
