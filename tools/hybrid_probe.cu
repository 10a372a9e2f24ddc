// hybrid_probe.cu -- can the FP64 pipe and the integer multiplier be driven at the same time?
// Warps with (warp % PERIOD) < NF run a register-only FP64 multiply-accumulate loop shaped like a 12x12-limb (22-bit limbs)
// lazy Fr product (144 DFMA per term into 23 column accumulators); the other warps run the integer product loop of the
// dense kernels (64 IMAD.WIDE per term).  Reports terms/s of each kind and the aggregate.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../mpc-protocols_b200/csrc/fr.cuh"
using namespace hb;

__global__ void __launch_bounds__(256) hybrid(uint32_t *sink, uint32_t seed, int iters, int nf, int period) {
    __shared__ uint32_t sb[64][8];
    __shared__ double sd[64][12];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sb[i >> 3][i & 7] = seed * (i + 17);
    for (int i = threadIdx.x; i < 768; i += blockDim.x) sd[i / 12][i % 12] = (double)((seed * (i + 29)) & 0x3fffff);
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    if ((warp % period) < nf) {
        double a[12], col[23];
#pragma unroll
        for (int i = 0; i < 12; ++i) a[i] = (double)((seed * (i + 3) + threadIdx.x) & 0x3fffff);
#pragma unroll
        for (int i = 0; i < 23; ++i) col[i] = 0.0;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            const double *b = sd[it & 63];
#pragma unroll
            for (int i = 0; i < 12; ++i) {
#pragma unroll
                for (int j = 0; j < 12; ++j) col[i + j] = fma(a[i], b[j], col[i + j]);
            }
            if ((it & 15) == 15) {  // keep the columns bounded (stand-in for the rare normalisation)
#pragma unroll
                for (int i = 0; i < 23; ++i) col[i] *= 0.0009765625;
            }
            a[it % 12 == 0 ? 0 : 1] += 1.0;
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 23; ++i) s += col[i];
        if (s == 1.2345) sink[0] = 1;
    } else {
        uint32_t a[8], b[8];
        for (int i = 0; i < 8; ++i) a[i] = seed * (i + 3) + threadIdx.x;
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
}

int main() {
    uint32_t *sink; cudaMalloc(&sink, 64);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 2048;
    const int cfg[][2] = {{0, 8}, {8, 8}, {4, 8}, {3, 8}, {5, 8}, {2, 8}};
    for (auto &c : cfg) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(e0); hybrid<<<blocks, threads>>>(sink, 12345, iters, c[0], c[1]); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
        }
        double warps = (double)blocks * 8, fw = warps * c[0] / c[1], iw = warps - fw;
        double ft = fw * 32 * iters / best / 1e6, itt = iw * 32 * iters / best / 1e6;
        printf("fp64 warps %d/%d: %.3f ms  fp64 %.1f GMAC/s  int %.1f GMAC/s  total %.1f GMAC/s\n", c[0], c[1], best, ft, itt, ft + itt);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
