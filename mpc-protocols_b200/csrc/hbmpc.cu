// hbmpc.cu -- the C ABI (include/hbmpc_b200.h) over the sm_100a kernels.  Host logic here is launch plumbing and
// table set-up only: no share, coefficient or codeword is ever processed on the CPU, and there is no fallback --
// without a usable CUDA device every entry point returns HBMPC_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/hbmpc_b200.h"
#include "fr.cuh"
#include "matvec.cuh"
#include "ntt.cuh"
#include "robust.cuh"
#include "tables.hpp"

using namespace hb;

// ------------------------------------------------------------------------------------------------ small kernels
namespace hb {

// K5: element-wise share algebra on canonical values.  HBM-bound (96 B per element).
__global__ void __launch_bounds__(256) elementwise_kernel(int op, long long count, const uint4 *a, const uint4 *b, uint4 *out, unsigned int *err) {
    unsigned bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        uint32_t x[8], y[8], z[8];
        load_fr(x, ldg_stream(a + i * 2), ldg_stream(a + i * 2 + 1));
        load_fr(y, ldg_stream(b + i * 2), ldg_stream(b + i * 2 + 1));
        bad |= (geq_mod(x) || geq_mod(y)) ? 1u : 0u;
        if (op == 0) fr_add(z, x, y);
        else if (op == 1) fr_sub(z, x, y);
        else {
            uint32_t r2[8], xm[8];
            r2_limbs(r2);
            mont_mul(xm, x, r2);  // x*R
            mont_mul(z, xm, y);   // x*y, canonical
        }
        stg_stream(out + i * 2, make_uint4(z[0], z[1], z[2], z[3]));
        stg_stream(out + i * 2 + 1, make_uint4(z[4], z[5], z[6], z[7]));
    }
    if (bad) atomicOr(err, 1u);
}

// out[b] = in[b*stride]  (secret = coefficient 0)
__global__ void gather_first_kernel(long long B, long long stride, const uint4 *in, uint4 *out) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        out[b * 2] = in[b * stride * 2];
        out[b * 2 + 1] = in[b * stride * 2 + 1];
    }
}

// Integer-pipe roofline probes: CHAINS independent dependent-chains per thread, fully unrolled.
template <int VARIANT>
__global__ void __launch_bounds__(256) imad_probe_kernel(unsigned int *sink, unsigned int seed, int iters) {
    unsigned int m0 = seed | 1u, m1 = (seed * 2654435761u) | 1u, m2 = m0 ^ 0x9e3779b9u, m3 = m1 + 0x7f4a7c15u;
    if (VARIANT == 0) {
        unsigned int x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m0), "r"(m1));
            }
        }
        unsigned int s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s ^= x[i];
        if (s == 0x12345u) sink[0] = s;
    } else if (VARIANT == 1) {
        unsigned long long x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(m0 + i), "r"(m1));
            }
        }
        unsigned long long s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s ^= x[i];
        if (s == 0x12345ull) sink[0] = (unsigned int)s;
    } else {
        // the product kernels' pattern: 4-lane carry chains (IMAD.WIDE.U32.X) + a carry counter on the ALU pipe
        unsigned int e[16], k[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 16; ++i) e[i] = threadIdx.x + i;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                chain4(e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], k[0], m0, m1, m2, m3, m0 + u);
                chain4(e[8], e[9], e[10], e[11], e[12], e[13], e[14], e[15], k[1], m1, m2, m3, m0, m1 + u);
                chain4(e[2], e[3], e[4], e[5], e[6], e[7], e[8], e[9], k[2], m2, m3, m0, m1, m2 + u);
                chain4(e[10], e[11], e[12], e[13], e[14], e[15], e[0], e[1], k[3], m3, m0, m1, m2, m3 + u);
            }
        }
        unsigned int s = k[0] ^ k[1] ^ k[2] ^ k[3];
#pragma unroll
        for (int i = 0; i < 16; ++i) s ^= e[i];
        if (s == 0x12345u) sink[0] = s;
    }
}

}  // namespace hb

// ------------------------------------------------------------------------------------------------ context
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct RecoverTables {
    // optimistic matvec
    int R = 0, C = 0, n_chk = 0, n_gate = 0, mout = 0;
    uint4 *M = nullptr;
    int *col_map = nullptr, *chk_map = nullptr;
    // robust
    int rmax = 0, fast = 0, nsyn_max = 0;
    int *att_P = nullptr, *att_nsyn = nullptr, *att_maxL = nullptr, *order = nullptr;
    long long *att_Hoff = nullptr, *att_uoff = nullptr;
    uint4 *H = nullptr, *uinv = nullptr, *xs = nullptr, *xinv = nullptr, *Lc = nullptr, *Veval = nullptr;
};

struct hbmpc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    bool async = false;
    uint64_t launches = 0;
    std::string err;
    int num_sms = 148;
    int matvec_regs[3] = {0, 0, 0};
    int ntt_ctas[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // resident CTAs per SM of ntt_kernel<LOGN>
    bool force_dense = false;                       // HBMPC_FORCE_DENSE=1: K1/K2 through the dense matvec kernel
    unsigned int *d_status = nullptr;  // [0] non-canonical input seen, [1] first failing item, [2] failing-item count
    unsigned int *h_status = nullptr;  // pinned
    int sticky = 0;
    std::map<std::string, uint4 *> matrices;          // Vandermonde matrices keyed by "V n cols"
    std::map<std::string, RecoverTables> recover;     // keyed by (n, d, t, ids, variant)
    std::vector<void *> owned;                        // device allocations freed at destroy
    DevBuf scratch[8];
};

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return HBMPC_CUDA_ERROR;                                                                 \
        }                                                                                            \
    } while (0)

static bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

static int scratch_get(hbmpc_ctx *ctx, int slot, size_t bytes, void **out) {
    DevBuf &b = ctx->scratch[slot];
    if (b.cap < bytes) {
        if (b.p) {
            CK(cudaStreamSynchronize(ctx->stream));
            CK(cudaFree(b.p));
            b.p = nullptr;
            b.cap = 0;
        }
        size_t cap = bytes + bytes / 8 + 256;
        CK(cudaMalloc(&b.p, cap));
        b.cap = cap;
    }
    *out = b.p;
    return 0;
}

template <typename T>
static int upload(hbmpc_ctx *ctx, const std::vector<T> &v, T **out) {
    void *p = nullptr;
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    CK(cudaMalloc(&p, bytes));
    ctx->owned.push_back(p);
    if (!v.empty()) CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // v may be a temporary
    *out = (T *)p;
    return 0;
}

static int upload_fr(hbmpc_ctx *ctx, const std::vector<HFr> &v, uint4 **out) {
    std::vector<uint32_t> w;
    to_u32(v, w);
    uint32_t *p = nullptr;
    int rc = upload(ctx, w, &p);
    *out = (uint4 *)p;
    return rc;
}

extern "C" int hbmpc_ctx_create(int device, hbmpc_ctx **out) {
    if (!out) return HBMPC_INVALID_INPUT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return HBMPC_NO_DEVICE;
    }
    if (cudaSetDevice(device) != cudaSuccess) return HBMPC_NO_DEVICE;
    hbmpc_ctx *ctx = new hbmpc_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        delete ctx;
        return HBMPC_NO_DEVICE;  // sm_100a only
    }
    ctx->num_sms = prop.multiProcessorCount;
    {
        const char *fd = getenv("HBMPC_FORCE_DENSE");
        ctx->force_dense = fd && fd[0] == '1';
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void **)&ctx->d_status, 16) != cudaSuccess || cudaMallocHost((void **)&ctx->h_status, 16) != cudaSuccess) {
        delete ctx;
        return HBMPC_NO_DEVICE;
    }
    cudaMemsetAsync(ctx->d_status, 0, 16, ctx->stream);
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, matvec_kernel<1>) == cudaSuccess) ctx->matvec_regs[0] = fa.numRegs;
    if (cudaFuncGetAttributes(&fa, matvec_kernel<2>) == cudaSuccess) ctx->matvec_regs[1] = fa.numRegs;
    if (cudaFuncGetAttributes(&fa, matvec_kernel<4>) == cudaSuccess) ctx->matvec_regs[2] = fa.numRegs;
    cudaFuncSetAttribute(matvec_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(matvec_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(matvec_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (cudaGetLastError() != cudaSuccess) {
        delete ctx;
        return HBMPC_NO_DEVICE;
    }
    *out = ctx;
    return HBMPC_SUCCESS;
}

extern "C" void hbmpc_ctx_destroy(hbmpc_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (void *p : ctx->owned) cudaFree(p);
    for (auto &b : ctx->scratch)
        if (b.p) cudaFree(b.p);
    if (ctx->d_status) cudaFree(ctx->d_status);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int hbmpc_ctx_set_stream(hbmpc_ctx *ctx, void *cuda_stream) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return HBMPC_SUCCESS;
}

extern "C" int hbmpc_ctx_set_async(hbmpc_ctx *ctx, int async) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    ctx->async = async != 0;
    return HBMPC_SUCCESS;
}

// reads the device status words, folds them into a ShareErrorCode and clears them
static int collect_status(hbmpc_ctx *ctx) {
    CK(cudaMemcpyAsync(ctx->h_status, ctx->d_status, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_status, 0, 16, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int rc = ctx->sticky;
    ctx->sticky = 0;
    if (ctx->h_status[0]) rc = HBMPC_INVALID_INPUT;
    else if (!rc && ctx->h_status[2]) rc = HBMPC_DECODING_ERROR;
    return rc;
}

extern "C" int hbmpc_ctx_synchronize(hbmpc_ctx *ctx) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    return collect_status(ctx);
}

extern "C" uint64_t hbmpc_ctx_launch_count(const hbmpc_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" const char *hbmpc_last_error(const hbmpc_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// finish a call: in synchronous mode wait and report the device-side status
static int finish(hbmpc_ctx *ctx) {
    if (ctx->async) return HBMPC_SUCCESS;
    return collect_status(ctx);
}

// ------------------------------------------------------------------------------------------------ staging of host buffers
struct Staged {
    const void *user = nullptr;
    void *dev = nullptr;
    size_t bytes = 0;
    bool host = false;
};
static int stage_in(hbmpc_ctx *ctx, int slot, const void *p, size_t bytes, Staged &s) {
    s.user = p;
    s.bytes = bytes;
    s.host = !is_device_ptr(p);
    if (!s.host) {
        s.dev = const_cast<void *>(p);
        return 0;
    }
    int rc = scratch_get(ctx, slot, bytes, &s.dev);
    if (rc) return rc;
    CK(cudaMemcpyAsync(s.dev, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
static int stage_out(hbmpc_ctx *ctx, int slot, void *p, size_t bytes, Staged &s) {
    s.user = p;
    s.bytes = bytes;
    s.host = !is_device_ptr(p);
    if (!s.host) {
        s.dev = p;
        return 0;
    }
    return scratch_get(ctx, slot, bytes, &s.dev);
}
static int unstage_out(hbmpc_ctx *ctx, Staged &s, bool &need_sync) {
    if (!s.host) return 0;
    CK(cudaMemcpyAsync(const_cast<void *>(s.user), s.dev, s.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    need_sync = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ matvec launch
static int launch_matvec(hbmpc_ctx *ctx, MatvecArgs a, int flag_words) {
    if (a.B == 0) return 0;
    int regs = ctx->matvec_regs[2] > 0 ? std::max(ctx->matvec_regs[0], std::max(ctx->matvec_regs[1], ctx->matvec_regs[2])) : 96;
    MatvecPlan p = matvec_plan(a.R, a.C, flag_words, regs);
    if (p.tbt == 0) {
        ctx->err = "matvec: no launch shape fits shared memory";
        return HBMPC_INVALID_INPUT;
    }
    a.rows_per_slice = p.rows_per_slice;
    a.flag_words = flag_words;
    a.err = ctx->d_status;
    long long tile = (long long)p.tbt * 32;
    long long ntiles = (a.B + tile - 1) / tile;
    long long gx = (long long)ctx->num_sms * p.ctas_per_sm / p.slices;
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    dim3 grid((unsigned)gx, (unsigned)p.slices), block(p.warps * 32);
    switch (p.tbt) {
        case 1: matvec_kernel<1><<<grid, block, p.smem, ctx->stream>>>(a); break;
        case 2: matvec_kernel<2><<<grid, block, p.smem, ctx->stream>>>(a); break;
        default: matvec_kernel<4><<<grid, block, p.smem, ctx->stream>>>(a); break;
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}


// ------------------------------------------------------------------------------------------------ NTT launch (K1/K2 on the domain)
template <int LOGN>
static int launch_ntt_t(hbmpc_ctx *ctx, const NttArgs &a) {
    const size_t smem = ntt_smem_bytes<LOGN>();
    if (ctx->ntt_ctas[LOGN] == 0) {
        CK(cudaFuncSetAttribute(ntt_kernel<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ntt_kernel<LOGN>, 256, smem));
        ctx->ntt_ctas[LOGN] = nb > 0 ? nb : 1;
    }
    const int ipc = ntt_items_per_cta<LOGN>();
    long long ntiles = (a.B + ipc - 1) / ipc;
    long long grid = std::min<long long>(ntiles, (long long)ctx->num_sms * ctx->ntt_ctas[LOGN]);
    ntt_kernel<LOGN><<<(unsigned)grid, 256, smem, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
static int launch_ntt(hbmpc_ctx *ctx, int logn, const NttArgs &a) {
    switch (logn) {
        case 1: return launch_ntt_t<1>(ctx, a);
        case 2: return launch_ntt_t<2>(ctx, a);
        case 3: return launch_ntt_t<3>(ctx, a);
        case 4: return launch_ntt_t<4>(ctx, a);
        case 5: return launch_ntt_t<5>(ctx, a);
        case 6: return launch_ntt_t<6>(ctx, a);
        case 7: return launch_ntt_t<7>(ctx, a);
        case 8: return launch_ntt_t<8>(ctx, a);
    }
    ctx->err = "ntt: unsupported domain size";
    return HBMPC_NO_SUITABLE_DOMAIN;
}

static int get_twiddles(hbmpc_ctx *ctx, int N, uint4 **out) {
    char key[64];
    snprintf(key, sizeof key, "W %d", N);
    auto it = ctx->matrices.find(key);
    if (it != ctx->matrices.end()) {
        *out = it->second;
        return 0;
    }
    std::vector<HFr> tw = domain_elements((size_t)N, (size_t)std::max(N / 2, 1));
    uint4 *d = nullptr;
    int rc = upload_fr(ctx, tw, &d);
    if (rc) return rc;
    ctx->matrices[key] = d;
    *out = d;
    return 0;
}

// out[b][j] = sum_k w_N^(jk) in[b][k]: NTT when the zero-padded input fits the domain, dense matvec otherwise
static int apply_domain_dev(hbmpc_ctx *ctx, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *out, int recipient_major);

static int get_vandermonde(hbmpc_ctx *ctx, size_t n, size_t cols, uint4 **out) {
    char key[64];
    snprintf(key, sizeof key, "V %zu %zu", n, cols);
    auto it = ctx->matrices.find(key);
    if (it != ctx->matrices.end()) {
        *out = it->second;
        return 0;
    }
    std::vector<HFr> pts = domain_elements(n, n);
    std::vector<HFr> V = vandermonde_on_points(pts, cols);
    uint4 *d = nullptr;
    int rc = upload_fr(ctx, V, &d);
    if (rc) return rc;
    ctx->matrices[key] = d;
    *out = d;
    return 0;
}

static int apply_matrix_dev(hbmpc_ctx *ctx, const uint4 *M, size_t rows, size_t cols, size_t B, const uint64_t *in, uint64_t *out,
                            int recipient_major) {
    Staged si, so;
    int rc = stage_in(ctx, 0, in, B * cols * 32, si);
    if (rc) return rc;
    rc = stage_out(ctx, 1, out, B * rows * 32, so);
    if (rc) return rc;
    MatvecArgs a{};
    a.M = M;
    a.in = (const uint4 *)si.dev;
    a.out = (uint4 *)so.dev;
    a.R = (int)rows;
    a.C = (int)cols;
    a.B = (long long)B;
    a.in_sb = (long long)cols;
    a.in_sc = 1;
    a.in_chunk_major = 1;
    if (recipient_major) { a.out_sb = 1; a.out_sr = (long long)B; }
    else { a.out_sb = (long long)rows; a.out_sr = 1; }
    rc = launch_matvec(ctx, a, 0);
    if (rc) return rc;
    bool ns = false;
    rc = unstage_out(ctx, so, ns);
    if (rc) return rc;
    if (ns && ctx->async) CK(cudaStreamSynchronize(ctx->stream));
    return finish(ctx);
}

static int apply_domain_dev(hbmpc_ctx *ctx, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *out, int recipient_major) {
    const int N = domain_size(n);
    int logn = 0;
    while ((1 << logn) < N) ++logn;
    if (ctx->force_dense || N < 2 || cols > (size_t)N) {
        uint4 *V = nullptr;
        int rc = get_vandermonde(ctx, n, cols, &V);
        if (rc) return rc;
        return apply_matrix_dev(ctx, V, n, cols, B, in, out, recipient_major);
    }
    uint4 *tw = nullptr;
    int rc = get_twiddles(ctx, N, &tw);
    if (rc) return rc;
    Staged si, so;
    if ((rc = stage_in(ctx, 0, in, B * cols * 32, si))) return rc;
    if ((rc = stage_out(ctx, 1, out, B * n * 32, so))) return rc;
    NttArgs a{};
    a.in = (const uint4 *)si.dev;
    a.out = (uint4 *)so.dev;
    a.tw = tw;
    a.B = (long long)B;
    a.in_sb = (long long)cols;
    a.in_sc = 1;
    if (recipient_major) { a.out_sb = 1; a.out_sr = (long long)B; }
    else { a.out_sb = (long long)n; a.out_sr = 1; }
    a.cols = (int)cols;
    a.n = (int)n;
    a.err = ctx->d_status;
    if ((rc = launch_ntt(ctx, logn, a))) return rc;
    bool ns = false;
    if ((rc = unstage_out(ctx, so, ns))) return rc;
    if (ns && ctx->async) CK(cudaStreamSynchronize(ctx->stream));
    return finish(ctx);
}

extern "C" int hbmpc_compute_shares_batch(hbmpc_ctx *ctx, size_t n, size_t d, size_t B, const uint64_t *coeffs, uint64_t *shares) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (n <= d) return HBMPC_INVALID_INPUT;          // robust_interpolate.rs:59-64
    if (!domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;  // :65-66
    if (B == 0) return HBMPC_SUCCESS;
    if (!coeffs || !shares) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    return apply_domain_dev(ctx, n, d + 1, B, coeffs, shares, 0);
}

extern "C" int hbmpc_apply_vandermonde_batch(hbmpc_ctx *ctx, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *out,
                                             int recipient_major) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (cols == 0 || cols > 256) return HBMPC_INVALID_INPUT;  // row length must equal shares.len() (share/mod.rs:54-60)
    if (!domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;
    if (B == 0) return HBMPC_SUCCESS;
    if (!in || !out) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    return apply_domain_dev(ctx, n, cols, B, in, out, recipient_major);
}

extern "C" int hbmpc_apply_matrix_batch(hbmpc_ctx *ctx, size_t rows, size_t cols, const uint64_t *matrix, size_t B, const uint64_t *in,
                                        uint64_t *out, int recipient_major) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (cols == 0 || cols > 256 || rows == 0 || rows > 256 || !matrix) return HBMPC_INVALID_INPUT;
    if (B == 0) return HBMPC_SUCCESS;
    if (!in || !out) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    std::vector<HFr> M(rows * cols);
    for (size_t i = 0; i < rows * cols; ++i) {
        if (hfr::geq_mod(matrix + 4 * i)) return HBMPC_INVALID_INPUT;
        M[i] = hfr::from_canon(matrix + 4 * i);
    }
    std::vector<uint32_t> w;
    to_u32(M, w);
    void *dM = nullptr;
    int rc = scratch_get(ctx, 2, w.size() * 4, &dM);
    if (rc) return rc;
    CK(cudaMemcpyAsync(dM, w.data(), w.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return apply_matrix_dev(ctx, (const uint4 *)dM, rows, cols, B, in, out, recipient_major);
}

// ------------------------------------------------------------------------------------------------ recovery tables
static int build_recover_tables(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const std::vector<int> &order,
                                const std::vector<size_t> &sorted_ids, bool want_flags, bool secrets_only, RecoverTables &T) {
    const size_t m = d + 1, needed = d + t + 1;
    std::vector<HFr> dom = domain_elements(n, n);
    std::vector<HFr> xs(S);
    for (size_t i = 0; i < S; ++i) xs[i] = dom[sorted_ids[i]];
    std::vector<HFr> sub(xs.begin(), xs.begin() + m);
    Lagrange L = lagrange_basis(sub);
    // optimistic matrix: check rows for sorted positions m.. (gate = first t of them), then coefficient rows
    const size_t n_chk = want_flags ? S - m : t;
    std::vector<HFr> chk_pts(xs.begin() + m, xs.begin() + m + n_chk);
    std::vector<HFr> chk_rows = lagrange_eval_rows(sub, L, chk_pts);
    const size_t mout = secrets_only ? 1 : m;
    std::vector<HFr> M(chk_rows);
    M.insert(M.end(), L.Lc.begin(), L.Lc.begin() + mout * m);
    T.R = (int)(n_chk + mout);
    T.C = (int)m;
    T.n_chk = (int)n_chk;
    T.n_gate = (int)t;
    T.mout = (int)mout;
    std::vector<int> col_map(m), chk_map(std::max<size_t>(n_chk, 1));
    for (size_t i = 0; i < m; ++i) col_map[i] = order[i];
    for (size_t r = 0; r < n_chk; ++r) chk_map[r] = order[m + r];
    int rc;
    if ((rc = upload_fr(ctx, M, &T.M))) return rc;
    if ((rc = upload(ctx, col_map, &T.col_map))) return rc;
    if ((rc = upload(ctx, chk_map, &T.chk_map))) return rc;
    if ((rc = upload(ctx, order, &T.order))) return rc;

    // robust tables
    T.rmax = (int)std::min(t, S - needed);
    T.fast = T.rmax >= 1 ? 1 : 0;
    const int natt = 1 + T.rmax;
    std::vector<int> att_P(natt), att_nsyn(natt), att_maxL(natt);
    std::vector<long long> att_Hoff(natt), att_uoff(natt);
    std::vector<HFr> H, U;
    // uinv_i^{(P)} = prod_{l<P, l != i} (x_i - x_l), built incrementally over P
    std::vector<HFr> uinv;
    auto extend_to = [&](size_t P) {
        while (uinv.size() < P) {
            size_t p = uinv.size();
            HFr self = hfr::ONE;
            for (size_t i = 0; i < p; ++i) {
                HFr df = hfr::sub(xs[i], xs[p]);
                uinv[i] = hfr::mul(uinv[i], df);
                self = hfr::mul(self, hfr::neg(df));
            }
            uinv.push_back(self);
        }
    };
    auto emit = [&](int a, size_t P, size_t nsyn, size_t maxL) {
        att_P[a] = (int)P;
        att_nsyn[a] = (int)nsyn;
        att_maxL[a] = (int)maxL;
        att_Hoff[a] = (long long)H.size();
        att_uoff[a] = (long long)U.size();
        std::vector<HFr> u(uinv.begin(), uinv.begin() + P);
        HFr one_c = {{1, 0, 0, 0}};
        for (size_t i = 0; i < P; ++i) U.push_back(hfr::mul(u[i], one_c));  // canonical
        hfr::batch_inv(u);
        size_t base = H.size();
        H.resize(base + nsyn * P);
        for (size_t i = 0; i < P; ++i) {
            HFr p = hfr::mul(u[i], hfr::R2);  // u_i * R^2
            for (size_t j = 0; j < nsyn; ++j) {
                H[base + j * P + i] = p;
                p = hfr::mul(p, xs[i]);
            }
        }
        T.nsyn_max = std::max(T.nsyn_max, (int)nsyn);
    };
    if (T.rmax >= 1) {
        for (int r = 1; r <= T.rmax; ++r) {
            size_t P = needed + r;
            extend_to(P);
            emit(r, P, t + r, (size_t)r);
        }
        extend_to(S);
        emit(0, S, S - m, std::min(t, (S - m) / 2));
    } else {
        att_P[0] = (int)S; att_nsyn[0] = 0; att_maxL[0] = 0; att_Hoff[0] = 0; att_uoff[0] = 0;
    }
    std::vector<HFr> xinv(xs);
    hfr::batch_inv(xinv);
    std::vector<HFr> Veval = vandermonde_on_points(xs, m);
    if ((rc = upload(ctx, att_P, &T.att_P))) return rc;
    if ((rc = upload(ctx, att_nsyn, &T.att_nsyn))) return rc;
    if ((rc = upload(ctx, att_maxL, &T.att_maxL))) return rc;
    if ((rc = upload(ctx, att_Hoff, &T.att_Hoff))) return rc;
    if ((rc = upload(ctx, att_uoff, &T.att_uoff))) return rc;
    if ((rc = upload_fr(ctx, H, &T.H))) return rc;
    if ((rc = upload_fr(ctx, U, &T.uinv))) return rc;
    if ((rc = upload_fr(ctx, xs, &T.xs))) return rc;
    if ((rc = upload_fr(ctx, xinv, &T.xinv))) return rc;
    if ((rc = upload_fr(ctx, L.Lc, &T.Lc))) return rc;
    if ((rc = upload_fr(ctx, Veval, &T.Veval))) return rc;
    return 0;
}

// shared implementation of K3 / K4-direct: element (item b, arrival j) of `in` at (b*in_sb + j*in_sc)
static int recover_impl(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *ids, size_t B, const uint64_t *in,
                        bool sender_major, uint64_t *coeffs, bool secrets_only, uint64_t *secrets, int32_t *path, uint64_t *flags) {
    // validation order of robust_interpolate.rs:290-341 / :100-142
    if (n < 3 * t + 1) return HBMPC_INVALID_INPUT;
    if (S == 0 || !ids) return HBMPC_INVALID_INPUT;
    if (B == 0) return HBMPC_INVALID_INPUT;  // "Empty batch"
    if (S > 256) return HBMPC_INVALID_INPUT;
    std::vector<int> order(S);
    for (size_t i = 0; i < S; ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ids[a] < ids[b]; });
    std::vector<size_t> sorted_ids(S);
    for (size_t i = 0; i < S; ++i) sorted_ids[i] = ids[order[i]];
    for (size_t i = 0; i < S; ++i) {
        if (i > 0 && sorted_ids[i] == sorted_ids[i - 1]) return HBMPC_INVALID_INPUT;
        if (sorted_ids[i] >= n) return HBMPC_INVALID_INPUT;
    }
    const size_t needed = d + t + 1, m = d + 1;
    if (S < needed) return HBMPC_INVALID_INPUT;
    if (!domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;
    if (!in || !path || (!coeffs && !secrets_only) || (secrets_only && !secrets)) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);

    const bool want_flags = flags != nullptr;
    std::string key = "R " + std::to_string(n) + " " + std::to_string(d) + " " + std::to_string(t) + (want_flags ? " f" : " -") +
                      (secrets_only ? " s" : " c");
    for (size_t i = 0; i < S; ++i) key += " " + std::to_string(ids[i]);
    auto it = ctx->recover.find(key);
    if (it == ctx->recover.end()) {
        RecoverTables T;
        int rc = build_recover_tables(ctx, n, d, t, S, order, sorted_ids, want_flags, secrets_only, T);
        if (rc) return rc;
        it = ctx->recover.emplace(key, T).first;
    }
    const RecoverTables &T = it->second;
    const int fw = want_flags ? (int)((S + 63) / 64) : 0;

    Staged si, sc, ss, sp, sf;
    int rc;
    if ((rc = stage_in(ctx, 0, in, B * S * 32, si))) return rc;
    uint64_t *co_user = secrets_only ? secrets : coeffs;
    if ((rc = stage_out(ctx, 1, co_user, B * T.mout * 32, sc))) return rc;
    if ((rc = stage_out(ctx, 2, path, B * 4, sp))) return rc;
    if (want_flags && (rc = stage_out(ctx, 3, flags, B * fw * 8, sf))) return rc;
    if (!secrets_only && secrets && (rc = stage_out(ctx, 4, secrets, B * 32, ss))) return rc;
    // scratch: fail bytes + list + counter
    void *aux = nullptr;
    size_t aux_bytes = ((B + 15) / 16) * 16 + B * 4 + 16;
    if ((rc = scratch_get(ctx, 5, aux_bytes, &aux))) return rc;
    unsigned char *fail = (unsigned char *)aux;
    unsigned int *list = (unsigned int *)((char *)aux + ((B + 15) / 16) * 16);
    unsigned int *count = list + B;
    CK(cudaMemsetAsync(fail, 0, ((B + 15) / 16) * 16, ctx->stream));
    CK(cudaMemsetAsync(count, 0, 16, ctx->stream));
    CK(cudaMemsetAsync(sp.dev, 0, B * 4, ctx->stream));
    if (want_flags) CK(cudaMemsetAsync(sf.dev, 0, B * fw * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_status + 1, 0xff, 4, ctx->stream));

    MatvecArgs a{};
    a.M = T.M;
    a.in = (const uint4 *)si.dev;
    a.out = (uint4 *)sc.dev;
    a.R = T.R;
    a.C = T.C;
    a.B = (long long)B;
    if (sender_major) { a.in_sb = 1; a.in_sc = (long long)B; a.in_chunk_major = 0; }
    else { a.in_sb = (long long)S; a.in_sc = 1; a.in_chunk_major = 1; }
    a.out_sb = T.mout;
    a.out_sr = 1;
    a.col_map = T.col_map;
    a.n_chk = T.n_chk;
    a.n_gate = T.n_gate;
    a.chk_map = T.chk_map;
    a.fail = fail;
    a.flags = want_flags ? (unsigned long long *)sf.dev : nullptr;
    if ((rc = launch_matvec(ctx, a, fw))) return rc;

    compact_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(fail, (long long)B, list, count);
    ctx->launches++;
    CK(cudaGetLastError());

    {
        RobustArgs r{};
        r.in = (const uint4 *)si.dev;
        r.in_sb = a.in_sb;
        r.in_sc = a.in_sc;
        r.B = (long long)B;
        r.list = list;
        r.count = count;
        r.S = (int)S; r.m = (int)m; r.t = (int)t; r.needed = (int)needed; r.rmax = T.rmax; r.fast = T.fast;
        r.att_P = T.att_P; r.att_nsyn = T.att_nsyn; r.att_maxL = T.att_maxL; r.att_Hoff = T.att_Hoff; r.att_uoff = T.att_uoff;
        r.H = T.H; r.uinv = T.uinv; r.xs = T.xs; r.xinv = T.xinv; r.Lc = T.Lc; r.Veval = T.Veval; r.order = T.order;
        r.coeffs = (uint4 *)sc.dev;
        r.mout = T.mout;
        r.path = (int *)sp.dev;
        r.flags = a.flags;
        r.flag_words = fw;
        r.first_fail = ctx->d_status + 1;
        r.fail_any = ctx->d_status + 2;
        WsLayout lay(T.nsyn_max, (int)t);
        const int threads = 128;
        long long blocks = std::min<long long>((long long)ctx->num_sms * 4, (long long)((B + threads - 1) / threads));
        if (blocks < 1) blocks = 1;
        void *ws = nullptr;
        if ((rc = scratch_get(ctx, 6, (size_t)blocks * threads * lay.total * 32 + 64, &ws))) return rc;
        r.ws = (uint4 *)ws;
        r.ws_elems = lay.total;
        robust_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(r);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    if (!secrets_only && secrets) {
        gather_first_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>((long long)B, (long long)m, (const uint4 *)sc.dev, (uint4 *)ss.dev);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    bool ns = false;
    if ((rc = unstage_out(ctx, sc, ns))) return rc;
    if ((rc = unstage_out(ctx, sp, ns))) return rc;
    if (want_flags && (rc = unstage_out(ctx, sf, ns))) return rc;
    if (!secrets_only && secrets && (rc = unstage_out(ctx, ss, ns))) return rc;
    if (ns && ctx->async) CK(cudaStreamSynchronize(ctx->stream));
    return finish(ctx);
}

extern "C" int hbmpc_batch_recover(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B,
                                   const uint64_t *evals, uint64_t *coeffs, int32_t *path, uint64_t *flags) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    return recover_impl(ctx, n, d, t, S, sender_ids, B, evals, true, coeffs, false, nullptr, path, flags);
}

extern "C" int hbmpc_batch_recover_secrets(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B,
                                           const uint64_t *evals, uint64_t *secrets, int32_t *path) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    return recover_impl(ctx, n, d, t, S, sender_ids, B, evals, true, nullptr, true, secrets, path, nullptr);
}

extern "C" int hbmpc_robust_interpolate_batch(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *ids, size_t B,
                                              const uint64_t *shares, uint64_t *coeffs, uint64_t *secrets, int32_t *path, uint64_t *flags) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    return recover_impl(ctx, n, d, t, S, ids, B, shares, false, coeffs, false, secrets, path, flags);
}

extern "C" int hbmpc_elementwise(hbmpc_ctx *ctx, int op, size_t count, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (op < 0 || op > 2) return HBMPC_INVALID_INPUT;
    if (count == 0) return HBMPC_SUCCESS;
    if (!a || !b || !out) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    Staged sa, sb, so;
    int rc;
    if ((rc = stage_in(ctx, 0, a, count * 32, sa))) return rc;
    if ((rc = stage_in(ctx, 1, b, count * 32, sb))) return rc;
    if ((rc = stage_out(ctx, 2, out, count * 32, so))) return rc;
    long long blocks = std::min<long long>((long long)ctx->num_sms * 8, (long long)((count + 255) / 256));
    elementwise_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(op, (long long)count, (const uint4 *)sa.dev, (const uint4 *)sb.dev, (uint4 *)so.dev,
                                                                  ctx->d_status);
    ctx->launches++;
    CK(cudaGetLastError());
    bool ns = false;
    if ((rc = unstage_out(ctx, so, ns))) return rc;
    if (ns && ctx->async) CK(cudaStreamSynchronize(ctx->stream));
    return finish(ctx);
}

extern "C" int hbmpc_measure_imad_peak(hbmpc_ctx *ctx, int variant, double *giga_inst_per_s, double *elapsed_ms) {
    if (!ctx || variant < 0 || variant > 2 || !giga_inst_per_s) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    void *sink = nullptr;
    int rc = scratch_get(ctx, 7, 256, &sink);
    if (rc) return rc;
    const int iters = 4096, blocks = ctx->num_sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0, ctx->stream));
        if (variant == 0) imad_probe_kernel<0><<<blocks, threads, 0, ctx->stream>>>((unsigned int *)sink, 12345u + rep, iters);
        else if (variant == 1) imad_probe_kernel<1><<<blocks, threads, 0, ctx->stream>>>((unsigned int *)sink, 12345u + rep, iters);
        else imad_probe_kernel<2><<<blocks, threads, 0, ctx->stream>>>((unsigned int *)sink, 12345u + rep, iters);
        ctx->launches++;
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // thread-level multiply-add instructions per thread: variant 0: 64/iter; 1: 64/iter; 2: 4*4*8 = 128/iter (+16 ALU adds)
    double per_thread = variant == 2 ? 128.0 * iters : 64.0 * iters;
    double total = per_thread * (double)blocks * threads;
    *giga_inst_per_s = total / (best * 1e-3) / 1e9;
    if (elapsed_ms) *elapsed_ms = best;
    return HBMPC_SUCCESS;
}
