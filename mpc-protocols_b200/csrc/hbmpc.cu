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
#include <initializer_list>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hbmpc_b200.h"
#include "fr.cuh"
#include "matvec.cuh"
#include "ntt.cuh"
#include "ntt16x.cuh"
#include "robust.cuh"
#include "sampler.cuh"
#include "tables.hpp"

using namespace hb;

// ------------------------------------------------------------------------------------------------ small kernels
namespace hb {

// K5: element-wise share algebra on canonical values.  HBM-bound (96 B per element).
__global__ void __launch_bounds__(256) elementwise_kernel(int op, long long count, const uint4 *a, const uint4 *b, uint4 *out, unsigned int *err) {
    fma_ballast(count < 0, err);
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
    if (bad) *(volatile unsigned int *)err = 1u;  // status words live in mapped host memory: plain store, every writer stores 1
}

// K5 fused: the share algebra of ONE protocol step in one pass over HBM instead of one pass per operator.
//   OP 0  triple mask      out0 = in0*in1 - in2                              (triple_generation.rs:332-340)  128 B per element instead of 192
//   OP 1  Beaver mask      out0 = in0 - in1, out1 = in2 - in3                (multiplication.rs:417-426)     one launch instead of two
//   OP 2  Beaver finalise  out0 = in0 - in3*in4 - in3*in2 - in4*in1          (multiplication.rs:79-97)       192 B per element instead of 576
// Values are canonical at the interface; the results are the canonical residues, i.e. bit-identical to the operator-by-operator route.
struct FusedArgs {
    const uint4 *in[5];
    uint4 *out[2];
    long long count;
    unsigned int *err;
};
template <int OP>
__global__ void __launch_bounds__(256) elementwise_fused_kernel(const FusedArgs a) {
    fma_ballast(a.count < 0, a.err);
    constexpr int NIN = OP == 0 ? 3 : OP == 1 ? 4 : 5;
    unsigned bad = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += (long long)gridDim.x * blockDim.x) {
        uint32_t v[NIN][8];
#pragma unroll
        for (int k = 0; k < NIN; ++k) {
            load_fr(v[k], ldg_stream(a.in[k] + i * 2), ldg_stream(a.in[k] + i * 2 + 1));
            bad |= geq_mod(v[k]) ? 1u : 0u;
        }
        uint32_t z[8];
        if (OP == 0) {
            k5_triple_mask(z, v[0], v[1], v[2]);
        } else if (OP == 1) {
            uint32_t z1[8];
            fr_sub(z, v[0], v[1]);
            fr_sub(z1, v[2], v[3]);
            stg_stream(a.out[1] + i * 2, make_uint4(z1[0], z1[1], z1[2], z1[3]));
            stg_stream(a.out[1] + i * 2 + 1, make_uint4(z1[4], z1[5], z1[6], z1[7]));
        } else {
            k5_beaver_finalize(z, v[0], v[1], v[2], v[3], v[4]);
        }
        stg_stream(a.out[0] + i * 2, make_uint4(z[0], z[1], z[2], z[3]));
        stg_stream(a.out[0] + i * 2 + 1, make_uint4(z[4], z[5], z[6], z[7]));
    }
    if (bad) *(volatile unsigned int *)a.err = 1u;
}

// out[b] = in[b*stride]  (secret = coefficient 0)
__global__ void gather_first_kernel(long long B, long long stride, const uint4 *in, uint4 *out) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        out[b * 2] = in[b * stride * 2];
        out[b * 2 + 1] = in[b * stride * 2 + 1];
    }
}

// NonRobustShare::recover_secret epilogue: status[b] = degree of the interpolant (highest non-zero coefficient, 0 for the zero
// polynomial like DensePolynomial::degree) or -DegreeMismatch when a coefficient above `deg` was non-zero; failing items are zeroed.
__global__ void degree_status_kernel(long long B, int m, uint4 *coeffs, const unsigned char *fail, int *status, uint4 *secrets) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        uint4 *c = coeffs + b * m * 2;
        int st = 0;
        if (fail[b]) {
            st = -HBMPC_DEGREE_MISMATCH;
            for (int k = 0; k < 2 * m; ++k) c[k] = make_uint4(0, 0, 0, 0);
        } else {
            for (int k = m - 1; k > 0; --k) {
                uint4 lo = c[2 * k], hi = c[2 * k + 1];
                if (lo.x | lo.y | lo.z | lo.w | hi.x | hi.y | hi.z | hi.w) { st = k; break; }
            }
        }
        status[b] = st;
        if (secrets) { secrets[2 * b] = c[0]; secrets[2 * b + 1] = c[1]; }
    }
}

// N1 wire records: ark-serialize writes a ShamirShare<F,1,_> as 32-byte LE canonical value + u64 id + u64 degree (48 bytes,
// 8-byte aligned inside a payload).  Split / join such records without a host-side repacking pass.
__global__ void unpack_records_kernel(long long count, const unsigned long long *rec, unsigned long long *values, unsigned long long *ids,
                                      unsigned long long *degrees) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long *r = rec + i * 6;
        values[i * 4 + 0] = r[0]; values[i * 4 + 1] = r[1]; values[i * 4 + 2] = r[2]; values[i * 4 + 3] = r[3];
        if (ids) ids[i] = r[4];
        if (degrees) degrees[i] = r[5];
    }
}
__global__ void pack_records_kernel(long long count, const unsigned long long *values, long long per_id, unsigned long long degree, unsigned long long *rec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long *r = rec + i * 6;
        r[0] = values[i * 4 + 0]; r[1] = values[i * 4 + 1]; r[2] = values[i * 4 + 2]; r[3] = values[i * 4 + 3];
        r[4] = (unsigned long long)(i / per_id);
        r[5] = degree;
    }
}

// Integer-pipe roofline probes (register-only, multiplicands depend on the running values so nothing is hoisted):
//   0: mad.lo.u32 (IMAD)            -- the "IMAD peak" of the north star: 64 lanes/clk/SM
//   1: IMAD.WIDE.U32(.X) 4-lane carry chains, the product kernels' instruction (32x32->64 multiply-add)
//   2: DFMA chains (FP64 pipe, for reference)
template <int VARIANT>
__global__ void __launch_bounds__(256) imad_probe_kernel(unsigned int *sink, unsigned int seed, int iters) {
    if (VARIANT == 0) {
        unsigned int x[16], m1 = (seed * 2654435761u) | 1u;
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int i = 0; i < 16; ++i) asm volatile("mad.lo.u32 %0, %0, %0, %1;" : "+r"(x[i]) : "r"(m1));
            }
        }
        unsigned int s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s ^= x[i];
        if (s == 0x12345u) sink[0] = s;
    } else if (VARIANT == 1) {
        unsigned long long l[8];
        unsigned int k = 0, m1 = seed * 77u + 5u;
#pragma unroll
        for (int i = 0; i < 8; ++i) l[i] = (unsigned long long)(threadIdx.x * 2654435761u + i) << 7;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                chain4w(l[0], l[1], l[2], l[3], k, (unsigned int)l[4], (unsigned int)l[5], (unsigned int)l[6], (unsigned int)l[7], m1 + u);
                chain4w(l[4], l[5], l[6], l[7], k, (unsigned int)l[0], (unsigned int)l[1], (unsigned int)l[2], (unsigned int)l[3], m1 - u);
            }
        }
        unsigned long long s = k;
#pragma unroll
        for (int i = 0; i < 8; ++i) s ^= l[i];
        if (s == 0x12345ull) sink[0] = (unsigned int)s;
    } else {
        double x[8], m = 1.0 + seed * 1e-9, c = seed * 1e-7;
#pragma unroll
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
#pragma unroll
        for (int i = 0; i < 8; ++i) s += x[i];
        if (s == 1.2345) sink[0] = 1u;
    }
}

// Latency probe of the multi-limb multiply-add: CH independent carry chains per thread (each a serial IMAD.WIDE.U32.X chain, as one
// row of a Montgomery product is), launched at a chosen number of resident warps per SM sub-partition.  Tells how much
// instruction-level parallelism x occupancy a product kernel needs before the multiplier pipe, not the chain latency, is the bound.
template <int CH>
__global__ void __launch_bounds__(128) wide_chain_probe_kernel(unsigned int *sink, unsigned int seed, int iters) {
    unsigned long long l[CH][4];
    unsigned int k = 0, m1 = seed * 77u + 5u;
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) l[c][i] = (unsigned long long)(threadIdx.x * 2654435761u + i + 4 * c) << 7;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int c = 0; c < CH; ++c)
                chain4w(l[c][0], l[c][1], l[c][2], l[c][3], k, (unsigned int)l[c][3], (unsigned int)l[c][2], (unsigned int)l[c][1], (unsigned int)l[c][0], m1 + u);
        }
    }
    unsigned long long s = k;
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) s ^= l[c][i];
    if (s == 0x12345ull) sink[0] = (unsigned int)s;
}

// Throughput probe of the Montgomery product itself: ILP independent register-resident product chains x <- x * w per thread.
template <int ILP>
__global__ void __launch_bounds__(128) mont_mul_probe_kernel(unsigned int *sink, unsigned int seed, int iters) {
    fma_ballast(iters < 0, sink);
    uint32_t x[ILP][8], w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = (seed + 0x9e3779b9u * (i + 1)) >> (i == 7 ? 2 : 0);
#pragma unroll
    for (int c = 0; c < ILP; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[c][i] = (threadIdx.x * 2654435761u + 977u * i + c) >> (i == 7 ? 2 : 0);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (ILP == 2) {
            uint32_t y0[8], y1[8];
            mont_mul2(y0, x[0], w, y1, x[ILP - 1], w);
#pragma unroll
            for (int i = 0; i < 8; ++i) { x[0][i] = y0[i]; x[ILP - 1][i] = y1[i]; }
        } else {
#pragma unroll
            for (int c = 0; c < ILP; ++c) {
                uint32_t y[8];
                mont_mul(y, x[c], w);
#pragma unroll
                for (int i = 0; i < 8; ++i) x[c][i] = y[i];
            }
        }
    }
    uint32_t s2 = 0;
#pragma unroll
    for (int c = 0; c < ILP; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) s2 ^= x[c][i];
    if (s2 == 0x12345u) sink[0] = s2;
}

}  // namespace hb

// ------------------------------------------------------------------------------------------------ context
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

// A lane = a stream with its own scratch.  Lane 0 is the context's (user-visible) stream and serves calls whose buffers
// are all device pointers; lanes 1..3 pipeline calls with host buffers chunk by chunk (H2D | kernels | D2H overlap).  In
// asynchronous mode host-buffer calls only enqueue and alternate between lanes 1..3 and 4..6, so that two consecutive calls run
// concurrently: a download-bound call (share generation) and an upload-bound one (recovery) then use both PCIe directions at once.
struct Lane {
    cudaStream_t stream = nullptr;
    DevBuf scratch[14];
};
static const int NLANES = 7, LANES_PER_SET = 3;

struct RecoverTables {
    // optimistic matvec
    int R = 0, C = 0, n_chk = 0, n_gate = 0, mout = 0;
    uint4 *M = nullptr;
    // robust
    int rmax = 0, fast = 0, nsyn_max = 0;
    int *att_P = nullptr, *att_nsyn = nullptr, *att_maxL = nullptr;
    long long *att_uoff = nullptr;
    int *sid = nullptr;
    uint4 *u2 = nullptr, *tw = nullptr, *ritw = nullptr, *uinv = nullptr, *xs = nullptr, *xinv = nullptr, *Lc = nullptr, *Veval = nullptr;
    // staged decoder (fast attempt on all S shares): syndrome weights by domain index, domain index -> sorted position
    uint4 *syn_wt = nullptr;
    int *pos_of_dom = nullptr;
    int nsyn0 = 0, maxL0 = 0;
    long long uoff0 = 0;
    // all-shares-present fast path (S == n == N): inverse NTT + degree check
    int fast_logn = 0;
    uint4 *itw = nullptr, *iscale = nullptr;
    // general optimistic check without flags: erasure-weighted inverse NTT + triangular coefficient recovery
    int er_logn = 0, er_zero_from = 0;
    bool er_all = false;  // tables built over ALL supplied ids (calls with flags): a chunk that passes has every flag clear
    int *er_row_len = nullptr, *er_row_start = nullptr;
    int er_h1 = 0, er_h2 = 0, er_hi_top = 0;   // two-sided recovery: low / high rows, index of the top coefficient of Q
    uint4 *er_wt = nullptr, *er_tri = nullptr;
    std::vector<void *> allocs;  // device memory of this entry (freed when the entry is evicted)
};

// a10 tables of one (n, deg, id SET): rows over the ids in ascending order; arrival order enters through per-call maps
struct NonRobustTables {
    uint4 *M = nullptr;
    int *chk_map = nullptr;
    int R = 0, n_chk = 0;
    // every domain point supplied (S == n == N): interpolation through all points is one inverse NTT
    int fast_logn = 0;
    uint4 *itw = nullptr, *iscale = nullptr;
    std::vector<void *> allocs;  // device memory of this entry (freed when the entry is evicted)
};

// Cache key of the constant tables of one (n, d, t, id SET, variant): the set is a 256-bit bitmap, so building and comparing a
// key costs a few words per call (arrival order is not part of it: it only changes the small per-call index maps).
struct TableKey {
    uint32_t kind = 0, n = 0, d = 0, t = 0;  // kind: bit0 flags wanted, bit1 secrets only, bit2 non-robust (a10) tables
    uint64_t idset[4] = {0, 0, 0, 0};
    bool operator<(const TableKey &o) const { return memcmp(this, &o, sizeof(TableKey)) < 0; }
};
static_assert(sizeof(TableKey) == 48, "TableKey must not contain padding (it is compared with memcmp)");

// Per-call index maps (arrival order -> sorted position and back) travel through a small ring of pinned host blocks, each with
// its own device block: the copy is a true asynchronous DMA from pinned memory, and a block is only refilled after the copy
// that read it has completed (event per block).  Device blocks are reused in stream order.
struct MapRing {
    static const int SLOTS = 8, INTS = 2048;
    int *h[SLOTS] = {};
    int *d[SLOTS] = {};
    cudaEvent_t ev[SLOTS] = {};
    bool used[SLOTS] = {};
    int next = 0;
};

struct hbmpc_ctx {
    int device = 0;
    Lane lanes[NLANES];
    bool own_stream = true;
    bool async = false;
    uint64_t launches = 0;
    std::string err;
    int num_sms = 148;
    int matvec_regs = 0;
    int ntt_ctas[6][9] = {};
    int ntt64_ctas[2] = {};                         // resident CTAs per SM of ntt64_cta_kernel<MODE>
    int ntt16x_ctas[3][8] = {};                     // resident CTAs per SM of ntt16x_kernel<LOGN, MODE>
    long long ntt16x_min = 16384;                   // HBMPC_NTT16X: 0 never use ntt16x_kernel, 2 use it for every batch size (tests); default: batches >= 16384
    bool ntt_cta = true;                            // HBMPC_NTT_CTA=0: never use ntt64_cta_kernel; 2: use it for every 64-point transform of any size
    bool ntt_cta_all = false;
    long long ntt_cta_min = 1024;                        // resident CTAs per SM of ntt_kernel<LOGN, MODE>
    size_t scan_max = 65536;                        // HBMPC_SCAN_MAX: batches up to this size skip the compaction pass
    bool no_staged_direct = false;                  // HBMPC_NO_STAGED_DIRECT=1: failing items of the all-points check take the dense check first
    bool no_speculation = false;                    // HBMPC_NO_SPECULATION=1: never try the persistent-attacker shortcut
    size_t staged_min = 4096;                       // HBMPC_STAGED_MIN: failing sets of at least this many items use the staged decoder
    int staged_seg = 16;                            // HBMPC_STAGED_SEG: Berlekamp-Massey iterations between two re-sorts
    unsigned int *h_spec = nullptr;                 // pinned: failing-item count + per-sender error histogram of the scout pass
    bool attack_seen = false;                       // the last recovery calls met large failing sets: batches <= scan_max are compacted too
    unsigned int dev_route_calls = 0;               // asynchronous recovery calls since the last status read that took the device-count staged route
    bool async_staged = true;                       // HBMPC_ASYNC_STAGED=0: asynchronous calls never take it; 2: always, at any batch size (tests)
    bool force_async_staged = false;
    bool no_sync_count = false;                     // HBMPC_NO_SYNC_COUNT=1: synchronous mid-size calls keep round 1's scan route (A/B runs)
    bool no_er_flags = false;                       // HBMPC_NO_ER_FLAGS=1: calls with flags on a sender subset go straight to the dense check
    bool no_fastpath = false;                       // HBMPC_NO_FASTPATH=1: K3 never takes the all-shares-present inverse-NTT path
    bool force_dense = false;                       // HBMPC_FORCE_DENSE=1: K1/K2 through the dense matvec kernel
    size_t chunk_bytes = 128u << 20;                // HBMPC_CHUNK_MB: target bytes per pipelined host copy (measured, e2e step of bench.py: 16 MB 75.3, 64 MB 71.8, 128 MB 67.6, 256 MB 67.6 ms)
    // status words in mapped pinned host memory (device view d_status, host view h_status): kernels store 1 on the rare
    // error, the host reads them after a stream synchronize -- no copy, no memset on the call path
    unsigned int *d_status = nullptr;  // [0] non-canonical input seen, [2] some item failed to decode
    unsigned int *h_status = nullptr;
    cudaEvent_t ev_main = nullptr;
    unsigned int *h_counts = nullptr;  // pinned: per-chunk count of items that failed the optimistic check (lean host path)
    size_t h_counts_cap = 0;
    std::map<std::string, uint4 *> matrices;          // Vandermonde / twiddle tables keyed by "V n cols" / "W N"
    std::map<TableKey, RecoverTables> recover;        // keyed by (n, d, t, id set, variant); bounded (256 entries)
    std::map<TableKey, struct NonRobustTables> *nonrobust = nullptr;   // same for the a10 tables
    std::vector<void *> owned;                        // device allocations freed at destroy
    MapRing maps;                                     // per-call arrival-order index maps (order, col_map, chk_map, in_map)
    int lane_set = 0;                                   // async mode: lane set (0: lanes 1..3, 1: lanes 4..6) of the next host-buffer call
    // dynamic tile queues of ntt16x_kernel: one self-resetting pair of counters per stream (kernels of one stream run in order)
    unsigned long long *work_pool = nullptr;            // [WORK_SLOTS][4]
    std::vector<cudaStream_t> work_streams;
    bool static_tiles = false;                          // HBMPC_STATIC_TILES=1: static round-robin tiles (measurement knob)
    cudaStream_t main_stream() const { return lanes[0].stream; }
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

static int scratch_get(hbmpc_ctx *ctx, Lane &ln, int slot, size_t bytes, void **out) {
    DevBuf &b = ln.scratch[slot];
    if (b.cap < bytes) {
        if (b.p) {
            CK(cudaStreamSynchronize(ln.stream));
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
static int upload(hbmpc_ctx *ctx, const std::vector<T> &v, T **out, std::vector<void *> *owner = nullptr) {
    void *p = nullptr;
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    CK(cudaMalloc(&p, bytes));
    (owner ? *owner : ctx->owned).push_back(p);
    if (!v.empty()) {
        // the consumers run on non-blocking streams that are not ordered against the legacy stream of a plain cudaMemcpy (which
        // may return before a pageable copy has landed): copy on the context's stream and wait for it -- tables are built once
        CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->main_stream()));
        CK(cudaStreamSynchronize(ctx->main_stream()));
    }
    *out = (T *)p;
    return 0;
}

// one block of the per-call map ring: fills it through `fill(int *host)` (returns the number of ints) and enqueues the copy
template <typename Fill>
static int maps_upload(hbmpc_ctx *ctx, Fill fill, const int **dev) {
    MapRing &r = ctx->maps;
    const int i = r.next;
    r.next = (r.next + 1) % MapRing::SLOTS;
    if (!r.h[i]) {
        CK(cudaMallocHost((void **)&r.h[i], MapRing::INTS * sizeof(int)));
        CK(cudaMalloc((void **)&r.d[i], MapRing::INTS * sizeof(int)));
        CK(cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming));
    }
    if (r.used[i]) CK(cudaEventSynchronize(r.ev[i]));
    const size_t count = fill(r.h[i]);
    CK(cudaMemcpyAsync(r.d[i], r.h[i], count * sizeof(int), cudaMemcpyHostToDevice, ctx->main_stream()));
    CK(cudaEventRecord(r.ev[i], ctx->main_stream()));
    r.used[i] = true;
    *dev = r.d[i];
    return 0;
}

static int upload_fr(hbmpc_ctx *ctx, const std::vector<HFr> &v, uint4 **out, std::vector<void *> *owner = nullptr) {
    std::vector<uint32_t> w;
    to_u32(v, w);
    uint32_t *p = nullptr;
    int rc = upload(ctx, w, &p, owner);
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
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return HBMPC_NO_DEVICE;  // sm_100a only
    hbmpc_ctx *ctx = new hbmpc_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    {
        const char *fd = getenv("HBMPC_FORCE_DENSE");
        ctx->force_dense = fd && fd[0] == '1';
        const char *ne = getenv("HBMPC_NO_ER_FLAGS");
        ctx->no_er_flags = ne && ne[0] == '1';
        const char *nf = getenv("HBMPC_NO_FASTPATH");
        ctx->no_fastpath = nf && nf[0] == '1';
        const char *sx = getenv("HBMPC_SCAN_MAX");
        if (sx) ctx->scan_max = (size_t)atoll(sx);
        const char *ns = getenv("HBMPC_NO_SPECULATION");
        ctx->no_speculation = ns && ns[0] == '1';
        const char *nc = getenv("HBMPC_NTT_CTA");
        if (nc) { ctx->ntt_cta = nc[0] == '1' || nc[0] == '2'; if (nc[0] == '2') { ctx->ntt_cta_min = 1; ctx->ntt_cta_all = true; } }
        const char *nx = getenv("HBMPC_NTT16X");
        if (nx) ctx->ntt16x_min = nx[0] == '0' ? LLONG_MAX : (nx[0] == '2' ? 1 : ctx->ntt16x_min);
        const char *nd = getenv("HBMPC_NO_STAGED_DIRECT");
        ctx->no_staged_direct = nd && nd[0] == '1';
        const char *sm = getenv("HBMPC_STAGED_MIN");
        if (sm) ctx->staged_min = (size_t)atoll(sm);
        const char *sg = getenv("HBMPC_STAGED_SEG");
        if (sg && atoi(sg) > 0) ctx->staged_seg = atoi(sg);
        const char *as = getenv("HBMPC_ASYNC_STAGED");
        if (as) { ctx->async_staged = as[0] != '0'; if (as[0] == '2') { ctx->attack_seen = true; ctx->force_async_staged = true; } }
        const char *nsc = getenv("HBMPC_NO_SYNC_COUNT");
        ctx->no_sync_count = nsc && nsc[0] == '1';
        const char *stl = getenv("HBMPC_STATIC_TILES");
        ctx->static_tiles = stl && stl[0] == '1';
        const char *cm = getenv("HBMPC_CHUNK_MB");
        if (cm && atoi(cm) > 0) ctx->chunk_bytes = (size_t)atoi(cm) << 20;
    }
    bool ok = true;
    for (int i = 0; i < NLANES && ok; ++i) ok = cudaStreamCreateWithFlags(&ctx->lanes[i].stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaHostAlloc((void **)&ctx->h_status, 16, cudaHostAllocMapped) == cudaSuccess &&
         cudaHostGetDevicePointer((void **)&ctx->d_status, ctx->h_status, 0) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming) == cudaSuccess &&
         cudaMallocHost((void **)&ctx->h_spec, 320 * sizeof(unsigned int)) == cudaSuccess;
    if (ok) {
        memset(ctx->h_status, 0, 16);
        cudaFuncAttributes fa;
        int regs = 0;
        if (cudaFuncGetAttributes(&fa, matvec_kernel<1>) == cudaSuccess) regs = std::max(regs, fa.numRegs);
        if (cudaFuncGetAttributes(&fa, matvec_kernel<2>) == cudaSuccess) regs = std::max(regs, fa.numRegs);
        if (cudaFuncGetAttributes(&fa, matvec_kernel<4>) == cudaSuccess) regs = std::max(regs, fa.numRegs);
        ctx->matvec_regs = regs;
        cudaFuncSetAttribute(matvec_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(matvec_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(matvec_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        ok = cudaGetLastError() == cudaSuccess && regs > 0;  // regs == 0: no sm_100a image for this device
    }
    if (!ok) {
        cudaGetLastError();
        delete ctx;
        return HBMPC_NO_DEVICE;
    }
    *out = ctx;
    return HBMPC_SUCCESS;
}

static void note_destroy_error(const char *what) {
    cudaError_t e = cudaGetLastError();  // also clears it: a failure here must not surface in another context's next call
    if (e != cudaSuccess && getenv("HBMPC_DEBUG")) fprintf(stderr, "hbmpc_ctx_destroy: %s: %s\n", what, cudaGetErrorString(e));
}

extern "C" void hbmpc_gl_release(const hbmpc_ctx *ctx);   // goldilocks.cu
extern "C" int hbmpc_ctx_device(const hbmpc_ctx *ctx) { return ctx ? ctx->device : -1; }
extern "C" void hbmpc_ctx_destroy(hbmpc_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    hbmpc_gl_release(ctx);
    for (auto &ln : ctx->lanes)
        if (ln.stream) cudaStreamSynchronize(ln.stream);
    note_destroy_error("sync");
    for (void *p : ctx->owned) cudaFree(p);
    note_destroy_error("owned");
    for (auto &e : ctx->recover)
        for (void *q : e.second.allocs) cudaFree(q);
    note_destroy_error("recover tables");
    for (auto &ln : ctx->lanes)
        for (auto &b : ln.scratch)
            if (b.p) cudaFree(b.p);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    for (int i = 0; i < MapRing::SLOTS; ++i) {
        if (ctx->maps.h[i]) cudaFreeHost(ctx->maps.h[i]);
        if (ctx->maps.d[i]) cudaFree(ctx->maps.d[i]);
        if (ctx->maps.ev[i]) cudaEventDestroy(ctx->maps.ev[i]);
    }
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    if (ctx->h_spec) cudaFreeHost(ctx->h_spec);
    note_destroy_error("host buffers");
    if (ctx->ev_main) cudaEventDestroy(ctx->ev_main);
    if (ctx->nonrobust)
        for (auto &e : *ctx->nonrobust)
            for (void *q : e.second.allocs) cudaFree(q);
    delete ctx->nonrobust;
    for (int i = 0; i < NLANES; ++i)
        if (ctx->lanes[i].stream && (i > 0 || ctx->own_stream)) cudaStreamDestroy(ctx->lanes[i].stream);
    note_destroy_error("streams");
    delete ctx;
}

extern "C" int hbmpc_ctx_set_stream(hbmpc_ctx *ctx, void *cuda_stream) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->main_stream());
    if (ctx->own_stream && ctx->main_stream()) cudaStreamDestroy(ctx->main_stream());
    ctx->lanes[0].stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return HBMPC_SUCCESS;
}

extern "C" int hbmpc_ctx_set_async(hbmpc_ctx *ctx, int async) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    ctx->async = async != 0;
    return HBMPC_SUCCESS;
}

// Device status words (mapped pinned memory): [0] non-canonical input seen, [1] densely failing batch seen by robust_kernel in
// scan mode, [2] some item failed to decode.  read_status waits for the context's stream, reads and clears them.
static int read_status(hbmpc_ctx *ctx, unsigned int &bad, unsigned int &undec) {
    CK(cudaStreamSynchronize(ctx->main_stream()));
    volatile unsigned int *hs = ctx->h_status;
    bad = hs[0];
    undec = hs[2];
    if (hs[1]) {  // the next recovery call counts its failing items (staged decoder)
        ctx->attack_seen = true;
        hs[1] = 0;
    }
    if (ctx->dev_route_calls) {   // asynchronous calls took the device-count route: none of them met a large failing set -> the attack is over
        if (!hs[3] && !ctx->force_async_staged) ctx->attack_seen = false;
        ctx->dev_route_calls = 0;
    }
    hs[3] = 0;
    hs[0] = 0;
    hs[2] = 0;
    return 0;
}
// status of the call that has just been enqueued and awaited (synchronous calls, calls with host buffers)
static int collect_status(hbmpc_ctx *ctx) {
    unsigned int bad = 0, undec = 0;
    int rc = read_status(ctx, bad, undec);
    if (rc) return rc;
    if (bad) return HBMPC_INVALID_INPUT;
    if (undec) return HBMPC_DECODING_ERROR;
    return HBMPC_SUCCESS;
}
// a call that failed half-way (CUDA error, allocation failure): nothing it left in the status words may surface in a later call
static int abandon_call(hbmpc_ctx *ctx, int rc) {
    for (auto &ln : ctx->lanes)
        if (ln.stream) cudaStreamSynchronize(ln.stream);
    cudaGetLastError();
    volatile unsigned int *hs = ctx->h_status;
    hs[0] = 0;
    hs[1] = 0;
    hs[2] = 0;
    hs[3] = 0;
    return rc;
}

extern "C" int hbmpc_ctx_synchronize(hbmpc_ctx *ctx) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    for (int i = 1; i < NLANES; ++i) CK(cudaStreamSynchronize(ctx->lanes[i].stream));   // enqueue-only host-buffer calls
    unsigned int bad = 0, undec = 0;
    int rc = read_status(ctx, bad, undec);
    if (rc) return rc;
    if (bad) return HBMPC_INVALID_INPUT;
    if (undec) return HBMPC_DECODING_ERROR;
    return HBMPC_SUCCESS;
}

extern "C" uint64_t hbmpc_ctx_launch_count(const hbmpc_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" const char *hbmpc_last_error(const hbmpc_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// ------------------------------------------------------------------------------------------------ batch buffers and chunking
// A batch buffer holds J records of `esz` bytes per item; record (b, j) lives at base + (b*sb + j*sj)*esz.
// item-major: sb = J, sj = 1 (coeffs[B][d+1], shares[B][n]);  record-major: sb = 1, sj = B (evals[S][B], out[n][B]).
struct BatchBuf {
    void *user = nullptr;
    bool host = false, present = false, record_major = false;
    long long J = 0;
    size_t esz = 32;
    size_t B = 0;
    size_t ld = 0;                           // record-major buffers: distance between records in items (B, or more when the call
                                             // works on a column range of a wider array: group calls shard the batch axis)
    const std::vector<int> *rows = nullptr;  // record-major host buffers: copy only these records (ascending); others stay stale
    void *const *ptrs = nullptr;             // record-major host buffers given as J separate arrays of B items (one message payload per
                                             // sender / recipient): record j lives at ptrs[j]; `user` is unused
};
static BatchBuf make_buf(const void *p, size_t B, long long J, bool record_major, size_t esz = 32, size_t ld = 0) {
    BatchBuf b;
    b.user = const_cast<void *>(p);
    b.present = p != nullptr;
    b.host = b.present && !is_device_ptr(p);
    b.J = J;
    b.record_major = record_major;
    b.esz = esz;
    b.B = B;
    b.ld = ld ? ld : B;
    return b;
}
// J host arrays of B items each (nullptr entries or device pointers are rejected by the callers)
static BatchBuf make_buf_ptrs(void *const *ptrs, size_t B, long long J, size_t esz = 32) {
    BatchBuf b;
    b.user = nullptr;
    b.present = ptrs != nullptr;
    b.host = true;
    b.J = J;
    b.record_major = true;
    b.esz = esz;
    b.B = B;
    b.ld = B;
    b.ptrs = ptrs;
    return b;
}
static bool host_ptr_array_ok(const void *const *ptrs, size_t J) {
    if (!ptrs) return false;
    for (size_t j = 0; j < J; ++j)
        if (!ptrs[j] || is_device_ptr(ptrs[j])) return false;
    return true;
}
// device view of the chunk [b0, b0+Bc)
struct ChunkView {
    void *dev = nullptr;
    long long sb = 0, sj = 0;
};
static int chunk_prepare(hbmpc_ctx *ctx, Lane &ln, int slot, const BatchBuf &bb, size_t b0, size_t Bc, bool copy_in, ChunkView &v) {
    if (!bb.present) return 0;
    if (!bb.host) {
        v.sb = bb.record_major ? 1 : bb.J;
        v.sj = bb.record_major ? (long long)bb.ld : 1;
        v.dev = (char *)bb.user + (size_t)b0 * (size_t)v.sb * bb.esz;
        return 0;
    }
    int rc = scratch_get(ctx, ln, slot, Bc * (size_t)bb.J * bb.esz, &v.dev);
    if (rc) return rc;
    if (bb.record_major && bb.ptrs) {   // one message payload per record: copied straight from where the messages lie
        v.sb = 1;
        v.sj = (long long)Bc;
        if (copy_in) {
            if (!bb.rows) {
                for (long long j = 0; j < bb.J; ++j)
                    CK(cudaMemcpyAsync((char *)v.dev + (size_t)j * Bc * bb.esz, (const char *)bb.ptrs[j] + b0 * bb.esz, Bc * bb.esz, cudaMemcpyHostToDevice, ln.stream));
            } else {
                for (int j : *bb.rows)
                    CK(cudaMemcpyAsync((char *)v.dev + (size_t)j * Bc * bb.esz, (const char *)bb.ptrs[j] + b0 * bb.esz, Bc * bb.esz, cudaMemcpyHostToDevice, ln.stream));
            }
        }
    } else if (bb.record_major) {
        v.sb = 1;
        v.sj = (long long)Bc;
        if (copy_in && !bb.rows)
            CK(cudaMemcpy2DAsync(v.dev, Bc * bb.esz, (char *)bb.user + b0 * bb.esz, bb.ld * bb.esz, Bc * bb.esz, (size_t)bb.J, cudaMemcpyHostToDevice, ln.stream));
        if (copy_in && bb.rows) {
            const std::vector<int> &r = *bb.rows;
            for (size_t i = 0; i < r.size();) {  // one 2D copy per run of consecutive records
                size_t k = i + 1;
                while (k < r.size() && r[k] == r[k - 1] + 1) ++k;
                CK(cudaMemcpy2DAsync((char *)v.dev + (size_t)r[i] * Bc * bb.esz, Bc * bb.esz, (char *)bb.user + ((size_t)r[i] * bb.ld + b0) * bb.esz,
                                     bb.ld * bb.esz, Bc * bb.esz, k - i, cudaMemcpyHostToDevice, ln.stream));
                i = k;
            }
        }
    } else {
        v.sb = bb.J;
        v.sj = 1;
        if (copy_in)
            CK(cudaMemcpyAsync(v.dev, (char *)bb.user + b0 * (size_t)bb.J * bb.esz, Bc * (size_t)bb.J * bb.esz, cudaMemcpyHostToDevice, ln.stream));
    }
    return 0;
}
static int chunk_commit(hbmpc_ctx *ctx, Lane &ln, const BatchBuf &bb, size_t b0, size_t Bc, const ChunkView &v) {
    if (!bb.present || !bb.host) return 0;
    if (bb.record_major && bb.ptrs) {
        for (long long j = 0; j < bb.J; ++j)
            CK(cudaMemcpyAsync((char *)bb.ptrs[j] + b0 * bb.esz, (const char *)v.dev + (size_t)j * Bc * bb.esz, Bc * bb.esz, cudaMemcpyDeviceToHost, ln.stream));
    } else if (bb.record_major)
        CK(cudaMemcpy2DAsync((char *)bb.user + b0 * bb.esz, bb.ld * bb.esz, v.dev, Bc * bb.esz, Bc * bb.esz, (size_t)bb.J, cudaMemcpyDeviceToHost, ln.stream));
    else
        CK(cudaMemcpyAsync((char *)bb.user + b0 * (size_t)bb.J * bb.esz, v.dev, Bc * (size_t)bb.J * bb.esz, cudaMemcpyDeviceToHost, ln.stream));
    return 0;
}

// Runs `body(lane, b0, Bc)` over the batch: one pass on lane 0 when every buffer is a device pointer, otherwise chunk by
// chunk round-robin over lanes 1..3 so that host->device copies, kernels and device->host copies of neighbouring chunks
// overlap.  Returns after the host buffers are complete (or, all-device in async mode, after enqueueing).
static size_t pick_chunk(const hbmpc_ctx *ctx, size_t B, size_t max_item_bytes) {
    // large chunks move at a higher PCIe rate, but a call should still be cut into enough pieces for its copies and kernels to overlap
    const size_t item = std::max<size_t>(max_item_bytes, 32);
    const size_t target = std::min<size_t>(ctx->chunk_bytes, std::max<size_t>((size_t)16 << 20, B * item / 12));
    size_t Bc = target / item;
    Bc = std::max<size_t>(Bc & ~(size_t)255, 1024);
    if (Bc >= B || B <= 4096) Bc = B;
    return Bc;
}
template <typename Body>
static int run_batched(hbmpc_ctx *ctx, size_t B, bool any_host, size_t max_item_bytes, Body body, bool collect = true) {
    if (!any_host) {
        int rc = body(ctx->lanes[0], (size_t)0, B);
        if (rc) return abandon_call(ctx, rc);
        return ctx->async ? HBMPC_SUCCESS : collect_status(ctx);
    }
    // asynchronous mode: enqueue only (the caller's buffers must stay valid until hbmpc_ctx_synchronize, which also returns the
    // status), on the lane set the previous host-buffer call did not use
    const int base = 1 + (ctx->async ? LANES_PER_SET * ctx->lane_set : 0);
    if (ctx->async) ctx->lane_set ^= 1;
    const size_t Bc = pick_chunk(ctx, B, max_item_bytes);
    CK(cudaEventRecord(ctx->ev_main, ctx->main_stream()));
    for (int i = 0; i < LANES_PER_SET; ++i) CK(cudaStreamWaitEvent(ctx->lanes[base + i].stream, ctx->ev_main, 0));
    int li = 0, rc = 0;
    for (size_t b0 = 0; b0 < B && !rc; b0 += Bc) {
        Lane &ln = ctx->lanes[base + li];
        li = (li + 1) % LANES_PER_SET;
        rc = body(ln, b0, std::min(Bc, B - b0));
    }
    if (ctx->async && !rc) return HBMPC_SUCCESS;
    for (int i = 0; i < LANES_PER_SET; ++i) {
        cudaError_t e = cudaStreamSynchronize(ctx->lanes[base + i].stream);
        if (e != cudaSuccess && !rc) {
            ctx->err = std::string("pipeline lane: ") + cudaGetErrorString(e);
            rc = HBMPC_CUDA_ERROR;
        }
    }
    if (rc) return abandon_call(ctx, rc);
    return collect ? collect_status(ctx) : HBMPC_SUCCESS;
}

// ------------------------------------------------------------------------------------------------ kernel launches
static int launch_matvec(hbmpc_ctx *ctx, Lane &ln, MatvecArgs a, int flag_words) {
    cudaStream_t st = ln.stream;
    if (a.B == 0) return 0;
    MatvecPlan p = matvec_plan(a.R, a.C, flag_words, ctx->matvec_regs > 0 ? ctx->matvec_regs : 112, a.B, ctx->num_sms);
    if (p.tbt == 0) {
        // too wide for one shared-memory tile: two column blocks into temporaries, then add + row semantics
        if (a.C < 2 || a.row_len) {
            ctx->err = "matvec: no launch shape fits shared memory";
            return HBMPC_INVALID_INPUT;
        }
        void *tmp = nullptr;
        const size_t tbytes = (size_t)a.B * a.R * 32;
        int rc = scratch_get(ctx, ln, 8, 2 * tbytes, &tmp);
        if (rc) return rc;
        uint4 *T1 = (uint4 *)tmp, *T2 = (uint4 *)((char *)tmp + tbytes);
        const int ld = a.M_ld ? a.M_ld : a.C, C1 = a.C / 2;
        for (int h = 0; h < 2; ++h) {
            MatvecArgs s = a;
            s.M = a.M + (size_t)(h ? C1 : 0) * 2;
            s.M_ld = ld;
            s.C = h ? a.C - C1 : C1;
            s.col0 = a.col0 + (h ? C1 : 0);
            s.out = h ? T2 : T1;
            s.out_sb = a.R;
            s.out_sr = 1;
            s.n_chk = 0;
            s.n_gate = 0;
            s.chk_map = nullptr;
            s.fail = nullptr;
            s.flags = nullptr;
            if ((rc = launch_matvec(ctx, ln, s, 0))) return rc;
        }
        a.flag_words = flag_words;
        matvec_combine_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(a, T1, T2);
        ctx->launches++;
        CK(cudaGetLastError());
        return 0;
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
        case 1: matvec_kernel<1><<<grid, block, p.smem, st>>>(a); break;
        case 2: matvec_kernel<2><<<grid, block, p.smem, st>>>(a); break;
        default: matvec_kernel<4><<<grid, block, p.smem, st>>>(a); break;
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

static unsigned long long *work_slot(hbmpc_ctx *ctx, cudaStream_t st);
template <int LOGN, int MODE>
static int launch_ntt_t(hbmpc_ctx *ctx, cudaStream_t st, const NttArgs &a0) {
    NttArgs a = a0;
    a.work = work_slot(ctx, st);
    const size_t smem = ntt_smem_bytes<LOGN>();
    int &ctas = ctx->ntt_ctas[MODE][LOGN];
    if (ctas == 0) {
        CK(cudaFuncSetAttribute(ntt_kernel<LOGN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ntt_kernel<LOGN, MODE>, HB_NTT_BLOCK, smem));
        ctas = nb > 0 ? nb : 1;
    }
    const int ipc = ntt_items_per_cta<LOGN>();
    long long ntiles = (a.B + ipc - 1) / ipc;
    long long grid = std::min<long long>(ntiles, (long long)ctx->num_sms * ctas);
    ntt_kernel<LOGN, MODE><<<(unsigned)grid, HB_NTT_BLOCK, smem, st>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
template <int MODE>
static int launch_ntt64_cta(hbmpc_ctx *ctx, cudaStream_t st, const NttArgs &a) {
    const size_t smem = ntt64_cta_smem_bytes();
    int &ctas = ctx->ntt64_ctas[MODE];
    if (ctas == 0) {
        CK(cudaFuncSetAttribute(ntt64_cta_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ntt64_cta_kernel<MODE>, HB_NTT64_IPC * 16, smem));
        ctas = nb > 0 ? nb : 1;
    }
    const long long ntiles = (a.B + HB_NTT64_IPC - 1) / HB_NTT64_IPC;
    const long long grid = std::min<long long>(ntiles, (long long)ctx->num_sms * ctas);
    ntt64_cta_kernel<MODE><<<(unsigned)grid, HB_NTT64_IPC * 16, smem, st>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
constexpr int WORK_SLOTS = 64;
// the tile-queue words of stream `st` (nullptr when the pool is exhausted: the kernel then walks its tiles statically)
static unsigned long long *work_slot(hbmpc_ctx *ctx, cudaStream_t st) {
    if (ctx->static_tiles) return nullptr;
    if (!ctx->work_pool) {
        void *p = nullptr;
        if (cudaMalloc(&p, WORK_SLOTS * 4 * sizeof(unsigned long long)) != cudaSuccess || cudaMemset(p, 0, WORK_SLOTS * 4 * sizeof(unsigned long long)) != cudaSuccess) {
            cudaGetLastError();
            if (p) cudaFree(p);
            ctx->static_tiles = true;
            return nullptr;
        }
        ctx->work_pool = (unsigned long long *)p;
        ctx->owned.push_back(p);
    }
    for (size_t i = 0; i < ctx->work_streams.size(); ++i)
        if (ctx->work_streams[i] == st) return ctx->work_pool + 4 * i;
    if (ctx->work_streams.size() >= (size_t)WORK_SLOTS) return nullptr;
    ctx->work_streams.push_back(st);
    return ctx->work_pool + 4 * (ctx->work_streams.size() - 1);
}
template <int LOGN, int MODE>
static int launch_ntt16x_t(hbmpc_ctx *ctx, cudaStream_t st, const NttArgs &a0) {
    NttArgs a = a0;
    a.work = work_slot(ctx, st);
    const size_t smem = ntt16x_smem_bytes<LOGN, MODE>();
    int &ctas = ctx->ntt16x_ctas[MODE][LOGN];
    if (ctas == 0) {
        CK(cudaFuncSetAttribute(ntt16x_kernel<LOGN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ntt16x_kernel<LOGN, MODE>, NTT16X_WARPS * 32, smem));
        ctas = nb > 0 ? nb : 1;
    }
    const int ipc = NTT16X_WARPS * (32 >> (LOGN - ntt16x_logs<LOGN>()));
    const long long ntiles = (a.B + ipc - 1) / ipc;
    const long long grid = std::min<long long>(ntiles, (long long)ctx->num_sms * ctas);
    ntt16x_kernel<LOGN, MODE><<<(unsigned)grid, NTT16X_WARPS * 32, smem, st>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
template <int MODE>
static int launch_ntt(hbmpc_ctx *ctx, cudaStream_t st, int logn, const NttArgs &a) {
    if constexpr (MODE <= 2) {
        // large batches of 16..128-point transforms: the barrier-free single-product-site kernel (ntt16x.cuh)
        if (logn >= 4 && logn <= 7 && a.B >= ctx->ntt16x_min && !a.item_list && !a.out_group32 && a.cols <= (1 << logn)) {
            switch (logn) {
                case 4: return launch_ntt16x_t<4, MODE>(ctx, st, a);
                case 5: return launch_ntt16x_t<5, MODE>(ctx, st, a);
                case 6: return launch_ntt16x_t<6, MODE>(ctx, st, a);
                case 7: return launch_ntt16x_t<7, MODE>(ctx, st, a);
            }
        }
    }
    if constexpr (MODE == 0 || MODE == 1) {
        // measured (profiles/r01i_*): the CTA-cooperative kernel wins where whole warps can skip products -- zero-padded inputs
        // (share generation: 22 or 43 coefficients of 64... up to half the domain) and the inverse transform; the full-width
        // forward transform (64 columns) is faster in the warp-per-item kernel
        if (logn == 6 && ctx->ntt_cta && a.B >= ctx->ntt_cta_min && (ctx->ntt_cta_all || MODE == 1 || a.cols <= 32))
            return launch_ntt64_cta<MODE>(ctx, st, a);
    }
    switch (logn) {
        case 1: return launch_ntt_t<1, MODE>(ctx, st, a);
        case 2: return launch_ntt_t<2, MODE>(ctx, st, a);
        case 3: return launch_ntt_t<3, MODE>(ctx, st, a);
        case 4: return launch_ntt_t<4, MODE>(ctx, st, a);
        case 5: return launch_ntt_t<5, MODE>(ctx, st, a);
        case 6: return launch_ntt_t<6, MODE>(ctx, st, a);
        case 7: return launch_ntt_t<7, MODE>(ctx, st, a);
        case 8: return launch_ntt_t<8, MODE>(ctx, st, a);
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

// inverse transform tables: w_N^{-k} (k < N/2) followed by N^{-1}, all in Montgomery form
static int get_inverse_twiddles(hbmpc_ctx *ctx, int N, uint4 **tw, uint4 **scale) {
    char key[64];
    snprintf(key, sizeof key, "WI %d", N);
    auto it = ctx->matrices.find(key);
    if (it == ctx->matrices.end()) {
        std::vector<HFr> d = domain_elements((size_t)N, (size_t)N);
        std::vector<HFr> t((size_t)std::max(N / 2, 1) + 1);
        t[0] = hfr::ONE;
        for (int k = 1; k < N / 2; ++k) t[k] = d[N - k];  // w^{-k} = w^{N-k}
        t[std::max(N / 2, 1)] = hfr::inv(hfr::from_u64((uint64_t)N));
        uint4 *dev = nullptr;
        int rc = upload_fr(ctx, t, &dev);
        if (rc) return rc;
        it = ctx->matrices.emplace(key, dev).first;
    }
    *tw = it->second;
    *scale = it->second + (size_t)std::max(N / 2, 1) * 2;
    return 0;
}

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

// ------------------------------------------------------------------------------------------------ K1 / K2
// out[b][r] = sum_c M[r][c] in[b][c]: M == nullptr selects the domain transform (NTT) with `n` outputs
static int apply_map(hbmpc_ctx *ctx, const uint4 *M, const uint4 *tw, int logn, size_t rows, size_t cols, size_t B, const uint64_t *in,
                     uint64_t *out, int recipient_major, size_t ld_out = 0, void *const *out_ptrs = nullptr) {
    BatchBuf bi = make_buf(in, B, (long long)cols, false);
    BatchBuf bo = out_ptrs ? make_buf_ptrs(out_ptrs, B, (long long)rows) : make_buf(out, B, (long long)rows, recipient_major != 0, 32, ld_out);
    auto body = [&](Lane &ln, size_t b0, size_t Bc) -> int {
        ChunkView vi, vo;
        int rc;
        if ((rc = chunk_prepare(ctx, ln, 0, bi, b0, Bc, true, vi))) return rc;
        if ((rc = chunk_prepare(ctx, ln, 1, bo, b0, Bc, false, vo))) return rc;
        if (M) {
            MatvecArgs a{};
            a.M = M;
            a.in = (const uint4 *)vi.dev;
            a.out = (uint4 *)vo.dev;
            a.R = (int)rows;
            a.C = (int)cols;
            a.B = (long long)Bc;
            a.in_sb = vi.sb; a.in_sc = vi.sj; a.in_chunk_major = 1;
            a.out_sb = vo.sb; a.out_sr = vo.sj;
            if ((rc = launch_matvec(ctx, ln, a, 0))) return rc;
        } else {
            NttArgs a{};
            a.in = (const uint4 *)vi.dev;
            a.out = (uint4 *)vo.dev;
            a.tw = tw;
            a.B = (long long)Bc;
            a.in_sb = vi.sb; a.in_sc = vi.sj;
            a.out_sb = vo.sb; a.out_sr = vo.sj;
            a.cols = (int)cols;
            a.n = (int)rows;
            a.err = ctx->d_status;
            if ((rc = launch_ntt<0>(ctx, ln.stream, logn, a))) return rc;
        }
        return chunk_commit(ctx, ln, bo, b0, Bc, vo);
    };
    return run_batched(ctx, B, bi.host || bo.host, std::max(rows, cols) * 32, body);
}

static int apply_domain(hbmpc_ctx *ctx, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *out, int recipient_major, size_t ld_out = 0,
                        void *const *out_ptrs = nullptr) {
    const int N = domain_size(n);
    int logn = 0;
    while ((1 << logn) < N) ++logn;
    if (ctx->force_dense || N < 2 || cols > (size_t)N) {
        uint4 *V = nullptr;
        int rc = get_vandermonde(ctx, n, cols, &V);
        if (rc) return rc;
        return apply_map(ctx, V, nullptr, 0, n, cols, B, in, out, recipient_major, ld_out, out_ptrs);
    }
    uint4 *tw = nullptr;
    int rc = get_twiddles(ctx, N, &tw);
    if (rc) return rc;
    return apply_map(ctx, nullptr, tw, logn, n, cols, B, in, out, recipient_major, ld_out, out_ptrs);
}

extern "C" int hbmpc_compute_shares_batch(hbmpc_ctx *ctx, size_t n, size_t d, size_t B, const uint64_t *coeffs, uint64_t *shares) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (n <= d) return HBMPC_INVALID_INPUT;                 // robust_interpolate.rs:59-64
    if (!domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;   // :65-66
    if (B == 0) return HBMPC_SUCCESS;
    if (!coeffs || !shares) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    return apply_domain(ctx, n, d + 1, B, coeffs, shares, 0);
}

extern "C" int hbmpc_apply_vandermonde_batch(hbmpc_ctx *ctx, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *out,
                                             int recipient_major) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (cols == 0 || cols > 256) return HBMPC_INVALID_INPUT;  // row length must equal shares.len() (share/mod.rs:54-60)
    if (!domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;
    if (B == 0) return HBMPC_SUCCESS;
    if (!in || !out) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    return apply_domain(ctx, n, cols, B, in, out, recipient_major);
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
    int rc = scratch_get(ctx, ctx->lanes[0], 9, w.size() * 4, &dM);
    if (rc) return rc;
    // stream-ordered after earlier calls that may still read the old matrix; w is pageable and freed on return, so wait for the copy
    CK(cudaMemcpyAsync(dM, w.data(), w.size() * 4, cudaMemcpyHostToDevice, ctx->main_stream()));
    CK(cudaStreamSynchronize(ctx->main_stream()));
    return apply_map(ctx, (const uint4 *)dM, nullptr, 0, rows, cols, B, in, out, recipient_major);
}

// ------------------------------------------------------------------------------------------------ recovery tables
static int build_recover_tables(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S,
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
    int rc;
    if ((rc = upload_fr(ctx, M, &T.M, &T.allocs))) return rc;

    // robust tables
    T.rmax = (int)std::min(t, S - needed);
    T.fast = T.rmax >= 1 ? 1 : 0;
    const int natt = 1 + T.rmax;
    std::vector<int> att_P(natt), att_nsyn(natt), att_maxL(natt);
    std::vector<long long> att_uoff(natt);
    std::vector<HFr> U2, U;
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
        att_uoff[a] = (long long)U.size();
        std::vector<HFr> u(uinv.begin(), uinv.begin() + P);
        HFr one_c = {{1, 0, 0, 0}};
        for (size_t i = 0; i < P; ++i) U.push_back(hfr::mul(u[i], one_c));  // canonical
        hfr::batch_inv(u);
        for (size_t i = 0; i < P; ++i) U2.push_back(hfr::mul(u[i], hfr::R2));  // u_i * R^2
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
        att_P[0] = (int)S; att_nsyn[0] = 0; att_maxL[0] = 0; att_uoff[0] = 0;
    }
    std::vector<HFr> xinv(xs);
    hfr::batch_inv(xinv);
    std::vector<HFr> Veval = vandermonde_on_points(xs, m);
    if ((rc = upload(ctx, att_P, &T.att_P, &T.allocs))) return rc;
    if ((rc = upload(ctx, att_nsyn, &T.att_nsyn, &T.allocs))) return rc;
    if ((rc = upload(ctx, att_maxL, &T.att_maxL, &T.allocs))) return rc;
    if ((rc = upload(ctx, att_uoff, &T.att_uoff, &T.allocs))) return rc;
    if ((rc = upload_fr(ctx, U2, &T.u2, &T.allocs))) return rc;
    {
        std::vector<int> sid(S);
        for (size_t i = 0; i < S; ++i) sid[i] = (int)sorted_ids[i];
        if ((rc = upload(ctx, sid, &T.sid, &T.allocs))) return rc;
        uint4 *sc = nullptr;
        if ((rc = get_twiddles(ctx, domain_size(n), &T.tw))) return rc;
        if ((rc = get_inverse_twiddles(ctx, domain_size(n), &T.ritw, &sc))) return rc;
    }
    if ((rc = upload_fr(ctx, U, &T.uinv, &T.allocs))) return rc;
    if (T.fast) {
        const int Nd = domain_size(n);
        std::vector<HFr> swt((size_t)Nd, hfr::ZERO);
        std::vector<int> pod((size_t)Nd, -1);
        for (size_t i = 0; i < S; ++i) {
            swt[sorted_ids[i]] = U2[(size_t)att_uoff[0] + i];
            pod[sorted_ids[i]] = (int)i;
        }
        if ((rc = upload_fr(ctx, swt, &T.syn_wt, &T.allocs))) return rc;
        if ((rc = upload(ctx, pod, &T.pos_of_dom, &T.allocs))) return rc;
        T.nsyn0 = att_nsyn[0];
        T.maxL0 = att_maxL[0];
        T.uoff0 = att_uoff[0];
    }
    if ((rc = upload_fr(ctx, xs, &T.xs, &T.allocs))) return rc;
    if ((rc = upload_fr(ctx, xinv, &T.xinv, &T.allocs))) return rc;
    if ((rc = upload_fr(ctx, L.Lc, &T.Lc, &T.allocs))) return rc;
    if ((rc = upload_fr(ctx, Veval, &T.Veval, &T.allocs))) return rc;
    // every point of the power-of-two domain supplied: coefficients by one inverse NTT, checked by "top coefficients vanish"
    const int N = domain_size(n);
    if (!ctx->no_fastpath && S == n && (size_t)N == n && N >= 2) {
        if ((rc = get_inverse_twiddles(ctx, N, &T.itw, &T.iscale))) return rc;
        while ((1 << T.fast_logn) < N) ++T.fast_logn;
    }
    // Erasure-weighted transform for the general case (any id subset, no flags): with X = the d+t+1 examined ids and
    // Zc(x) = prod_{k < N, k not in X} (x - w^k), the word y'_k = y_k*Zc(w^k) (k in X, zero elsewhere) is the evaluation
    // vector of Q = P_X*Zc where P_X interpolates the examined shares; deg P_X <= d  <=>  the coefficients N-t .. N-1 of
    // Q = INTT(y') vanish, and then P = Q*(N*Zc)^{-1} mod x^(d+1): a (d+1) x (d+1) triangular matrix.
    // With flags the same transform runs over ALL S supplied ids (Zc over the ids that are absent): a chunk in which every supplied
    // share lies on one degree-d polynomial has path 0 and no flag set; the others go to the dense check, which examines the prefix.
    if (!ctx->no_fastpath && T.fast_logn == 0 && N >= 2 && !(want_flags && ctx->no_er_flags)) {
        const size_t xcount = want_flags ? S : needed;
        T.er_all = want_flags;
        std::vector<char> inX(N, 0);
        for (size_t i = 0; i < xcount; ++i) inX[sorted_ids[i]] = 1;
        std::vector<HFr> domN = domain_elements((size_t)N, (size_t)N);
        std::vector<HFr> Z(1, hfr::ONE);  // coefficients of Zc, low degree first
        for (int k = 0; k < N; ++k) {
            if (inX[k]) continue;
            HFr nx = hfr::neg(domN[k]);
            Z.push_back(hfr::ZERO);
            for (size_t i = Z.size() - 1; i >= 1; --i) Z[i] = hfr::add(Z[i - 1], hfr::mul(Z[i], nx));
            Z[0] = hfr::mul(Z[0], nx);
        }
        std::vector<HFr> wt(N, hfr::ZERO);
        for (int k = 0; k < N; ++k) {
            if (!inX[k]) continue;
            HFr acc = hfr::ZERO;
            for (size_t i = Z.size(); i-- > 0;) acc = hfr::add(hfr::mul(acc, domN[k]), Z[i]);
            wt[k] = acc;
        }
        // P = Q / (N*Zc), from BOTH ends of Q (Q = P*N*Zc is exact when the check passes, deg Q <= d + zdeg, zdeg = deg Zc):
        //   low half   p_k     = sum_{i<=k} q_i          * Wlo[k-i],  k < h1,   Wlo = (N*Zc)^{-1}      mod x^h1
        //   high half  p_{d-i} = sum_{j<=i} q_{d+zdeg-j} * Whi[i-j],  i < h2,   Whi = rev(N*Zc)^{-1}   mod x^h2
        // two triangles of h1 and h2 = m - h1 rows (132 terms at m = 22) instead of one of m rows (253).  The transform stores
        // q_0 .. q_{h1-1} and then q_{d+zdeg}, q_{d+zdeg-1}, ... (NttArgs::hi_top / hi_cnt); row k of the matrix reads its
        // row_len[k] columns from column row_start[k].  Secrets only (mout = 1): one product with Wlo[0].
        HFr Nf = hfr::from_u64((uint64_t)N);
        const size_t zdeg = Z.size() - 1;
        const size_t h1 = (mout + 1) / 2, h2 = mout - h1;
        auto series_inverse = [&](const std::vector<HFr> &A, size_t terms) {   // (sum A_i x^i)^{-1} mod x^terms
            std::vector<HFr> W(terms, hfr::ZERO);
            if (!terms) return W;
            W[0] = hfr::inv(A[0]);
            const HFr nW0 = hfr::neg(W[0]);
            for (size_t k = 1; k < terms; ++k) {
                HFr acc = hfr::ZERO;
                for (size_t i = 1; i <= k && i < A.size(); ++i) acc = hfr::add(acc, hfr::mul(A[i], W[k - i]));
                W[k] = hfr::mul(nW0, acc);
            }
            return W;
        };
        std::vector<HFr> Zs(Z.size()), Zr(Z.size());
        for (size_t i = 0; i <= zdeg; ++i) Zs[i] = hfr::mul(Z[i], Nf);
        for (size_t i = 0; i <= zdeg; ++i) Zr[i] = Zs[zdeg - i];
        const std::vector<HFr> Wlo = series_inverse(Zs, h1), Whi = series_inverse(Zr, h2);
        std::vector<HFr> tri(mout * mout, hfr::ZERO);
        std::vector<int> row_len(mout), row_start(mout);
        for (size_t k = 0; k < h1; ++k) {
            row_start[k] = 0;
            row_len[k] = (int)k + 1;
            for (size_t i = 0; i <= k; ++i) tri[k * mout + i] = Wlo[k - i];
        }
        for (size_t k = h1; k < mout; ++k) {   // k = d - i
            const size_t i = (mout - 1) - k;
            row_start[k] = (int)h1;
            row_len[k] = (int)i + 1;
            for (size_t j = 0; j <= i; ++j) tri[k * mout + h1 + j] = Whi[i - j];
        }
        if ((rc = upload(ctx, row_len, &T.er_row_len, &T.allocs))) return rc;
        if ((rc = upload(ctx, row_start, &T.er_row_start, &T.allocs))) return rc;
        if ((rc = upload_fr(ctx, wt, &T.er_wt, &T.allocs))) return rc;
        if ((rc = upload_fr(ctx, tri, &T.er_tri, &T.allocs))) return rc;
        T.er_h1 = (int)h1;
        T.er_h2 = (int)h2;
        T.er_hi_top = (int)((mout - 1) + zdeg);   // mout = d + 1 whenever h2 > 0
        if (!T.itw && (rc = get_inverse_twiddles(ctx, N, &T.itw, &T.iscale))) return rc;
        T.er_zero_from = N - (int)(xcount - m);  // deg Q <= d + N - |X|
        while ((1 << T.er_logn) < N) ++T.er_logn;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ K4 staged decoder
// Decodes the failing items list[first .. cnt) (cnt known on the host) wave by wave through the stages described in
// robust.cuh; what the fast attempt cannot decode is collected in a second list and finished by robust_kernel's exact path.
// direct != nullptr: the items come straight from the all-points NTT check (S == n == N); paths include 0, the coefficients
// (INTT of the received word, left in `coeffs` by that check) are corrected by one sparse inverse transform per item, and the
// leftovers are returned in direct[0] (list) / direct[1] (count) for the caller's dense check instead of being decoded here.
// bytes of wave workspace per slot / slots per wave
static size_t staged_slot_bytes(const RecoverTables &T, int t) {
    const size_t tp = (size_t)t + 2, syn_ld = (size_t)std::max(T.nsyn0, 1);
    return (2 * (syn_ld + 2 * tp + 1) + 5 * tp + 1) * 32 + 2 * 16 + 2 * 4 + 1 + 16 + 32 + 1 + 4 + 1;
}
static size_t staged_wave_slots(const hbmpc_ctx *ctx, const RecoverTables &T, int t) {
    size_t budget = (size_t)6144 << 20;  // of 180 GB
    if (const char *wm = getenv("HBMPC_STAGED_WS_MB")) if (atoll(wm) > 0) budget = (size_t)atoll(wm) << 20;
    (void)ctx;
    return std::max<size_t>(budget / staged_slot_bytes(T, t), 1024) & ~(size_t)1023;
}
// list2_cap: capacity of the leftover list (0: the number of items of this call); append: keep the leftovers collected by an
// earlier call of the same capacity; hist_slots: r.hist samples the first hist_slots slots of every wave.
static int staged_decode(hbmpc_ctx *ctx, Lane &ln, const RecoverTables &T, const RobustArgs &r_in, const int *in_map, unsigned int first,
                         unsigned int cnt, long long fb_blocks, int fb_threads, unsigned int **direct = nullptr, unsigned int list2_cap = 0,
                         bool append = false, unsigned int hist_slots = 0, const unsigned int *cnt_dev = nullptr) {
    // cnt_dev != nullptr (device-count mode, asynchronous calls): `cnt` is only an upper bound (one wave: the caller checks that it
    // fits), the number of items is read on the device by every stage; slots beyond it are dead from the start
    if (cnt <= first) return 0;
    RobustArgs r = r_in;
    r.skip_coeffs = direct ? 1 : 0;
    cudaStream_t st = ln.stream;
    const int logn = r.logn, N = 1 << logn;
    const int tp = r.t + 2, syn_ld = std::max(T.nsyn0, 1), seg = ctx->staged_seg;
    const int nseg = (T.nsyn0 + seg - 1) / seg;
    const size_t total = cnt - first;
    if (list2_cap < total) list2_cap = (unsigned int)total;
    size_t Wmax = staged_wave_slots(ctx, T, r.t);
    if (Wmax > total) Wmax = total;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = al(off + bytes); return o; };
    size_t o_synG[2], o_lamG[2], o_bpG[2], o_bdisP[2], o_stateP[2], o_originP[2];
    size_t o_keyP, o_lam, o_bp, o_om, o_num, o_den, o_state, o_mask, o_perm, o_key, o_runs, o_okf;
    void *wsp = nullptr, *aux = nullptr;
    int rc;
    for (;;) {  // the wave shrinks when the device cannot spare the workspace
        const size_t Wg = (Wmax + 31) / 32 * 32;  // group-interleaved arrays hold whole groups of 32 positions
        off = 0;
        for (int i = 0; i < 2; ++i) {
            o_synG[i] = take(Wg * syn_ld * 32);
            o_lamG[i] = take(Wg * tp * 32);
            o_bpG[i] = take(Wg * tp * 32);
            o_bdisP[i] = take(Wg * 32);
            o_stateP[i] = take(Wg * 16);
            o_originP[i] = take(Wg * 4);
        }
        o_keyP = take(Wg);
        o_lam = take(Wmax * tp * 32); o_bp = take(Wmax * tp * 32); o_om = take(Wmax * tp * 32);
        o_num = take(Wmax * tp * 32); o_den = take(Wmax * tp * 32);
        o_state = take(Wmax * 16); o_mask = take(Wmax * 32); o_perm = take(Wmax * 4); o_key = take(Wmax);
        o_runs = take(Wmax * 32); o_okf = take(Wmax);
        rc = scratch_get(ctx, ln, 10, off, &wsp);
        if (!rc) break;
        cudaGetLastError();
        if (Wmax <= 4096) return rc;
        Wmax = std::max<size_t>((Wmax / 2) & ~(size_t)1023, 1024);
    }
    const size_t hist_bytes = (size_t)(nseg + 1) * 256 * 4;
    if ((rc = scratch_get(ctx, ln, 11, al(hist_bytes) + (size_t)list2_cap * 4 + 256, &aux))) return rc;
    unsigned int *hist = (unsigned int *)aux;
    unsigned int *count2 = (unsigned int *)((char *)aux + al(hist_bytes));
    unsigned int *list2 = count2 + 4;
    if (!append) CK(cudaMemsetAsync(count2, 0, 16, st));
    char *w8 = (char *)wsp;
    StagedArgs sa{};
    sa.lam = (uint4 *)(w8 + o_lam); sa.bp = (uint4 *)(w8 + o_bp); sa.om = (uint4 *)(w8 + o_om);
    sa.num = (uint4 *)(w8 + o_num); sa.den = (uint4 *)(w8 + o_den); sa.state = (int4 *)(w8 + o_state);
    sa.rootmask = (unsigned int *)(w8 + o_mask); sa.key = (unsigned char *)(w8 + o_key); sa.keyP = (unsigned char *)(w8 + o_keyP);
    unsigned int *perm = (unsigned int *)(w8 + o_perm);
    sa.syn_ld = syn_ld; sa.tp = tp; sa.nsyn = T.nsyn0; sa.maxL = T.maxL0;
    sa.pos_of_dom = T.pos_of_dom;
    sa.uinv0 = T.uinv + T.uoff0 * 2;
    sa.list2 = list2; sa.count2 = count2;
    sa.direct = direct ? 1 : 0;
    sa.cnt_dev = cnt_dev;
    if (cnt_dev) {   // every wave of this mode re-raises the "attack continues" word while its failing set stays large
        sa.attack_flag = ctx->d_status + 3;
        sa.attack_min = (unsigned int)std::max<size_t>(ctx->staged_min / 2, 1);
    }
    sa.hist_slots = r.hist ? (hist_slots ? hist_slots : 0xffffffffu) : 0u;
    sa.runs = (uint4 *)(w8 + o_runs); sa.okf = (unsigned char *)(w8 + o_okf);
    // resident CTAs of the Berlekamp-Massey kernel are capped through its dynamic shared memory size: the live state of the
    // resident positions should stay inside the L2 (HBMPC_BM_CTAS per SM, default 5 = the register limit)
    size_t bm_smem = 0;
    if (const char *bc = getenv("HBMPC_BM_CTAS")) {
        const int c = atoi(bc);
        if (c >= 1 && c < HB_BM_MINB) {
            bm_smem = (size_t)(220 * 1024 / c - 2048) & ~(size_t)1023;
            CK(cudaFuncSetAttribute(bm_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bm_smem));
        }
    }
    const bool prof = getenv("HBMPC_STAGED_PROF") != nullptr;
    cudaEvent_t pe[8] = {};
    float pms[7] = {};
    if (prof) for (auto &e : pe) cudaEventCreate(&e);
    auto sort_by = [&](const unsigned char *key, unsigned int W, unsigned int *h) -> int {
        sort_hist_kernel<<<std::min<unsigned int>((W + 1023) / 1024, (unsigned int)ctx->num_sms * 2), 256, 0, st>>>(key, W, h);
        sort_scan_kernel<<<1, 32, 0, st>>>(h);
        sort_scatter_kernel<<<std::min<unsigned int>((W + 2047) / 2048, (unsigned int)ctx->num_sms * 2), 256, 0, st>>>(key, W, h, perm);
        ctx->launches += 3;
        CK(cudaGetLastError());
        return 0;
    };
    for (size_t w0 = first; w0 < cnt; w0 += Wmax) {
        const unsigned int W = (unsigned int)std::min<size_t>(Wmax, cnt - w0);
        const unsigned int gb = (W + 127) / 128;
        sa.W = W;
        sa.list_first = (unsigned int)w0;
        for (int i = 0; i < 2; ++i) {
            sa.synG[i] = (uint4 *)(w8 + o_synG[i]); sa.lamG[i] = (uint4 *)(w8 + o_lamG[i]); sa.bpG[i] = (uint4 *)(w8 + o_bpG[i]);
            sa.bdisP[i] = (uint4 *)(w8 + o_bdisP[i]); sa.stateP[i] = (int4 *)(w8 + o_stateP[i]); sa.originP[i] = (unsigned int *)(w8 + o_originP[i]);
        }
        CK(cudaMemsetAsync(sa.rootmask, 0, (size_t)W * 32, st));
        CK(cudaMemsetAsync(hist, 0, hist_bytes, st));
        if (prof) cudaEventRecord(pe[0], st);
        {   // 1. syndromes of the weighted word (all S shares), written by position (= slot at this point)
            NttArgs na{};
            na.in = r.in; na.in_sb = r.in_sb; na.in_sc = r.in_sc;
            na.out = sa.synG[0]; na.out_sb = syn_ld; na.out_sr = 1; na.out_group32 = 1;
            na.tw = T.tw;
            na.B = (long long)W;
            na.cols = N; na.n = N;
            na.err = ctx->d_status;
            na.in_map = in_map;
            na.wt = T.syn_wt;
            na.m = N + 1; na.mout = T.nsyn0;
            na.item_list = r.list + w0;
            na.b_dev = cnt_dev; na.b_first = (unsigned int)w0;
            if ((rc = launch_ntt<2>(ctx, st, logn, na))) return rc;
        }
        if (prof) cudaEventRecord(pe[1], st);
        // 2. Berlekamp-Massey in segments; positions re-sorted by locator degree (and their state moved) in between
        for (int k = 0; k < nseg; ++k) {
            sa.j0 = k * seg;
            sa.j1 = std::min(T.nsyn0, sa.j0 + seg);
            bm_segment_kernel<<<gb, 128, bm_smem, st>>>(sa);
            ctx->launches++;
            CK(cudaGetLastError());
            if (k + 1 < nseg) {
                if ((rc = sort_by(sa.keyP, W, hist + (size_t)k * 256))) return rc;
                sa.perm = perm;
                permute_kernel<<<(W + 255) / 256, 256, 0, st>>>(sa);
                ctx->launches++;
                CK(cudaGetLastError());
                std::swap(sa.synG[0], sa.synG[1]); std::swap(sa.lamG[0], sa.lamG[1]); std::swap(sa.bpG[0], sa.bpG[1]);
                std::swap(sa.bdisP[0], sa.bdisP[1]); std::swap(sa.stateP[0], sa.stateP[1]); std::swap(sa.originP[0], sa.originP[1]);
            }
        }
        if (prof) cudaEventRecord(pe[2], st);
        // 3. Omega, Lambda' and zero padding, per slot; final order of the slots by locator degree
        omega_kernel<<<gb, 128, 0, st>>>(sa);
        ctx->launches++;
        CK(cudaGetLastError());
        if ((rc = sort_by(sa.key, W, hist + (size_t)nseg * 256))) return rc;
        sa.perm = perm;
        if (prof) cudaEventRecord(pe[3], st);
        NttArgs nb{};
        nb.in_sb = tp; nb.in_sc = 1;
        nb.out_sb = tp; nb.out_sr = 1;
        nb.tw = T.ritw;
        nb.B = (long long)W;
        nb.cols = tp; nb.n = N;
        nb.err = ctx->d_status;
        nb.rootmask = sa.rootmask;
        nb.idset = in_map;
        nb.b_dev = cnt_dev; nb.b_first = (unsigned int)w0;
        nb.in = sa.lam;   // 4. Chien search
        if ((rc = launch_ntt<3>(ctx, st, logn, nb))) return rc;
        if (prof) cudaEventRecord(pe[4], st);
        nb.in = sa.om; nb.out = sa.num;   // 5. Forney numerators / denominators at the roots
        if ((rc = launch_ntt<4>(ctx, st, logn, nb))) return rc;
        nb.in = sa.bp; nb.out = sa.den;
        if ((rc = launch_ntt<4>(ctx, st, logn, nb))) return rc;
        if (prof) cudaEventRecord(pe[5], st);
        // 6. error values, path, corrected coefficients, flags
        staged_prefix_kernel<<<gb, 128, 0, st>>>(r, sa);
        staged_invert_kernel<<<(W + 128 * HB_INV_BATCH - 1) / (128 * HB_INV_BATCH), 128, 0, st>>>(sa);
        staged_finish_kernel<<<gb, 128, 0, st>>>(r, sa);
        ctx->launches += 3;
        CK(cudaGetLastError());
        if (direct && !r.hist_only) {  // coefficients: INTT(y) - INTT(e)
            NttArgs nc{};
            nc.in = sa.num; nc.in_sb = tp; nc.in_sc = 1;
            nc.out = r.coeffs; nc.out_sb = r.mout; nc.out_sr = 1;
            nc.tw = T.itw;
            nc.scale = T.iscale;
            nc.B = (long long)W;
            nc.cols = N; nc.n = N;
            nc.err = ctx->d_status;
            nc.mout = r.mout;
            nc.rootmask = sa.rootmask;
            nc.item_list = r.list + w0;
            nc.b_dev = cnt_dev; nc.b_first = (unsigned int)w0;
            if ((rc = launch_ntt<5>(ctx, st, logn, nc))) return rc;
        }
        if (prof) {
            cudaEventRecord(pe[6], st);
            cudaEventSynchronize(pe[6]);
            for (int i = 0; i < 6; ++i) { float ms = 0; cudaEventElapsedTime(&ms, pe[i], pe[i + 1]); pms[i] += ms; }
        }
    }
    if (prof) {
        fprintf(stderr, "staged_decode items=%zu: syndromes %.3f ms, berlekamp-massey+sorts %.3f, omega+sort %.3f, chien %.3f, forney %.3f, finish %.3f\n", total,
                pms[0], pms[1], pms[2], pms[3], pms[4], pms[5]);
        for (auto &e : pe) cudaEventDestroy(e);
    }
    if (direct) {
        direct[0] = list2;
        direct[1] = count2;
        return 0;
    }
    // leftovers: the literal OEC rounds
    RobustArgs rf = r;
    rf.list = list2;
    rf.count = count2;
    rf.fail_scan = nullptr;
    rf.fast = 0;
    rf.list_first = 0;
    rf.list_max = 0;
    robust_kernel<<<(unsigned)fb_blocks, fb_threads, 0, st>>>(rf);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ K3 / K4
// shared implementation: element (item b, arrival j) of `in` is sender-major [S][B] (K3) or codeword-major [B][S] (K4)
static int recover_impl(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *ids, size_t B, const uint64_t *in,
                        bool sender_major, uint64_t *coeffs, bool secrets_only, uint64_t *secrets, int32_t *path, uint64_t *flags,
                        size_t ld_in = 0, void *const *in_ptrs = nullptr) {
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
    if ((!in && !in_ptrs) || !path || (!coeffs && !secrets_only) || (secrets_only && !secrets)) return HBMPC_INVALID_INPUT;
    if (in_ptrs && (!sender_major || !host_ptr_array_ok((const void *const *)in_ptrs, S))) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);

    const bool want_flags = flags != nullptr;
    TableKey key;
    key.kind = (want_flags ? 1u : 0u) | (secrets_only ? 2u : 0u);
    key.n = (uint32_t)n; key.d = (uint32_t)d; key.t = (uint32_t)t;
    for (size_t i = 0; i < S; ++i) key.idset[ids[i] >> 6] |= 1ull << (ids[i] & 63);  // the field tables depend on the id SET only
    auto it = ctx->recover.find(key);
    if (it == ctx->recover.end()) {
        if (ctx->recover.size() >= 256) {  // bounded cache: sender sets differ from session to session
            for (auto &ln : ctx->lanes) CK(cudaStreamSynchronize(ln.stream));
            for (auto &e : ctx->recover)
                for (void *q : e.second.allocs) cudaFree(q);
            ctx->recover.clear();
        }
        RecoverTables T;
        int rc = build_recover_tables(ctx, n, d, t, S, sorted_ids, want_flags, secrets_only, T);
        if (rc) {
            for (void *q : T.allocs) cudaFree(q);
            return abandon_call(ctx, rc);
        }
        it = ctx->recover.emplace(key, T).first;
    }
    const RecoverTables &T = it->second;
    // arrival-order index maps, rebuilt per call (messages arrive in a different order every session): a few hundred ints
    struct { const int *order, *col_map, *chk_map, *in_map, *er_in_map; } P{};
    {
        const int N = domain_size(n);
        const size_t off_in = S, off_er = S + (size_t)N;
        const int *base = nullptr;
        int rc = maps_upload(ctx, [&](int *blk) -> size_t {
            for (size_t i = 0; i < S; ++i) blk[i] = order[i];                  // [0, S): order; col_map = its first m entries
            int *in_map = blk + off_in, *er_map = blk + off_er;
            for (int k = 0; k < 2 * N; ++k) in_map[k] = -1;
            for (size_t i = 0; i < S; ++i) in_map[sorted_ids[i]] = order[i];
            for (size_t i = 0; i < needed; ++i) er_map[sorted_ids[i]] = order[i];
            return S + 2 * (size_t)N;
        }, &base);
        if (rc) return abandon_call(ctx, rc);
        P.order = base;
        P.col_map = base;            // column c of the optimistic matrix reads the share of sorted position c
        P.chk_map = base + m;        // check row r compares with the share of sorted position m + r
        P.in_map = base + off_in;
        P.er_in_map = base + off_er;
    }
    const int fw = want_flags ? (int)((S + 63) / 64) : 0;
    const bool want_secrets = !secrets_only && secrets != nullptr;

    BatchBuf bi = in_ptrs ? make_buf_ptrs(in_ptrs, B, (long long)S) : make_buf(in, B, (long long)S, sender_major, 32, sender_major ? ld_in : 0);
    BatchBuf bc = make_buf(secrets_only ? secrets : coeffs, B, T.mout, false);
    BatchBuf bp = make_buf(path, B, 1, false, 4);
    BatchBuf bf = make_buf(flags, B, fw > 0 ? fw : 1, false, 8);
    BatchBuf bs = make_buf(want_secrets ? secrets : nullptr, B, 1, false);
    const WsLayout lay(T.nsyn_max, (int)t, domain_size(n));

    // Lean host path: when the shares arrive in host memory sender-major and no flags are wanted, only the d+t+1 examined
    // sender vectors are uploaded and the optimistic check runs on them (PCIe, not the SMs, bounds such calls); chunks in
    // which some item fails are re-run afterwards with every sender vector uploaded (robust decoding needs them all).
    std::vector<int> lean_rows;
    BatchBuf bi_lean = bi;
    const bool lean = bi.host && sender_major && !want_flags && S > needed && !ctx->async;   // (needs a host decision between its two phases)
    size_t chunk_full = 0;
    if (lean) {
        for (size_t i = 0; i < needed; ++i) lean_rows.push_back(order[i]);
        std::sort(lean_rows.begin(), lean_rows.end());
        bi_lean.rows = &lean_rows;
        chunk_full = pick_chunk(ctx, B, std::max<size_t>(S, T.mout) * 32);
        const size_t nchunks = (B + chunk_full - 1) / chunk_full;
        if (ctx->h_counts_cap < nchunks) {
            if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
            ctx->h_counts = nullptr;
            CK(cudaMallocHost((void **)&ctx->h_counts, (nchunks + 64) * sizeof(unsigned int)));
            ctx->h_counts_cap = nchunks + 64;
        }
    }
    bool lean_phase = lean;
    auto body = [&](Lane &ln, size_t b0, size_t Bc) -> int {
        ChunkView vi, vc, vp, vf, vs;
        int rc;
        if ((rc = chunk_prepare(ctx, ln, 0, lean_phase ? bi_lean : bi, b0, Bc, true, vi))) return rc;
        if ((rc = chunk_prepare(ctx, ln, 1, bc, b0, Bc, false, vc))) return rc;
        if ((rc = chunk_prepare(ctx, ln, 2, bp, b0, Bc, false, vp))) return rc;
        if (want_flags && (rc = chunk_prepare(ctx, ln, 3, bf, b0, Bc, false, vf))) return rc;
        if (want_secrets && (rc = chunk_prepare(ctx, ln, 4, bs, b0, Bc, false, vs))) return rc;
        // scratch: fail bytes + list + counter
        void *aux = nullptr;
        const size_t fail_bytes = ((Bc + 15) / 16) * 16;
        if ((rc = scratch_get(ctx, ln, 5, 2 * (fail_bytes + Bc * 4 + 16), &aux))) return rc;
        unsigned char *fail = (unsigned char *)aux;
        unsigned int *list = (unsigned int *)((char *)aux + fail_bytes);
        unsigned int *count = list + Bc;
        unsigned char *fail1 = (unsigned char *)(count + 4);
        unsigned int *list1 = (unsigned int *)(fail1 + fail_bytes);
        unsigned int *count1 = list1 + Bc;
        CK(cudaMemsetAsync(aux, 0, 2 * (fail_bytes + Bc * 4 + 16), ln.stream));  // fail bytes, lists and counters of both levels
        if (want_flags) CK(cudaMemsetAsync(vf.dev, 0, Bc * fw * 8, ln.stream));

        const bool fastN = T.fast_logn > 0 && !lean_phase;
        if (!(fastN || T.er_logn > 0)) CK(cudaMemsetAsync(vp.dev, 0, Bc * 4, ln.stream));  // the NTT checks write path = 0 themselves
        if (fastN) {
            // optimistic-optimistic: all n = N shares on one degree-d polynomial <=> the top N-m coefficients of the inverse
            // NTT vanish; then the lowest d+t+1 agree as well (path 0, no flags).  Items that fail go to the dense check.
            NttArgs na{};
            na.in = (const uint4 *)vi.dev;
            na.out = (uint4 *)vc.dev;
            na.tw = T.itw;
            na.B = (long long)Bc;
            na.in_sb = vi.sb; na.in_sc = vi.sj;
            na.out_sb = T.mout; na.out_sr = 1;
            na.cols = (int)S;
            na.n = (int)S;
            na.err = ctx->d_status;
            na.in_map = P.in_map;
            na.scale = T.iscale;
            na.m = (int)m;
            na.mout = T.mout;
            na.fail = fail1;
            na.fail_list = list1; na.fail_count = count1;   // the check lists its failing items itself (no compaction launch)
            na.path = (int *)vp.dev;
            if ((rc = launch_ntt<1>(ctx, ln.stream, T.fast_logn, na))) return rc;
        }

        // decoder arguments (shared by the staged decoder and robust_kernel)
        RobustArgs r{};
        r.in = (const uint4 *)vi.dev;
        r.in_sb = vi.sb; r.in_sc = vi.sj;
        r.B = (long long)Bc;
        r.list = list;
        r.count = count;
        r.S = (int)S; r.m = (int)m; r.t = (int)t; r.needed = (int)needed; r.rmax = T.rmax; r.fast = T.fast;
        r.att_P = T.att_P; r.att_nsyn = T.att_nsyn; r.att_maxL = T.att_maxL; r.att_uoff = T.att_uoff;
        r.u2 = T.u2; r.sid = T.sid; r.tw = T.tw; r.itw = T.ritw; r.uinv = T.uinv;
        { int lg = 0; while ((1 << lg) < domain_size(n)) ++lg; r.logn = lg; } r.xs = T.xs; r.xinv = T.xinv; r.Lc = T.Lc; r.Veval = T.Veval; r.order = P.order;
        r.coeffs = (uint4 *)vc.dev;
        r.mout = T.mout;
        r.path = (int *)vp.dev;
        r.flags = want_flags ? (unsigned long long *)vf.dev : nullptr;
        r.flag_words = fw;
        r.fail_any = ctx->d_status + 2;
        const int threads = 128;
        long long blocks = std::min<long long>((long long)ctx->num_sms * HB_ROBUST_MINB, (long long)((Bc + threads - 1) / threads));
        if (blocks < 1) blocks = 1;

        // Large failing sets of the all-points check (synchronous calls): straight to the staged decoder, which settles path 0
        // (errors beyond the examined prefix only), the OEC round and the flags from the error positions; only what it cannot
        // decode takes the dense check below.
        // session-sized batches: no compaction pass -- the decoder's threads look at fail[] themselves and compute Lc*y; unless the
        // context has just seen an attack (large failing sets): then the count is worth a synchronisation, because it opens the
        // staged decoder
        // asynchronous calls of a context under attack (learnt at the last hbmpc_ctx_synchronize) compact too: the staged decoder then
        // runs in device-count mode (no host decision), see staged_decode
        const bool dev_route = ctx->async && ctx->async_staged && ctx->attack_seen && T.fast && (Bc >= 2048 || ctx->force_async_staged) && !lean_phase;
        const bool scan_early = Bc <= ctx->scan_max && !lean_phase && !(ctx->attack_seen && !ctx->async && Bc >= 2048) && !dev_route;
        // Synchronous calls of 2048 .. scan_max chunks: the failing items are compacted and robust_kernel is enqueued for them at once,
        // told to do nothing when there are >= SPEC_MIN of them; the call's one synchronisation then also brings the count, and a
        // large failing set takes the scouts / staged decoder below -- on the FIRST attacked call, at the price of one or two
        // near-empty launches for honest ones (round 1 learnt of an attack only for the context's next calls).
        const bool spec_small = scan_early && !ctx->async && Bc >= 2048 && T.fast && T.rmax >= 1 && !ctx->no_sync_count;
        const bool scan = scan_early && !spec_small;
        bool staged_direct = false;
        unsigned int *dense_list = list1, *dense_count = count1;
        if (fastN && T.fast && !ctx->async && !scan_early) {
            CK(cudaMemcpyAsync(ctx->h_spec, count1, sizeof(unsigned int), cudaMemcpyDeviceToHost, ln.stream));
            CK(cudaStreamSynchronize(ln.stream));
            const unsigned int c1 = ctx->h_spec[0];
            if (Bc <= ctx->scan_max && c1 < ctx->staged_min / 2) ctx->attack_seen = false;  // the attack is over
            if (c1 >= ctx->staged_min && !ctx->no_staged_direct) {
                // More than one wave of work: a few scouts first (hist_only: nothing is written).  When the same <= t senders
                // are wrong in (almost) every scout, the persistent-attacker shortcut further down (dense interpolation from the
                // senders believed honest) beats decoding every item.  Smaller failing sets are decoded without asking.
                const unsigned int W1 = (unsigned int)staged_wave_slots(ctx, T, (int)t), SC = 64;
                RobustArgs rd = r;
                rd.list = list1;
                rd.count = count1;
                unsigned int *left[2] = {nullptr, nullptr};
                bool persistent = false;
                if (!ctx->no_speculation && c1 > W1) {
                    void *histbuf = nullptr;
                    if ((rc = scratch_get(ctx, ln, 9, 4096, &histbuf))) return rc;
                    unsigned int *hist = (unsigned int *)histbuf;
                    CK(cudaMemsetAsync(hist, 0, S * sizeof(unsigned int), ln.stream));
                    RobustArgs rs = rd;
                    rs.hist = hist;
                    rs.hist_only = 1;
                    if ((rc = staged_decode(ctx, ln, T, rs, P.in_map, 0, SC, blocks, threads, left, c1))) return rc;
                    CK(cudaMemcpyAsync(ctx->h_spec + 8, hist, S * sizeof(unsigned int), cudaMemcpyDeviceToHost, ln.stream));
                    CK(cudaStreamSynchronize(ln.stream));
                    size_t nsus = 0;
                    for (size_t j = 0; j < S; ++j)
                        if (ctx->h_spec[8 + j] >= SC * 3 / 4) ++nsus;  // persistent attackers are wrong (almost) every time
                    persistent = nsus >= 1 && nsus <= t && S - nsus >= m;
                }
                if (!persistent) {
                    if ((rc = staged_decode(ctx, ln, T, rd, P.in_map, 0, c1, blocks, threads, left))) return rc;
                    staged_direct = true;
                    dense_list = left[0];
                    dense_count = left[1];
                }  // else: every failing item takes the dense check and the shortcut
            }
        }

        const bool er_any = T.er_logn > 0 && !fastN, er_all = er_any && T.er_all;
        const bool erasure = er_any && !er_all;   // the erasure check replaces the dense check (no flags wanted)
        if (er_any) {
            void *tmp = nullptr;
            if ((rc = scratch_get(ctx, ln, 7, Bc * (size_t)T.mout * 32, &tmp))) return rc;
            NttArgs na{};
            na.in = (const uint4 *)vi.dev;
            na.out = (uint4 *)tmp;
            na.tw = T.itw;
            na.B = (long long)Bc;
            na.in_sb = vi.sb; na.in_sc = vi.sj;
            na.out_sb = T.mout; na.out_sr = 1;
            na.cols = 1 << T.er_logn;
            na.n = 1 << T.er_logn;
            na.err = ctx->d_status;
            na.in_map = er_all ? P.in_map : P.er_in_map;
            na.wt = T.er_wt;
            na.m = T.er_zero_from;
            na.mout = T.er_h1;
            na.hi_top = T.er_hi_top;
            na.hi_cnt = T.er_h2;
            na.fail = er_all ? fail1 : fail;
            na.fail_list = er_all ? list1 : (scan ? nullptr : list);     // the check lists its failing items itself
            na.fail_count = er_all ? count1 : (scan ? nullptr : count);
            na.path = (int *)vp.dev;
            if ((rc = launch_ntt<2>(ctx, ln.stream, T.er_logn, na))) return rc;
            MatvecArgs tr{};
            tr.M = T.er_tri;
            tr.in = (const uint4 *)tmp;
            tr.out = (uint4 *)vc.dev;
            tr.R = T.mout;
            tr.C = T.mout;
            tr.B = (long long)Bc;
            tr.in_sb = T.mout; tr.in_sc = 1; tr.in_chunk_major = 1;
            tr.out_sb = T.mout; tr.out_sr = 1;
            tr.row_len = T.er_row_len;
            tr.row_start = T.er_row_start;
            if ((rc = launch_matvec(ctx, ln, tr, 0))) return rc;
            // (er_all: like the all-points check, only the chunks it rejects -- listed in list1 -- see the dense check, which sets the flags)
        }
        MatvecArgs a{};
        if (fastN || er_all) { a.item_list = dense_list; a.item_count = dense_count; }
        a.M = T.M;
        a.in = (const uint4 *)vi.dev;
        a.out = (uint4 *)vc.dev;
        a.R = T.R;
        a.C = T.C;
        a.B = (long long)Bc;
        a.in_sb = vi.sb; a.in_sc = vi.sj;
        a.in_chunk_major = sender_major ? 0 : 1;
        a.out_sb = T.mout;
        a.out_sr = 1;
        a.col_map = P.col_map;
        a.n_chk = T.n_chk;
        a.n_gate = T.n_gate;
        a.chk_map = P.chk_map;
        a.fail = fail;
        if (!scan) { a.fail_list = list; a.fail_count = count; }   // list mode: the check appends its failing items (no compaction launch)
        a.flags = want_flags ? (unsigned long long *)vf.dev : nullptr;
        if (!erasure && (rc = launch_matvec(ctx, ln, a, fw))) return rc;
        if (erasure && !lean_phase && !scan) {
            // the robust decoder corrects Lc*y[lowest d+1] by linearity: provide it for the failing items
            MatvecArgs lc{};
            lc.M = T.Lc;
            lc.in = (const uint4 *)vi.dev;
            lc.out = (uint4 *)vc.dev;
            lc.R = T.mout;
            lc.C = (int)m;
            lc.B = (long long)Bc;
            lc.in_sb = vi.sb; lc.in_sc = vi.sj;
            lc.in_chunk_major = sender_major ? 0 : 1;
            lc.out_sb = T.mout; lc.out_sr = 1;
            lc.col_map = P.col_map;
            lc.item_list = list; lc.item_count = count;
            if ((rc = launch_matvec(ctx, ln, lc, 0))) return rc;
        }

        if (lean_phase) {
            CK(cudaMemcpyAsync(ctx->h_counts + b0 / chunk_full, count, sizeof(unsigned int), cudaMemcpyDeviceToHost, ln.stream));
            if ((rc = chunk_commit(ctx, ln, bc, b0, Bc, vc))) return rc;
            return chunk_commit(ctx, ln, bp, b0, Bc, vp);
        }
        r.fail_scan = scan ? fail : nullptr;
        r.dense_fail_flag = (Bc >= 2048 && (scan || ctx->async)) ? ctx->d_status + 1 : nullptr;
        r.attack_min = (unsigned int)ctx->staged_min;
        r.need_lc = (scan && erasure) ? 1 : 0;
        void *ws = nullptr;
        if ((rc = scratch_get(ctx, ln, 6, (size_t)blocks * threads * lay.total * 32 + 64, &ws))) return rc;
        r.ws = (uint4 *)ws;
        r.ws_elems = lay.total;
        auto launch_robust = [&](const RobustArgs &ra) -> int {
            robust_kernel<<<(unsigned)blocks, threads, 0, ln.stream>>>(ra);
            ctx->launches++;
            CK(cudaGetLastError());
            return 0;
        };
        // Persistent-attacker shortcut (synchronous calls with many failing items): decode a few scouts, and if the same
        // <= t senders are wrong in most of them, interpolate every other failing item from senders believed honest and
        // verify against ALL supplied shares with one dense launch; whatever is not explained by <= t errors goes to the
        // full decoder.  Results are identical to decoding everything (see robust.cuh: spec_finalize_kernel).
        const unsigned int SCOUTS = 64, SPEC_MIN = 1024;
        bool done = false;
        // large failing sets go through the staged decoder (needs the count on the host: synchronous calls only)
        auto decode_list = [&](const RobustArgs &ra, unsigned int cnt_host) -> int {
            if (T.fast && !staged_direct && !ra.hist && cnt_host != UINT_MAX && cnt_host > ra.list_first && (size_t)(cnt_host - ra.list_first) >= ctx->staged_min)
                return staged_decode(ctx, ln, T, ra, P.in_map, ra.list_first, cnt_host, blocks, threads);
            if (dev_route && cnt_host == UINT_MAX && !ra.hist && ra.list_first == 0) {
                ctx->dev_route_calls++;
                return staged_decode(ctx, ln, T, ra, P.in_map, 0, (unsigned int)Bc, blocks, threads, nullptr, 0, false, 0, ra.count);
            }
            return launch_robust(ra);
        };
        unsigned int cnt_host = UINT_MAX;
        if (!scan && !ctx->async && T.rmax >= 1) {
            if (spec_small) {   // enqueued before the count is known: decodes a small failing set, skips a large one, reports the count
                RobustArgs rq = r;
                rq.skip_above = SPEC_MIN;
                rq.count_out = ctx->h_spec;   // pinned host memory (device-accessible under unified addressing): no copy to wait for
                if ((rc = launch_robust(rq))) return rc;
            } else {
                CK(cudaMemcpyAsync(ctx->h_spec, count, sizeof(unsigned int), cudaMemcpyDeviceToHost, ln.stream));
            }
            CK(cudaStreamSynchronize(ln.stream));
            cnt_host = ((volatile unsigned int *)ctx->h_spec)[0];
            if (spec_small && cnt_host < SPEC_MIN) done = true;
            if (spec_small && cnt_host >= ctx->staged_min) ctx->attack_seen = true;   // later calls of this shape may skip the dense check
            if (Bc <= ctx->scan_max && cnt_host < ctx->staged_min / 2 && !staged_direct) ctx->attack_seen = false;  // the attack is over
        }
        if (cnt_host != UINT_MAX && !ctx->no_speculation) {
            const unsigned int cnt = cnt_host;
            if (cnt >= SPEC_MIN) {
                // ---- scouts
                void *histbuf = nullptr;
                if ((rc = scratch_get(ctx, ln, 9, 4096, &histbuf))) return rc;
                unsigned int *hist = (unsigned int *)histbuf;
                CK(cudaMemsetAsync(hist, 0, S * sizeof(unsigned int), ln.stream));
                RobustArgs rs = r;
                rs.list_first = 0;
                rs.list_max = SCOUTS;
                rs.hist = hist;
                rs.clear_fail = fail;
                rs.need_lc = erasure ? 1 : 0;
                if ((rc = launch_robust(rs))) return rc;
                CK(cudaMemcpyAsync(ctx->h_spec + 8, hist, S * sizeof(unsigned int), cudaMemcpyDeviceToHost, ln.stream));
                CK(cudaStreamSynchronize(ln.stream));
                std::vector<char> suspect(S, 0);
                size_t nsus = 0;
                for (size_t j = 0; j < S; ++j)
                    if (ctx->h_spec[8 + j] >= SCOUTS / 2) { suspect[j] = 1; ++nsus; }
                if (nsus >= 1 && nsus <= t && S - nsus >= m) {
                    // ---- tables for "interpolate from the lowest m unsuspected ids, compare with every other supplied share"
                    std::vector<HFr> dom = domain_elements(n, n);
                    std::vector<int> I, O, pos_of(S);
                    for (size_t i = 0; i < S; ++i) {
                        pos_of[order[i]] = (int)i;
                        if (!suspect[order[i]] && I.size() < m) I.push_back(order[i]);
                        else O.push_back(order[i]);
                    }
                    std::vector<HFr> xi(m), xo(O.size());
                    for (size_t i = 0; i < m; ++i) xi[i] = dom[ids[I[i]]];
                    for (size_t i = 0; i < O.size(); ++i) xo[i] = dom[ids[O[i]]];
                    Lagrange Ls = lagrange_basis(xi);
                    std::vector<HFr> Ms = lagrange_eval_rows(xi, Ls, xo);
                    Ms.insert(Ms.end(), Ls.Lc.begin(), Ls.Lc.begin() + (size_t)T.mout * m);
                    std::vector<uint32_t> mw;
                    to_u32(Ms, mw);
                    std::vector<int> maps(I);
                    maps.insert(maps.end(), O.begin(), O.end());
                    maps.insert(maps.end(), pos_of.begin(), pos_of.end());
                    const int fws = (int)((S + 63) / 64);
                    const size_t off_maps = ((mw.size() * 4 + 255) / 256) * 256, off_sf = off_maps + ((maps.size() * 4 + 255) / 256) * 256;
                    void *sp = nullptr;
                    if ((rc = scratch_get(ctx, ln, 9, off_sf + Bc * (size_t)fws * 8 + 4096, &sp))) return rc;
                    CK(cudaMemcpyAsync(sp, mw.data(), mw.size() * 4, cudaMemcpyHostToDevice, ln.stream));
                    CK(cudaMemcpyAsync((char *)sp + off_maps, maps.data(), maps.size() * 4, cudaMemcpyHostToDevice, ln.stream));
                    unsigned long long *sflags = (unsigned long long *)((char *)sp + off_sf);
                    CK(cudaMemsetAsync(sflags, 0, Bc * (size_t)fws * 8, ln.stream));
                    ctx->h_spec[1] = cnt - SCOUTS;
                    CK(cudaMemcpyAsync(count1, ctx->h_spec + 1, sizeof(unsigned int), cudaMemcpyHostToDevice, ln.stream));
                    const int *dmaps = (const int *)((char *)sp + off_maps);
                    MatvecArgs sm{};
                    sm.M = (const uint4 *)sp;
                    sm.in = (const uint4 *)vi.dev;
                    sm.out = (uint4 *)vc.dev;
                    sm.R = (int)(O.size() + T.mout);
                    sm.C = (int)m;
                    sm.B = (long long)Bc;
                    sm.in_sb = vi.sb; sm.in_sc = vi.sj;
                    sm.in_chunk_major = sender_major ? 0 : 1;
                    sm.out_sb = T.mout; sm.out_sr = 1;
                    sm.col_map = dmaps;
                    sm.n_chk = (int)O.size();
                    sm.n_gate = 0;
                    sm.chk_map = dmaps + m;
                    sm.flags = sflags;
                    sm.item_list = list + SCOUTS;
                    sm.item_count = count1;
                    if ((rc = launch_matvec(ctx, ln, sm, fws))) return rc;
                    SpecArgs sa{};
                    sa.list = list; sa.first = SCOUTS; sa.count = cnt;
                    sa.sflags = sflags;
                    sa.flags = a.flags;
                    sa.flag_words = fws;
                    sa.pos_of = dmaps + S;
                    sa.S = (int)S; sa.t = (int)t; sa.needed = (int)needed; sa.rmax = T.rmax;
                    sa.path = (int *)vp.dev;
                    sa.fail = fail;
                    sa.fail_any = ctx->d_status + 2;
                    sa.coeffs = (uint4 *)vc.dev;
                    sa.mout = T.mout;
                    spec_finalize_kernel<<<ctx->num_sms * 4, 256, 0, ln.stream>>>(sa);
                    ctx->launches++;
                    CK(cudaGetLastError());
                    // ---- what speculation could not explain: full decoder (the scouts are already done)
                    CK(cudaMemsetAsync(count, 0, sizeof(unsigned int), ln.stream));
                    compact_kernel<<<ctx->num_sms * 4, 256, 0, ln.stream>>>(fail, (long long)Bc, list, count);
                    ctx->launches++;
                    CK(cudaGetLastError());
                    RobustArgs rr = r;
                    rr.need_lc = 1;  // the speculative launch overwrote coeffs
                    CK(cudaMemcpyAsync(ctx->h_spec, count, sizeof(unsigned int), cudaMemcpyDeviceToHost, ln.stream));
                    CK(cudaStreamSynchronize(ln.stream));
                    if ((rc = decode_list(rr, ctx->h_spec[0]))) return rc;
                } else {
                    RobustArgs rr = r;
                    rr.list_first = SCOUTS;
                    if ((rc = decode_list(rr, cnt))) return rc;
                }
                done = true;
            }
        }
        if (!done && (rc = decode_list(r, cnt_host))) return rc;

        if (want_secrets) {
            gather_first_kernel<<<ctx->num_sms * 4, 256, 0, ln.stream>>>((long long)Bc, (long long)m, (const uint4 *)vc.dev, (uint4 *)vs.dev);
            ctx->launches++;
            CK(cudaGetLastError());
        }
        if ((rc = chunk_commit(ctx, ln, bc, b0, Bc, vc))) return rc;
        if ((rc = chunk_commit(ctx, ln, bp, b0, Bc, vp))) return rc;
        if (want_flags && (rc = chunk_commit(ctx, ln, bf, b0, Bc, vf))) return rc;
        if (want_secrets && (rc = chunk_commit(ctx, ln, bs, b0, Bc, vs))) return rc;
        return 0;
    };
    const bool any_host = bi.host || bc.host || bp.host || (want_flags && bf.host) || (want_secrets && bs.host);
    if (!lean) return run_batched(ctx, B, any_host, std::max<size_t>(S, T.mout) * 32, body);
    int rc = run_batched(ctx, B, true, std::max<size_t>(S, T.mout) * 32, body, false);
    if (rc) return rc;
    lean_phase = false;
    for (size_t b0 = 0, c = 0; b0 < B; b0 += chunk_full, ++c) {
        if (!ctx->h_counts[c]) continue;
        if ((rc = body(ctx->lanes[1], b0, std::min(chunk_full, B - b0)))) return rc;
    }
    CK(cudaStreamSynchronize(ctx->lanes[1].stream));
    return collect_status(ctx);
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

// N1: the per-sender / per-recipient vectors as they lie in the message payloads (one host array per sender or recipient, e.g.
// payload + 8 of an ark-serialize Vec<F>): no host-side gather into one contiguous [S][B] array, no scatter out of [n][B]
// (batch_recon.rs:174-175, 339, 419; common/utils.rs:3-21).
extern "C" int hbmpc_batch_recover_msgs(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B,
                                        const uint64_t *const *sender_evals, uint64_t *coeffs, int32_t *path, uint64_t *flags) {
    if (!ctx || !sender_evals) return HBMPC_INVALID_INPUT;
    return recover_impl(ctx, n, d, t, S, sender_ids, B, nullptr, true, coeffs, false, nullptr, path, flags, 0, (void *const *)sender_evals);
}
extern "C" int hbmpc_batch_recover_secrets_msgs(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B,
                                                const uint64_t *const *sender_evals, uint64_t *secrets, int32_t *path) {
    if (!ctx || !sender_evals) return HBMPC_INVALID_INPUT;
    return recover_impl(ctx, n, d, t, S, sender_ids, B, nullptr, true, nullptr, true, secrets, path, nullptr, 0, (void *const *)sender_evals);
}
extern "C" int hbmpc_apply_vandermonde_msgs(hbmpc_ctx *ctx, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *const *recipient_out) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (cols == 0 || cols > 256) return HBMPC_INVALID_INPUT;
    if (!domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;
    if (B == 0) return HBMPC_SUCCESS;
    if (!in || !host_ptr_array_ok((const void *const *)recipient_out, n)) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    return apply_domain(ctx, n, cols, B, in, nullptr, 1, 0, (void *const *)recipient_out);
}

// ------------------------------------------------------------------------------------------------ a10: NonRobustShare::recover_secret, batched
static std::map<TableKey, NonRobustTables> &nr_cache(hbmpc_ctx *ctx) {
    if (!ctx->nonrobust) ctx->nonrobust = new std::map<TableKey, NonRobustTables>();
    return *ctx->nonrobust;
}

extern "C" int hbmpc_nonrobust_recover_batch(hbmpc_ctx *ctx, size_t n, size_t deg, size_t S, const size_t *ids, size_t B, const uint64_t *shares,
                                             int sender_major, uint64_t *coeffs, uint64_t *secrets, int32_t *status) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    // validation order of common/share/shamir.rs:204-232
    if (S == 0 || !ids) return HBMPC_INVALID_INPUT;
    if (S > 256) return HBMPC_INVALID_INPUT;
    std::vector<int> order(S);  // order[i] = arrival index of the share with the i-th smallest id
    for (size_t i = 0; i < S; ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return ids[a] < ids[b]; });
    for (size_t i = 1; i < S; ++i)
        if (ids[order[i]] == ids[order[i - 1]]) return HBMPC_INVALID_INPUT;
    if (S < deg + 1) return HBMPC_INSUFFICIENT_SHARES;
    if (!domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;
    for (size_t i = 0; i < S; ++i)
        if (ids[i] >= n) return HBMPC_INVALID_INPUT;
    if (B == 0) return HBMPC_SUCCESS;
    if (!shares || !coeffs || !status) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    const size_t m = deg + 1;
    // Tables are keyed by the id SET (like K3): interpolation through all supplied points does not depend on the order in which
    // they arrived, which is network- and adversary-influenced; the order only enters through the per-call column map.
    TableKey key;
    key.kind = 4u;
    key.n = (uint32_t)n; key.d = (uint32_t)deg;
    for (size_t i = 0; i < S; ++i) key.idset[ids[i] >> 6] |= 1ull << (ids[i] & 63);
    auto &cache = nr_cache(ctx);
    auto it = cache.find(key);
    if (it == cache.end()) {
        if (cache.size() >= 256) {  // bounded, like the K3 cache
            for (auto &ln : ctx->lanes) CK(cudaStreamSynchronize(ln.stream));
            for (auto &e : cache)
                for (void *q : e.second.allocs) cudaFree(q);
            cache.clear();
        }
        std::vector<HFr> dom = domain_elements(n, n), xs(S);
        for (size_t i = 0; i < S; ++i) xs[i] = dom[ids[order[i]]];
        Lagrange L = lagrange_basis(xs);  // Lc[k][i]: coefficient k of the basis polynomial of the i-th smallest id
        // rows: coefficients deg+1 .. S-1 (must vanish: the DegreeMismatch check of shamir.rs:234-237), then 0 .. deg
        std::vector<HFr> M(L.Lc.begin() + m * S, L.Lc.end());
        M.insert(M.end(), L.Lc.begin(), L.Lc.begin() + m * S);
        NonRobustTables T;
        T.n_chk = (int)(S - m);
        T.R = (int)S;
        std::vector<int> chk(std::max<size_t>(S - m, 1), -1);
        int rc;
        if ((rc = upload_fr(ctx, M, &T.M, &T.allocs)) || (rc = upload(ctx, chk, &T.chk_map, &T.allocs))) {
            for (void *q : T.allocs) cudaFree(q);
            return abandon_call(ctx, rc);
        }
        const int N = domain_size(n);
        if (!ctx->no_fastpath && S == n && (size_t)N == n && N >= 2) {
            if ((rc = get_inverse_twiddles(ctx, N, &T.itw, &T.iscale))) return abandon_call(ctx, rc);
            while ((1 << T.fast_logn) < N) ++T.fast_logn;
        }
        it = cache.emplace(key, T).first;
    }
    const NonRobustTables &T = it->second;
    // per-call maps: column c (sorted position) reads arrival order[c]; domain index -> arrival index for the inverse transform
    const int *d_order = nullptr, *d_in_map = nullptr;
    {
        const int N = domain_size(n);
        const int *base = nullptr;
        int rc = maps_upload(ctx, [&](int *blk) -> size_t {
            for (size_t i = 0; i < S; ++i) blk[i] = order[i];
            for (int k = 0; k < N; ++k) blk[S + k] = -1;
            for (size_t i = 0; i < S; ++i) blk[S + ids[i]] = (int)i;
            return S + (size_t)N;
        }, &base);
        if (rc) return abandon_call(ctx, rc);
        d_order = base;
        d_in_map = base + S;
    }
    BatchBuf bi = make_buf(shares, B, (long long)S, sender_major != 0), bc = make_buf(coeffs, B, (long long)m, false);
    BatchBuf bs = make_buf(secrets, B, 1, false), bst = make_buf(status, B, 1, false, 4);
    auto body = [&](Lane &ln, size_t b0, size_t Bc) -> int {
        ChunkView vi, vc, vs, vst;
        int rc;
        if ((rc = chunk_prepare(ctx, ln, 0, bi, b0, Bc, true, vi))) return rc;
        if ((rc = chunk_prepare(ctx, ln, 1, bc, b0, Bc, false, vc))) return rc;
        if ((rc = chunk_prepare(ctx, ln, 2, bst, b0, Bc, false, vst))) return rc;
        if (secrets && (rc = chunk_prepare(ctx, ln, 4, bs, b0, Bc, false, vs))) return rc;
        void *aux = nullptr;
        const size_t fail_bytes = ((Bc + 15) / 16) * 16;
        if ((rc = scratch_get(ctx, ln, 5, fail_bytes, &aux))) return rc;
        CK(cudaMemsetAsync(aux, 0, fail_bytes, ln.stream));
        if (T.fast_logn > 0) {
            NttArgs na{};
            na.in = (const uint4 *)vi.dev;
            na.out = (uint4 *)vc.dev;
            na.tw = T.itw;
            na.B = (long long)Bc;
            na.in_sb = vi.sb; na.in_sc = vi.sj;
            na.out_sb = (long long)m; na.out_sr = 1;
            na.cols = (int)S;
            na.n = (int)S;
            na.err = ctx->d_status;
            na.in_map = d_in_map;
            na.scale = T.iscale;
            na.m = (int)m;
            na.mout = (int)m;
            na.fail = (unsigned char *)aux;
            if ((rc = launch_ntt<1>(ctx, ln.stream, T.fast_logn, na))) return rc;
        } else {
            MatvecArgs a{};
            a.M = T.M;
            a.in = (const uint4 *)vi.dev;
            a.out = (uint4 *)vc.dev;
            a.R = T.R;
            a.C = (int)S;
            a.B = (long long)Bc;
            a.in_sb = vi.sb; a.in_sc = vi.sj;
            a.in_chunk_major = sender_major ? 0 : 1;
            a.out_sb = (long long)m;
            a.out_sr = 1;
            a.col_map = d_order;
            a.n_chk = T.n_chk;
            a.n_gate = T.n_chk;
            a.chk_map = T.chk_map;
            a.fail = (unsigned char *)aux;
            if ((rc = launch_matvec(ctx, ln, a, 0))) return rc;
        }
        degree_status_kernel<<<ctx->num_sms * 4, 256, 0, ln.stream>>>((long long)Bc, (int)m, (uint4 *)vc.dev, (const unsigned char *)aux, (int *)vst.dev,
                                                                     secrets ? (uint4 *)vs.dev : nullptr);
        ctx->launches++;
        CK(cudaGetLastError());
        if ((rc = chunk_commit(ctx, ln, bc, b0, Bc, vc))) return rc;
        if ((rc = chunk_commit(ctx, ln, bst, b0, Bc, vst))) return rc;
        if (secrets && (rc = chunk_commit(ctx, ln, bs, b0, Bc, vs))) return rc;
        return 0;
    };
    return run_batched(ctx, B, bi.host || bc.host || bst.host || (secrets && bs.host), S * 32, body);
}

// ------------------------------------------------------------------------------------------------ K5
extern "C" int hbmpc_elementwise(hbmpc_ctx *ctx, int op, size_t count, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (op < 0 || op > 2) return HBMPC_INVALID_INPUT;
    if (count == 0) return HBMPC_SUCCESS;
    if (!a || !b || !out) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    BatchBuf ba = make_buf(a, count, 1, false), bb = make_buf(b, count, 1, false), bo = make_buf(out, count, 1, false);
    auto body = [&](Lane &ln, size_t b0, size_t Bc) -> int {
        ChunkView va, vb, vo;
        int rc;
        if ((rc = chunk_prepare(ctx, ln, 0, ba, b0, Bc, true, va))) return rc;
        if ((rc = chunk_prepare(ctx, ln, 1, bb, b0, Bc, true, vb))) return rc;
        if ((rc = chunk_prepare(ctx, ln, 2, bo, b0, Bc, false, vo))) return rc;
        long long blocks = std::min<long long>((long long)ctx->num_sms * 8, (long long)((Bc + 255) / 256));
        elementwise_kernel<<<(unsigned)blocks, 256, 0, ln.stream>>>(op, (long long)Bc, (const uint4 *)va.dev, (const uint4 *)vb.dev, (uint4 *)vo.dev,
                                                                   ctx->d_status);
        ctx->launches++;
        CK(cudaGetLastError());
        return chunk_commit(ctx, ln, bo, b0, Bc, vo);
    };
    return run_batched(ctx, count, ba.host || bb.host || bo.host, 32, body);
}

extern "C" int hbmpc_share_algebra_fused(hbmpc_ctx *ctx, int op, size_t count, const uint64_t *const *in, uint64_t *const *out) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (op < 0 || op > 2) return HBMPC_INVALID_INPUT;
    if (count == 0) return HBMPC_SUCCESS;
    if (!in || !out) return HBMPC_INVALID_INPUT;
    const int nin = op == 0 ? 3 : op == 1 ? 4 : 5, nout = op == 1 ? 2 : 1;
    for (int k = 0; k < nin; ++k)
        if (!in[k]) return HBMPC_INVALID_INPUT;
    for (int k = 0; k < nout; ++k)
        if (!out[k]) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    BatchBuf bi[5], bo[2];
    bool host = false;
    for (int k = 0; k < nin; ++k) { bi[k] = make_buf(in[k], count, 1, false); host = host || bi[k].host; }
    for (int k = 0; k < nout; ++k) { bo[k] = make_buf(out[k], count, 1, false); host = host || bo[k].host; }
    auto body = [&](Lane &ln, size_t b0, size_t Bc) -> int {
        ChunkView vi[5], vo[2];
        int rc;
        for (int k = 0; k < nin; ++k)
            if ((rc = chunk_prepare(ctx, ln, k, bi[k], b0, Bc, true, vi[k]))) return rc;
        for (int k = 0; k < nout; ++k)
            if ((rc = chunk_prepare(ctx, ln, 5 + k, bo[k], b0, Bc, false, vo[k]))) return rc;
        FusedArgs fa{};
        for (int k = 0; k < nin; ++k) fa.in[k] = (const uint4 *)vi[k].dev;
        for (int k = 0; k < nout; ++k) fa.out[k] = (uint4 *)vo[k].dev;
        fa.count = (long long)Bc;
        fa.err = ctx->d_status;
        const unsigned blocks = (unsigned)std::min<long long>((long long)ctx->num_sms * 8, (long long)((Bc + 255) / 256));
        if (op == 0) elementwise_fused_kernel<0><<<blocks, 256, 0, ln.stream>>>(fa);
        else if (op == 1) elementwise_fused_kernel<1><<<blocks, 256, 0, ln.stream>>>(fa);
        else elementwise_fused_kernel<2><<<blocks, 256, 0, ln.stream>>>(fa);
        ctx->launches++;
        CK(cudaGetLastError());
        for (int k = 0; k < nout; ++k)
            if ((rc = chunk_commit(ctx, ln, bo[k], b0, Bc, vo[k]))) return rc;
        return 0;
    };
    return run_batched(ctx, count, host, 32, body);
}

// ------------------------------------------------------------------------------------------------ N1: 48-byte share records
extern "C" int hbmpc_unpack_share_records(hbmpc_ctx *ctx, size_t count, const void *records, uint64_t *values, uint64_t *ids, uint64_t *degrees) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (count == 0) return HBMPC_SUCCESS;
    if (!records || !values) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    BatchBuf br = make_buf(records, count, 1, false, 48), bv = make_buf(values, count, 1, false, 32);
    BatchBuf bi = make_buf(ids, count, 1, false, 8), bd = make_buf(degrees, count, 1, false, 8);
    auto body = [&](Lane &ln, size_t b0, size_t Bc) -> int {
        ChunkView vr, vv, vi, vd;
        int rc;
        if ((rc = chunk_prepare(ctx, ln, 0, br, b0, Bc, true, vr))) return rc;
        if ((rc = chunk_prepare(ctx, ln, 1, bv, b0, Bc, false, vv))) return rc;
        if ((rc = chunk_prepare(ctx, ln, 2, bi, b0, Bc, false, vi))) return rc;
        if ((rc = chunk_prepare(ctx, ln, 3, bd, b0, Bc, false, vd))) return rc;
        long long blocks = std::min<long long>((long long)ctx->num_sms * 8, (long long)((Bc + 255) / 256));
        unpack_records_kernel<<<(unsigned)blocks, 256, 0, ln.stream>>>((long long)Bc, (const unsigned long long *)vr.dev, (unsigned long long *)vv.dev,
                                                                     (unsigned long long *)vi.dev, (unsigned long long *)vd.dev);
        ctx->launches++;
        CK(cudaGetLastError());
        if ((rc = chunk_commit(ctx, ln, bv, b0, Bc, vv))) return rc;
        if ((rc = chunk_commit(ctx, ln, bi, b0, Bc, vi))) return rc;
        return chunk_commit(ctx, ln, bd, b0, Bc, vd);
    };
    return run_batched(ctx, count, br.host || bv.host || (ids && bi.host) || (degrees && bd.host), 48, body);
}

// records[i] = (values[i], id = i / per_id, degree): per_id consecutive values belong to the same share id
// (per_id = 1: the n shares of one sharing, ids 0..n-1; per_id = B: recipient-major [n][B] batches)
extern "C" int hbmpc_pack_share_records(hbmpc_ctx *ctx, size_t count, const uint64_t *values, size_t per_id, size_t degree, void *records) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (count == 0) return HBMPC_SUCCESS;
    if (!records || !values || per_id == 0) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    BatchBuf br = make_buf(records, count, 1, false, 48), bv = make_buf(values, count, 1, false, 32);
    if (br.host || bv.host) {
        // ids depend on the global index: keep host calls in one piece (these payloads are message sized)
        Lane &ln = ctx->lanes[1];
        CK(cudaEventRecord(ctx->ev_main, ctx->main_stream()));
        CK(cudaStreamWaitEvent(ln.stream, ctx->ev_main, 0));
        ChunkView vr, vv;
        int rc;
        if ((rc = chunk_prepare(ctx, ln, 0, bv, 0, count, true, vv))) return rc;
        if ((rc = chunk_prepare(ctx, ln, 1, br, 0, count, false, vr))) return rc;
        long long blocks = std::min<long long>((long long)ctx->num_sms * 8, (long long)((count + 255) / 256));
        pack_records_kernel<<<(unsigned)blocks, 256, 0, ln.stream>>>((long long)count, (const unsigned long long *)vv.dev, (long long)per_id, degree, (unsigned long long *)vr.dev);
        ctx->launches++;
        CK(cudaGetLastError());
        if ((rc = chunk_commit(ctx, ln, br, 0, count, vr))) return rc;
        CK(cudaStreamSynchronize(ln.stream));
        return HBMPC_SUCCESS;
    }
    long long blocks = std::min<long long>((long long)ctx->num_sms * 8, (long long)((count + 255) / 256));
    pack_records_kernel<<<(unsigned)blocks, 256, 0, ctx->main_stream()>>>((long long)count, (const unsigned long long *)values, (long long)per_id, degree, (unsigned long long *)records);
    ctx->launches++;
    CK(cudaGetLastError());
    return ctx->async ? HBMPC_SUCCESS : collect_status(ctx);
}

// ------------------------------------------------------------------------------------------------ N4: device-side sampling
// accepted elements 0 .. want-1 of the Fp::rand stream of StdRng::from_seed(seed), laid out by `mode` (sampler.cuh); out: device or host
static int sample_stream(hbmpc_ctx *ctx, const uint8_t *seed, unsigned long long want, int mode, int per, size_t out_elems, uint64_t *out,
                         const uint64_t *secrets, size_t B) {
    if (!ctx || !seed || !out) return HBMPC_INVALID_INPUT;
    if (want == 0) return HBMPC_SUCCESS;
    cudaSetDevice(ctx->device);
    Lane &ln = ctx->lanes[0];
    cudaStream_t st = ln.stream;
    const bool host_out = !is_device_ptr(out);
    void *dout = out;
    int rc;
    if (host_out && (rc = scratch_get(ctx, ln, 1, out_elems * 32, &dout))) return rc;
    SampleArgs a{};
    for (int i = 0; i < 8; ++i) a.key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
    a.out = (uint4 *)dout;
    a.want = want;
    a.mode = mode;
    a.per = per;
    // acceptance probability r / 2^255 = 0.9057: examine want / 0.9 candidates (+ slack); more only if that was not enough
    unsigned long long ncand = (unsigned long long)((double)want / 0.9) + 8192;
    for (;;) {
        if (ncand >= (1ull << 32)) return HBMPC_INVALID_INPUT;
        const unsigned int nblk = (unsigned int)((ncand + SAMPLE_THREADS - 1) / SAMPLE_THREADS);
        void *cnt = nullptr;
        if ((rc = scratch_get(ctx, ln, 0, (size_t)nblk * 4 + 64, &cnt))) return rc;
        a.ncand = ncand;
        a.counts = (unsigned int *)cnt + 16;
        a.total = (unsigned long long *)cnt;
        sample_count_kernel<<<nblk, SAMPLE_THREADS, 0, st>>>(a);
        sample_scan_kernel<<<1, 1024, 0, st>>>(a.counts, nblk, a.total);
        ctx->launches += 2;
        CK(cudaGetLastError());
        unsigned long long total = 0;
        CK(cudaMemcpyAsync(&total, a.total, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (total >= want) {
            sample_emit_kernel<<<nblk, SAMPLE_THREADS, 0, st>>>(a);
            ctx->launches++;
            CK(cudaGetLastError());
            break;
        }
        ncand *= 2;
    }
    if (mode == 2 && secrets) {   // coefficient 0 of every sharing = the caller's secret
        const bool host_sec = !is_device_ptr(secrets);
        if (host_sec) CK(cudaMemcpy2DAsync(dout, (size_t)(per + 1) * 32, secrets, 32, 32, B, cudaMemcpyHostToDevice, st));
        else CK(cudaMemcpy2DAsync(dout, (size_t)(per + 1) * 32, secrets, 32, 32, B, cudaMemcpyDeviceToDevice, st));
    }
    if (host_out) CK(cudaMemcpyAsync(out, dout, out_elems * 32, cudaMemcpyDeviceToHost, st));
    if (host_out || !ctx->async) CK(cudaStreamSynchronize(st));
    return HBMPC_SUCCESS;
}

extern "C" int hbmpc_sample_fr_batch(hbmpc_ctx *ctx, const uint8_t *seed32, size_t count, uint64_t *out) {
    return sample_stream(ctx, seed32, count, 0, 0, count, out, nullptr, 0);
}
extern "C" int hbmpc_sample_polynomials(hbmpc_ctx *ctx, const uint8_t *seed32, size_t B, size_t d, const uint64_t *secrets, uint64_t *coeffs) {
    if (B == 0) return HBMPC_SUCCESS;
    if (d > 255) return HBMPC_INVALID_INPUT;
    if (secrets) return sample_stream(ctx, seed32, (unsigned long long)B * (d + 1), 2, (int)d, B * (d + 1), coeffs, secrets, B);
    return sample_stream(ctx, seed32, (unsigned long long)B * (d + 2), 1, (int)d, B * (d + 1), coeffs, nullptr, B);
}

// K1 with the reference's own argument meaning: RobustShare::compute_shares(secret, n, degree, None, rng) (robust_interpolate.rs:52-82)
// for B secrets whose polynomials are drawn on the device from StdRng::from_seed(seed32) exactly as B consecutive calls on one
// generator would draw them (d+1 draws per sharing, the first overwritten by the secret).  A dealer uploads 32 bytes per secret.
extern "C" int hbmpc_share_secrets_batch(hbmpc_ctx *ctx, const uint8_t *seed32, size_t n, size_t d, size_t B, const uint64_t *secrets,
                                         uint64_t *shares, uint64_t *coeffs_out) {
    if (!ctx || !seed32) return HBMPC_INVALID_INPUT;
    if (n <= d) return HBMPC_INVALID_INPUT;                 // robust_interpolate.rs:59-64
    if (!domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;   // :65-66
    if (B == 0) return HBMPC_SUCCESS;
    if (!secrets || !shares || d > 255) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    if (ctx->async)   // the coefficient scratch of an earlier call may still be read by its pipelined chunks
        for (int i = 1; i < NLANES; ++i) CK(cudaStreamSynchronize(ctx->lanes[i].stream));
    void *dc = nullptr;
    const bool own = !(coeffs_out && is_device_ptr(coeffs_out));
    int rc;
    if (own) {
        if ((rc = scratch_get(ctx, ctx->lanes[0], 12, B * (d + 1) * 32, &dc))) return rc;
    } else dc = coeffs_out;
    if ((rc = sample_stream(ctx, seed32, (unsigned long long)B * (d + 1), 2, (int)d, B * (d + 1), (uint64_t *)dc, secrets, B))) return rc;
    if (own && coeffs_out) {
        CK(cudaMemcpyAsync(coeffs_out, dc, B * (d + 1) * 32, cudaMemcpyDeviceToHost, ctx->main_stream()));
        CK(cudaStreamSynchronize(ctx->main_stream()));
    }
    return apply_domain(ctx, n, d + 1, B, (const uint64_t *)dc, shares, 0);
}

// ------------------------------------------------------------------------------------------------ single-process multi-GPU groups
// The reference party is ONE process (honeybadger/mod.rs:245-257) that issues all sessions' work before awaiting (:1362-1375); its
// batches are maps over independent secrets / chunks / codewords.  A group owns one context per device and splits every
// batch into contiguous ranges [g*B/G, (g+1)*B/G), one host thread per device, constant tables replicated, NO collective: results
// land in the caller's (host) buffers in batch order.  Sender-major / recipient-major arrays are sharded along their batch axis
// through the leading dimension of the per-device calls.
struct hbmpc_group {
    std::vector<hbmpc_ctx *> ctx;
};

extern "C" int hbmpc_group_create(const int *devices, size_t n_devices, hbmpc_group **out) {
    if (!out || !devices || n_devices == 0 || n_devices > 64) return HBMPC_INVALID_INPUT;
    *out = nullptr;
    hbmpc_group *g = new hbmpc_group();
    for (size_t i = 0; i < n_devices; ++i) {
        hbmpc_ctx *c = nullptr;
        int rc = hbmpc_ctx_create(devices[i], &c);
        if (rc != HBMPC_SUCCESS) {
            for (hbmpc_ctx *p : g->ctx) hbmpc_ctx_destroy(p);
            delete g;
            return rc;
        }
        g->ctx.push_back(c);
    }
    *out = g;
    return HBMPC_SUCCESS;
}
extern "C" void hbmpc_group_destroy(hbmpc_group *g) {
    if (!g) return;
    for (hbmpc_ctx *p : g->ctx) hbmpc_ctx_destroy(p);
    delete g;
}
extern "C" size_t hbmpc_group_size(const hbmpc_group *g) { return g ? g->ctx.size() : 0; }
extern "C" hbmpc_ctx *hbmpc_group_ctx(hbmpc_group *g, size_t i) { return (g && i < g->ctx.size()) ? g->ctx[i] : nullptr; }
// contiguous range of member i: [lo, hi)
extern "C" void hbmpc_group_shard_range(const hbmpc_group *g, size_t B, size_t i, size_t *lo, size_t *hi) {
    const size_t G = g ? g->ctx.size() : 1;
    if (lo) *lo = B * i / G;
    if (hi) *hi = B * (i + 1) / G;
}

// runs fn(member, lo, hi) on one host thread per device; returns the first non-zero status in member order
template <typename Fn>
static int group_run(hbmpc_group *g, size_t B, Fn fn) {
    const size_t G = g->ctx.size();
    std::vector<int> rcs(G, 0);
    std::vector<std::thread> th;
    for (size_t i = 0; i < G; ++i) {
        const size_t lo = B * i / G, hi = B * (i + 1) / G;
        if (hi == lo) continue;
        th.emplace_back([&, i, lo, hi]() {
            cudaSetDevice(g->ctx[i]->device);
            rcs[i] = fn(g->ctx[i], lo, hi);
        });
    }
    for (auto &t : th) t.join();
    for (size_t i = 0; i < G; ++i)
        if (rcs[i]) return rcs[i];
    return HBMPC_SUCCESS;
}
static bool group_host_only(std::initializer_list<const void *> ptrs) {
    for (const void *p : ptrs)
        if (p && is_device_ptr(p)) return false;
    return true;
}

extern "C" int hbmpc_group_compute_shares_batch(hbmpc_group *g, size_t n, size_t d, size_t B, const uint64_t *coeffs, uint64_t *shares) {
    if (!g) return HBMPC_INVALID_INPUT;
    if (n <= d) return HBMPC_INVALID_INPUT;
    if (!domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;
    if (B == 0) return HBMPC_SUCCESS;
    if (!coeffs || !shares || !group_host_only({coeffs, shares})) return HBMPC_INVALID_INPUT;   // a group call takes host buffers
    return group_run(g, B, [&](hbmpc_ctx *c, size_t lo, size_t hi) {
        return hbmpc_compute_shares_batch(c, n, d, hi - lo, coeffs + lo * (d + 1) * 4, shares + lo * n * 4);
    });
}

extern "C" int hbmpc_group_apply_vandermonde_batch(hbmpc_group *g, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *out, int recipient_major) {
    if (!g) return HBMPC_INVALID_INPUT;
    if (cols == 0 || cols > 256) return HBMPC_INVALID_INPUT;
    if (!domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;
    if (B == 0) return HBMPC_SUCCESS;
    if (!in || !out || !group_host_only({in, out})) return HBMPC_INVALID_INPUT;
    return group_run(g, B, [&](hbmpc_ctx *c, size_t lo, size_t hi) {
        cudaSetDevice(c->device);
        return apply_domain(c, n, cols, hi - lo, in + lo * cols * 4, recipient_major ? out + lo * 4 : out + lo * n * 4, recipient_major, recipient_major ? B : 0);
    });
}

static int group_recover(hbmpc_group *g, size_t n, size_t d, size_t t, size_t S, const size_t *ids, size_t B, const uint64_t *in, bool sender_major,
                         uint64_t *coeffs, bool secrets_only, uint64_t *secrets, int32_t *path, uint64_t *flags) {
    if (!g) return HBMPC_INVALID_INPUT;
    if (B == 0) return HBMPC_INVALID_INPUT;
    if (!group_host_only({in, coeffs, secrets, path, flags})) return HBMPC_INVALID_INPUT;
    const size_t m = d + 1, fw = (S + 63) / 64;
    return group_run(g, B, [&](hbmpc_ctx *c, size_t lo, size_t hi) {
        return recover_impl(c, n, d, t, S, ids, hi - lo, sender_major ? in + lo * 4 : in + lo * S * 4, sender_major, coeffs ? coeffs + lo * m * 4 : nullptr,
                            secrets_only, secrets ? secrets + lo * 4 : nullptr, path ? path + lo : nullptr, flags ? flags + lo * fw : nullptr, sender_major ? B : 0);
    });
}
extern "C" int hbmpc_group_batch_recover(hbmpc_group *g, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B, const uint64_t *evals,
                                         uint64_t *coeffs, int32_t *path, uint64_t *flags) {
    return group_recover(g, n, d, t, S, sender_ids, B, evals, true, coeffs, false, nullptr, path, flags);
}
extern "C" int hbmpc_group_batch_recover_secrets(hbmpc_group *g, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B,
                                                 const uint64_t *evals, uint64_t *secrets, int32_t *path) {
    return group_recover(g, n, d, t, S, sender_ids, B, evals, true, nullptr, true, secrets, path, nullptr);
}
extern "C" int hbmpc_group_robust_interpolate_batch(hbmpc_group *g, size_t n, size_t d, size_t t, size_t S, const size_t *ids, size_t B, const uint64_t *shares,
                                                    uint64_t *coeffs, uint64_t *secrets, int32_t *path, uint64_t *flags) {
    return group_recover(g, n, d, t, S, ids, B, shares, false, coeffs, false, secrets, path, flags);
}

extern "C" int hbmpc_measure_imad_peak(hbmpc_ctx *ctx, int variant, double *giga_inst_per_s, double *elapsed_ms) {
    if (!ctx || variant < 0 || variant > 2 || !giga_inst_per_s) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    void *sink = nullptr;
    int rc = scratch_get(ctx, ctx->lanes[0], 8, 256, &sink);
    if (rc) return rc;
    const int iters = 4096, blocks = ctx->num_sms * 8, threads = 256;
    cudaStream_t st = ctx->main_stream();
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0, st));
        if (variant == 0) imad_probe_kernel<0><<<blocks, threads, 0, st>>>((unsigned int *)sink, 12345u + rep, iters);
        else if (variant == 1) imad_probe_kernel<1><<<blocks, threads, 0, st>>>((unsigned int *)sink, 12345u + rep, iters);
        else imad_probe_kernel<2><<<blocks, threads, 0, st>>>((unsigned int *)sink, 12345u + rep, iters);
        ctx->launches++;
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // thread-level instructions per thread per iteration: 64 in every variant (16 chains x 4, 2 x 4 lanes x 8, 8 x 8)
    double total = 64.0 * iters * (double)blocks * threads;
    *giga_inst_per_s = total / (best * 1e-3) / 1e9;
    if (elapsed_ms) *elapsed_ms = best;
    return HBMPC_SUCCESS;
}

// chains in {1, 2, 4, 8} independent carry chains per thread, warps_per_smsp resident warps per sub-partition (1 .. 16)
extern "C" int hbmpc_measure_wide_chains(hbmpc_ctx *ctx, int chains, int warps_per_smsp, double *giga_inst_per_s) {
    if (!ctx || !giga_inst_per_s || warps_per_smsp < 1 || warps_per_smsp > 16) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    void *sink = nullptr;
    int rc = scratch_get(ctx, ctx->lanes[0], 8, 256, &sink);
    if (rc) return rc;
    const int iters = 2048, blocks = ctx->num_sms * warps_per_smsp, threads = 128;   // one CTA = one warp per sub-partition
    cudaStream_t st = ctx->main_stream();
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0, st));
        switch (chains) {
            case 1: wide_chain_probe_kernel<1><<<blocks, threads, 0, st>>>((unsigned int *)sink, 12345u + rep, iters); break;
            case 2: wide_chain_probe_kernel<2><<<blocks, threads, 0, st>>>((unsigned int *)sink, 12345u + rep, iters); break;
            case 4: wide_chain_probe_kernel<4><<<blocks, threads, 0, st>>>((unsigned int *)sink, 12345u + rep, iters); break;
            case 8: wide_chain_probe_kernel<8><<<blocks, threads, 0, st>>>((unsigned int *)sink, 12345u + rep, iters); break;
            default: cudaEventDestroy(e0); cudaEventDestroy(e1); return HBMPC_INVALID_INPUT;
        }
        ctx->launches++;
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double total = 4.0 * 4 * chains * iters * (double)blocks * threads;   // IMAD.WIDE.U32.X per thread: 4 per chain call (one per mad.lo/madc.hi pair)
    *giga_inst_per_s = total / (best * 1e-3) / 1e9;
    return HBMPC_SUCCESS;
}

// Montgomery products per second (1e9) of a register-only loop: ilp = 1 (one serial chain of products per thread), 2 (two chains
// through mont_mul2), 4 (four chains through mont_mul), at warps_per_smsp resident warps per sub-partition
extern "C" int hbmpc_measure_mont_mul(hbmpc_ctx *ctx, int ilp, int warps_per_smsp, double *giga_products_per_s) {
    if (!ctx || !giga_products_per_s || warps_per_smsp < 1 || warps_per_smsp > 16) return HBMPC_INVALID_INPUT;
    cudaSetDevice(ctx->device);
    void *sink = nullptr;
    int rc = scratch_get(ctx, ctx->lanes[0], 8, 256, &sink);
    if (rc) return rc;
    const int iters = 1024, blocks = ctx->num_sms * warps_per_smsp, threads = 128;
    cudaStream_t st = ctx->main_stream();
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0, st));
        switch (ilp) {
            case 1: mont_mul_probe_kernel<1><<<blocks, threads, 0, st>>>((unsigned int *)sink, 12345u + rep, iters); break;
            case 2: mont_mul_probe_kernel<2><<<blocks, threads, 0, st>>>((unsigned int *)sink, 12345u + rep, iters); break;
            case 4: mont_mul_probe_kernel<4><<<blocks, threads, 0, st>>>((unsigned int *)sink, 12345u + rep, iters); break;
            default: cudaEventDestroy(e0); cudaEventDestroy(e1); return HBMPC_INVALID_INPUT;
        }
        ctx->launches++;
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *giga_products_per_s = (double)ilp * iters * (double)blocks * threads / (best * 1e-3) / 1e9;
    return HBMPC_SUCCESS;
}

#ifdef HB_ROBUST_PROF
// profiling builds only (tools/robust_prof.py): per-phase cycle sums of robust_kernel
extern "C" void hbmpc_debug_read_robust_prof(unsigned long long *out) {
    cudaMemcpyFromSymbol(out, hb::g_robust_prof, sizeof(unsigned long long) * 8);
}
#endif
