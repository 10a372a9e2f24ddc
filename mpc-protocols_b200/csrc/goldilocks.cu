// goldilocks.cu -- the share path over the reference's second field, GoldilocksField = Fp64<MontBackend<p = 2^64 - 2^32 + 1, generator 7>>
// (/root/reference/mpc/src/common/math/goldilocks.rs:4-13; the RandBit / PRandInt pipeline instantiates the same generic sharing code
// with it: honeybadger/mod.rs, honeybadger/preprocessing.rs).  SURVEY.md 8f N4 (tail).
//
// Same boundary conventions as hbmpc_b200.h: an element is its CANONICAL value (one uint64_t < p), arrays are dense and row-major in the
// reference's shapes, error codes are ShareErrorCode, every pointer may be a host or a device pointer.  Same mathematics as the Fr path:
// evaluation points are the elements of GeneralEvaluationDomain::new(n) (radix-2, size N = next_pow2(n), element(j) = w_N^j with
// w_N = g^((p-1)/N), g = 7: ark-ff derives TWO_ADIC_ROOT_OF_UNITY = g^((p-1)/2^32) = 1753635133440165772), shares are P(w_N^j),
// reconstruction interpolates through the lowest d+1 ids and checks the next t (batch_recover_secret's optimistic path,
// robust_interpolate.rs:343-428) or through all supplied points with a degree check (NonRobustShare::recover_secret, shamir.rs:199-239).
// With 8-byte elements these calls are HBM / PCIe bound, not multiplier bound: one kernel (a constant-matrix product with lazily
// accumulated 192-bit sums, one reduction per output) serves all of them.  The error-correcting decoder (OEC / Gao) is NOT
// instantiated for this field: chunks that fail the optimistic check are reported in path[] (-DecodingError) for the caller to hand
// to the reference's CPU decoder.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/hbmpc_b200.h"

extern "C" int hbmpc_ctx_device(const hbmpc_ctx *ctx);   // hbmpc.cu

namespace {

constexpr uint64_t GL_P = 0xffffffff00000001ull;
constexpr uint64_t GL_EPS = 0xffffffffull;             // 2^64 mod p
constexpr uint64_t GL_ROOT32 = 1753635133440165772ull;  // 7^((p-1)/2^32): generator of the 2^32-th roots of unity

// ---- host arithmetic (table construction only)
inline uint64_t h_mul(uint64_t a, uint64_t b) { return (uint64_t)((unsigned __int128)a * b % GL_P); }
inline uint64_t h_add(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a + b) % GL_P); }
inline uint64_t h_sub(uint64_t a, uint64_t b) { return a >= b ? a - b : a + (GL_P - b); }
inline uint64_t h_pow(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = h_mul(r, a);
        a = h_mul(a, a);
        e >>= 1;
    }
    return r;
}
inline uint64_t h_inv(uint64_t a) { return h_pow(a, GL_P - 2); }
inline int gl_domain_size(size_t n) {
    if (n == 0 || n > 256) return 0;
    int N = 1;
    while ((size_t)N < n) N <<= 1;
    return N;
}
inline std::vector<uint64_t> gl_domain(size_t n) {
    const int N = gl_domain_size(n);
    uint64_t w = GL_ROOT32;
    for (uint64_t k = (1ull << 32) / (uint64_t)N; k > 1; k >>= 1) w = h_mul(w, w);
    std::vector<uint64_t> x(n);
    uint64_t p = 1;
    for (size_t j = 0; j < n; ++j) { x[j] = p; p = h_mul(p, w); }
    return x;
}
// coefficients of the Lagrange basis polynomials of the points xs: Lc[k*m + i] = coefficient k of L_i
inline std::vector<uint64_t> gl_lagrange_coeffs(const std::vector<uint64_t> &xs) {
    const size_t m = xs.size();
    std::vector<uint64_t> A(m + 1, 0), Lc(m * m, 0);
    A[0] = 1;
    size_t deg = 0;
    for (size_t i = 0; i < m; ++i) {  // A *= (x - xs[i])
        const uint64_t nx = h_sub(0, xs[i]);
        A[deg + 1] = A[deg];
        for (size_t k = deg; k >= 1; --k) A[k] = h_add(A[k - 1], h_mul(A[k], nx));
        A[0] = h_mul(A[0], nx);
        ++deg;
    }
    for (size_t i = 0; i < m; ++i) {
        // q = A / (x - xs[i]) by synthetic division, denominator = q(xs[i])
        std::vector<uint64_t> q(m, 0);
        uint64_t carry = 0;
        for (size_t k = m; k-- > 0;) {
            q[k] = h_add(A[k + 1], carry);
            carry = h_mul(q[k], xs[i]);
        }
        uint64_t den = 0;
        for (size_t k = m; k-- > 0;) den = h_add(h_mul(den, xs[i]), q[k]);
        const uint64_t w = h_inv(den);
        for (size_t k = 0; k < m; ++k) Lc[k * m + i] = h_mul(q[k], w);
    }
    return Lc;
}

// ---- device arithmetic
__device__ __forceinline__ uint64_t gl_reduce128(uint64_t lo, uint64_t hi) {
    // 2^64 = 2^32 - 1, 2^96 = -1 (mod p):  lo + hi_lo * (2^32 - 1) - hi_hi
    const uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    uint64_t t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS;               // borrow: + 2^64 = + (2^32 - 1) too much was added, i.e. subtract EPS
    const uint64_t t1 = hi_lo * GL_EPS;         // < 2^64
    uint64_t r = t0 + t1;
    if (r < t1) r += GL_EPS;                    // carry: 2^64 = EPS
    if (r >= GL_P) r -= GL_P;
    return r;
}
__device__ __forceinline__ uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128(a * b, __umul64hi(a, b)); }
__device__ __forceinline__ uint64_t gl_add(uint64_t a, uint64_t b) {
    uint64_t r = a + b;
    if (r < a || r >= GL_P) r -= GL_P;
    return r;
}
__device__ __forceinline__ uint64_t gl_sub(uint64_t a, uint64_t b) { return a >= b ? a - b : a + (GL_P - b); }

struct GlMatArgs {
    const uint64_t *M;       // [R][C] canonical
    const uint64_t *in;      // element (b, c) at in[b*in_sb + col_map[c]*in_sc]
    uint64_t *out;           // element (b, r) at out[b*out_sb + r*out_sr], rows n_chk .. R-1 (row r stores to index r - n_chk)
    long long B, in_sb, in_sc, out_sb, out_sr;
    int R, C, n_chk;
    const int *col_map;      // [C] or nullptr
    const int *chk_map;      // [n_chk]: check row r must equal in[b][chk_map[r]] (< 0: must equal zero)
    unsigned char *fail;     // [B] set to 1 when a check row disagrees
    unsigned int *err;       // set to 1 when a non-canonical input is seen
};
// out[b][r] = sum_c M[r][c] * in[b][c]: one thread per (b, r); the 128-bit products are summed in 192 bits and reduced once.
__global__ void __launch_bounds__(256) gl_matvec_kernel(const GlMatArgs a) {
    extern __shared__ uint64_t sM[];
    for (int i = threadIdx.x; i < a.R * a.C; i += blockDim.x) sM[i] = a.M[i];
    __syncthreads();
    const long long total = a.B * a.R;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / a.R;
        const int r = (int)(idx - b * a.R);
        const uint64_t *row = sM + (size_t)r * a.C;
        const uint64_t *x = a.in + b * a.in_sb;
        uint64_t lo = 0, mid = 0, hi = 0;
        bool bad = false;
        for (int c = 0; c < a.C; ++c) {
            const uint64_t v = x[(long long)(a.col_map ? a.col_map[c] : c) * a.in_sc];
            bad |= v >= GL_P;
            const uint64_t pl = row[c] * v, ph = __umul64hi(row[c], v);
            lo += pl;
            const uint64_t c0 = lo < pl ? 1ull : 0ull;
            mid += ph;
            const uint64_t c1 = mid < ph ? 1ull : 0ull;
            mid += c0;
            hi += c1 + (mid < c0 ? 1ull : 0ull);
        }
        if (bad) *(volatile unsigned int *)a.err = 1u;
        // hi * 2^128 = -hi * 2^32 (mod p), hi < 2^9
        const uint64_t res = gl_sub(gl_reduce128(lo, mid), gl_mul(hi, 1ull << 32));
        if (r < a.n_chk) {
            const int cm = a.chk_map[r];
            const uint64_t want = cm < 0 ? 0ull : x[(long long)cm * a.in_sc];
            if (cm >= 0 && want >= GL_P) *(volatile unsigned int *)a.err = 1u;
            if (res != want) a.fail[b] = 1;
        } else {
            a.out[b * a.out_sb + (long long)(r - a.n_chk) * a.out_sr] = res;
        }
    }
}
__global__ void __launch_bounds__(256) gl_elementwise_kernel(int op, long long count, const uint64_t *a, const uint64_t *b, uint64_t *out, unsigned int *err) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const uint64_t x = a[i], y = b[i];
        if (x >= GL_P || y >= GL_P) *(volatile unsigned int *)err = 1u;
        out[i] = op == 0 ? gl_add(x, y) : (op == 1 ? gl_sub(x, y) : gl_mul(x, y));
    }
}
// path[b] = 0 or -DecodingError; failing items' outputs are zeroed; status words as in the Fr path
__global__ void gl_finish_kernel(long long B, int m, const unsigned char *fail, uint64_t *coeffs, int *path, uint64_t *secrets, int code, unsigned int *undec) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        const bool f = fail[b] != 0;
        if (f) {
            for (int k = 0; k < m; ++k) coeffs[b * m + k] = 0;
            *(volatile unsigned int *)undec = 1u;
        }
        if (path) path[b] = f ? -code : 0;
        if (secrets) secrets[b] = f ? 0ull : coeffs[b * m];
    }
}
// NonRobustShare::recover_secret epilogue: status[b] = degree of the interpolant, or -DegreeMismatch
__global__ void gl_degree_kernel(long long B, int m, const unsigned char *fail, uint64_t *coeffs, int *status, uint64_t *secrets) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        if (fail[b]) {
            for (int k = 0; k < m; ++k) coeffs[b * m + k] = 0;
            status[b] = -HBMPC_DEGREE_MISMATCH;
            if (secrets) secrets[b] = 0;
            continue;
        }
        int deg = 0;
        for (int k = m - 1; k > 0; --k)
            if (coeffs[b * m + k]) { deg = k; break; }
        status[b] = deg;
        if (secrets) secrets[b] = coeffs[b * m];
    }
}

// ---- a small context of its own per device (stream, scratch, status word): the Goldilocks calls share nothing with the Fr kernels
struct GlState {
    cudaStream_t st = nullptr;
    unsigned int *d_status = nullptr, *h_status = nullptr;
    void *scratch[6] = {};
    size_t cap[6] = {};
};
std::mutex g_mu;
std::map<const hbmpc_ctx *, GlState> g_state;

bool dev_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}
int gl_get(const hbmpc_ctx *ctx, GlState **out) {
    if (cudaSetDevice(hbmpc_ctx_device(ctx)) != cudaSuccess) { cudaGetLastError(); return HBMPC_NO_DEVICE; }
    std::lock_guard<std::mutex> lk(g_mu);
    GlState &s = g_state[ctx];
    if (!s.st) {
        if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess) return HBMPC_CUDA_ERROR;
        if (cudaHostAlloc((void **)&s.h_status, 16, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void **)&s.d_status, s.h_status, 0) != cudaSuccess) return HBMPC_CUDA_ERROR;
        memset(s.h_status, 0, 16);
    }
    *out = &s;
    return 0;
}
int gl_scratch(GlState *s, int slot, size_t bytes, void **out) {
    if (s->cap[slot] < bytes) {
        if (s->scratch[slot]) { cudaStreamSynchronize(s->st); cudaFree(s->scratch[slot]); s->scratch[slot] = nullptr; s->cap[slot] = 0; }
        const size_t cap = bytes + bytes / 8 + 256;
        if (cudaMalloc(&s->scratch[slot], cap) != cudaSuccess) { cudaGetLastError(); return HBMPC_CUDA_ERROR; }
        s->cap[slot] = cap;
    }
    *out = s->scratch[slot];
    return 0;
}
// stage a user array on the device (identity for device pointers)
int gl_in(GlState *s, int slot, const void *user, size_t bytes, const void **dev) {
    if (dev_ptr(user)) { *dev = user; return 0; }
    void *d = nullptr;
    int rc = gl_scratch(s, slot, bytes, &d);
    if (rc) return rc;
    if (cudaMemcpyAsync(d, user, bytes, cudaMemcpyHostToDevice, s->st) != cudaSuccess) return HBMPC_CUDA_ERROR;
    *dev = d;
    return 0;
}
int gl_out(GlState *s, int slot, void *user, size_t bytes, void **dev) {
    if (!user) { *dev = nullptr; return 0; }
    if (dev_ptr(user)) { *dev = user; return 0; }
    return gl_scratch(s, slot, bytes, dev);
}
int gl_back(GlState *s, void *user, const void *dev, size_t bytes) {
    if (!user || user == dev) return 0;
    return cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, s->st) == cudaSuccess ? 0 : HBMPC_CUDA_ERROR;
}
int gl_status(GlState *s) {
    if (cudaStreamSynchronize(s->st) != cudaSuccess) { cudaGetLastError(); return HBMPC_CUDA_ERROR; }
    volatile unsigned int *hs = s->h_status;
    const unsigned int bad = hs[0], undec = hs[2];
    hs[0] = 0; hs[2] = 0;
    if (bad) return HBMPC_INVALID_INPUT;
    if (undec) return HBMPC_DECODING_ERROR;
    return HBMPC_SUCCESS;
}
int gl_launch_matvec(GlState *s, GlMatArgs a) {
    a.err = s->d_status;
    const size_t smem = (size_t)a.R * a.C * 8;
    if (smem > 200 * 1024) return HBMPC_INVALID_INPUT;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(gl_matvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return HBMPC_CUDA_ERROR;
    const long long total = a.B * a.R;
    const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
    gl_matvec_kernel<<<grid, 256, smem, s->st>>>(a);
    return cudaGetLastError() == cudaSuccess ? 0 : HBMPC_CUDA_ERROR;
}
// out = V(n x cols) * in over the domain
int gl_apply(hbmpc_ctx *ctx, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *out, int recipient_major) {
    GlState *s = nullptr;
    int rc = gl_get(ctx, &s);
    if (rc) return rc;
    std::vector<uint64_t> x = gl_domain(n), V(n * cols);
    for (size_t j = 0; j < n; ++j) {
        uint64_t p = 1;
        for (size_t k = 0; k < cols; ++k) { V[j * cols + k] = p; p = h_mul(p, x[j]); }
    }
    void *dV = nullptr;
    if ((rc = gl_scratch(s, 0, V.size() * 8, &dV))) return rc;
    if (cudaMemcpyAsync(dV, V.data(), V.size() * 8, cudaMemcpyHostToDevice, s->st) != cudaSuccess) return HBMPC_CUDA_ERROR;
    if (cudaStreamSynchronize(s->st) != cudaSuccess) return HBMPC_CUDA_ERROR;   // V is freed on return
    const void *din = nullptr;
    void *dout = nullptr;
    if ((rc = gl_in(s, 1, in, B * cols * 8, &din))) return rc;
    if ((rc = gl_out(s, 2, out, B * n * 8, &dout))) return rc;
    GlMatArgs a{};
    a.M = (const uint64_t *)dV; a.in = (const uint64_t *)din; a.out = (uint64_t *)dout;
    a.B = (long long)B; a.R = (int)n; a.C = (int)cols;
    a.in_sb = (long long)cols; a.in_sc = 1;
    a.out_sb = recipient_major ? 1 : (long long)n;
    a.out_sr = recipient_major ? (long long)B : 1;
    if ((rc = gl_launch_matvec(s, a))) return rc;
    if ((rc = gl_back(s, out, dout, B * n * 8))) return rc;
    return gl_status(s);
}
// shared body of batch_recover (check_from = d+1 .. d+t+1 checked, first d+1 interpolated) and of the non-robust recovery
// (interpolation through ALL points: coefficient rows above `deg` are zero checks)
int gl_recover(hbmpc_ctx *ctx, size_t n, size_t deg, size_t n_examined, bool all_points, size_t S, const size_t *ids, size_t B, const uint64_t *in,
               bool sender_major, uint64_t *coeffs, uint64_t *secrets, int32_t *path_or_status) {
    GlState *s = nullptr;
    int rc = gl_get(ctx, &s);
    if (rc) return rc;
    std::vector<int> order(S);
    for (size_t i = 0; i < S; ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ids[a] < ids[b]; });
    for (size_t i = 0; i < S; ++i) {
        if (ids[order[i]] >= n) return HBMPC_INVALID_INPUT;
        if (i > 0 && ids[order[i]] == ids[order[i - 1]]) return HBMPC_INVALID_INPUT;
    }
    const std::vector<uint64_t> dom = gl_domain(n);
    const size_t m = deg + 1;
    std::vector<uint64_t> M;
    std::vector<int> col_map, chk_map;
    size_t C, n_chk;
    if (!all_points) {
        // lowest d+1 ids interpolate, the next n_examined - (d+1) ids are checked: rows = [L_i(x_s) for the checked s; Lc]
        std::vector<uint64_t> xs(m);
        for (size_t i = 0; i < m; ++i) xs[i] = dom[ids[order[i]]];
        const std::vector<uint64_t> Lc = gl_lagrange_coeffs(xs);
        C = m; n_chk = n_examined - m;
        M.assign((n_chk + m) * C, 0);
        for (size_t r = 0; r < n_chk; ++r) {
            const uint64_t x = dom[ids[order[m + r]]];
            for (size_t i = 0; i < m; ++i) {
                uint64_t acc = 0;
                for (size_t k = m; k-- > 0;) acc = h_add(h_mul(acc, x), Lc[k * m + i]);
                M[r * C + i] = acc;
            }
            chk_map.push_back(order[m + r]);
        }
        for (size_t k = 0; k < m; ++k)
            for (size_t i = 0; i < m; ++i) M[(n_chk + k) * C + i] = Lc[k * m + i];
        for (size_t i = 0; i < m; ++i) col_map.push_back(order[i]);
    } else {
        // interpolation through all S points: coefficient rows deg+1 .. S-1 must vanish (check rows against zero), rows 0 .. deg are stored
        std::vector<uint64_t> xs(S);
        for (size_t i = 0; i < S; ++i) xs[i] = dom[ids[order[i]]];
        const std::vector<uint64_t> Lc = gl_lagrange_coeffs(xs);
        C = S; n_chk = S - m;
        M.assign(S * C, 0);
        for (size_t r = 0; r < n_chk; ++r) {
            for (size_t i = 0; i < S; ++i) M[r * C + i] = Lc[(m + r) * S + i];
            chk_map.push_back(-1);
        }
        for (size_t k = 0; k < m; ++k)
            for (size_t i = 0; i < S; ++i) M[(n_chk + k) * C + i] = Lc[k * S + i];
        for (size_t i = 0; i < S; ++i) col_map.push_back(order[i]);
    }
    std::vector<int> maps(col_map);
    maps.insert(maps.end(), chk_map.begin(), chk_map.end());
    void *dM = nullptr, *dmaps = nullptr, *dfail = nullptr;
    if ((rc = gl_scratch(s, 0, M.size() * 8, &dM))) return rc;
    if ((rc = gl_scratch(s, 4, maps.size() * 4 + 16, &dmaps))) return rc;
    if ((rc = gl_scratch(s, 5, B + 16, &dfail))) return rc;
    if (cudaMemcpyAsync(dM, M.data(), M.size() * 8, cudaMemcpyHostToDevice, s->st) != cudaSuccess) return HBMPC_CUDA_ERROR;
    if (cudaMemcpyAsync(dmaps, maps.data(), maps.size() * 4, cudaMemcpyHostToDevice, s->st) != cudaSuccess) return HBMPC_CUDA_ERROR;
    if (cudaMemsetAsync(dfail, 0, B, s->st) != cudaSuccess) return HBMPC_CUDA_ERROR;
    if (cudaStreamSynchronize(s->st) != cudaSuccess) return HBMPC_CUDA_ERROR;   // M / maps are freed on return
    const void *din = nullptr;
    void *dco = nullptr, *dsec = nullptr, *dps = nullptr;
    if ((rc = gl_in(s, 1, in, B * S * 8, &din))) return rc;
    if ((rc = gl_out(s, 2, coeffs, B * m * 8, &dco))) return rc;
    if (!dco && (rc = gl_scratch(s, 2, B * m * 8, &dco))) return rc;   // secrets-only callers
    void *dsec_user = nullptr;
    if (secrets) {
        if (dev_ptr(secrets)) dsec = secrets;
        else {
            // secrets share scratch slot 3 with the path / status words: [B] u64 then [B] i32
            if ((rc = gl_scratch(s, 3, B * 12 + 64, &dsec_user))) return rc;
            dsec = dsec_user;
        }
    }
    if (dev_ptr(path_or_status)) dps = path_or_status;
    else {
        if (!dsec_user && (rc = gl_scratch(s, 3, B * 12 + 64, &dsec_user))) return rc;
        dps = (char *)dsec_user + B * 8;
    }
    GlMatArgs a{};
    a.M = (const uint64_t *)dM; a.in = (const uint64_t *)din; a.out = (uint64_t *)dco;
    a.B = (long long)B; a.R = (int)(n_chk + m); a.C = (int)C; a.n_chk = (int)n_chk;
    a.in_sb = sender_major ? 1 : (long long)S;
    a.in_sc = sender_major ? (long long)B : 1;
    a.out_sb = (long long)m; a.out_sr = 1;
    a.col_map = (const int *)dmaps; a.chk_map = (const int *)dmaps + col_map.size();
    a.fail = (unsigned char *)dfail;
    if ((rc = gl_launch_matvec(s, a))) return rc;
    if (all_points) gl_degree_kernel<<<148 * 4, 256, 0, s->st>>>((long long)B, (int)m, (const unsigned char *)dfail, (uint64_t *)dco, (int *)dps, (uint64_t *)dsec);
    else gl_finish_kernel<<<148 * 4, 256, 0, s->st>>>((long long)B, (int)m, (const unsigned char *)dfail, (uint64_t *)dco, (int *)dps, (uint64_t *)dsec, HBMPC_DECODING_ERROR, s->d_status + 2);
    if (cudaGetLastError() != cudaSuccess) return HBMPC_CUDA_ERROR;
    if ((rc = gl_back(s, coeffs, dco, B * m * 8))) return rc;
    if ((rc = gl_back(s, secrets, dsec, B * 8))) return rc;
    if ((rc = gl_back(s, path_or_status, dps, B * 4))) return rc;
    return gl_status(s);
}

}  // namespace

extern "C" int hbmpc_gl_compute_shares_batch(hbmpc_ctx *ctx, size_t n, size_t d, size_t B, const uint64_t *coeffs, uint64_t *shares) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (n <= d) return HBMPC_INVALID_INPUT;
    if (!gl_domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;
    if (B == 0) return HBMPC_SUCCESS;
    if (!coeffs || !shares) return HBMPC_INVALID_INPUT;
    return gl_apply(ctx, n, d + 1, B, coeffs, shares, 0);
}
extern "C" int hbmpc_gl_apply_vandermonde_batch(hbmpc_ctx *ctx, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *out, int recipient_major) {
    if (!ctx || cols == 0 || cols > 256) return HBMPC_INVALID_INPUT;
    if (!gl_domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;
    if (B == 0) return HBMPC_SUCCESS;
    if (!in || !out) return HBMPC_INVALID_INPUT;
    return gl_apply(ctx, n, cols, B, in, out, recipient_major);
}
extern "C" int hbmpc_gl_batch_recover(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B, const uint64_t *evals,
                                      uint64_t *coeffs, uint64_t *secrets, int32_t *path) {
    if (!ctx) return HBMPC_INVALID_INPUT;
    if (n < 3 * t + 1 || S == 0 || !sender_ids || B == 0 || S > 256) return HBMPC_INVALID_INPUT;   // robust_interpolate.rs:290-341
    if (S < d + t + 1) return HBMPC_INVALID_INPUT;
    if (!gl_domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;
    if (!evals || !path || (!coeffs && !secrets)) return HBMPC_INVALID_INPUT;
    return gl_recover(ctx, n, d, d + t + 1, false, S, sender_ids, B, evals, true, coeffs, secrets, path);
}
extern "C" int hbmpc_gl_nonrobust_recover_batch(hbmpc_ctx *ctx, size_t n, size_t deg, size_t S, const size_t *ids, size_t B, const uint64_t *shares,
                                                int sender_major, uint64_t *coeffs, uint64_t *secrets, int32_t *status) {
    if (!ctx || S == 0 || !ids || S > 256) return HBMPC_INVALID_INPUT;            // shamir.rs:204-216
    if (S < deg + 1) return HBMPC_INSUFFICIENT_SHARES;
    if (!gl_domain_size(n)) return HBMPC_NO_SUITABLE_DOMAIN;
    if (B == 0) return HBMPC_SUCCESS;
    if (!shares || !coeffs || !status) return HBMPC_INVALID_INPUT;
    return gl_recover(ctx, n, deg, S, true, S, ids, B, shares, sender_major != 0, coeffs, secrets, status);
}
extern "C" int hbmpc_gl_elementwise(hbmpc_ctx *ctx, int op, size_t count, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    if (!ctx || op < 0 || op > 2) return HBMPC_INVALID_INPUT;
    if (count == 0) return HBMPC_SUCCESS;
    if (!a || !b || !out) return HBMPC_INVALID_INPUT;
    GlState *s = nullptr;
    int rc = gl_get(ctx, &s);
    if (rc) return rc;
    const void *da = nullptr, *db = nullptr;
    void *dout = nullptr;
    if ((rc = gl_in(s, 0, a, count * 8, &da))) return rc;
    if ((rc = gl_in(s, 1, b, count * 8, &db))) return rc;
    if ((rc = gl_out(s, 2, out, count * 8, &dout))) return rc;
    gl_elementwise_kernel<<<148 * 8, 256, 0, s->st>>>(op, (long long)count, (const uint64_t *)da, (const uint64_t *)db, (uint64_t *)dout, s->d_status);
    if (cudaGetLastError() != cudaSuccess) return HBMPC_CUDA_ERROR;
    if ((rc = gl_back(s, out, dout, count * 8))) return rc;
    return gl_status(s);
}
// releases the Goldilocks state of a context (called by hbmpc_ctx_destroy)
extern "C" void hbmpc_gl_release(const hbmpc_ctx *ctx) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_state.find(ctx);
    if (it == g_state.end()) return;
    GlState &s = it->second;
    if (s.st) { cudaStreamSynchronize(s.st); cudaStreamDestroy(s.st); }
    for (void *p : s.scratch) if (p) cudaFree(p);
    if (s.h_status) cudaFreeHost(s.h_status);
    g_state.erase(it);
}
