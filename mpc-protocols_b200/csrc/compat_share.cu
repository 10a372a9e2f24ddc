// compat_share.cu -- the reference's single-item share C ABI (include/hbmpc_compat_share.h) as thin wrappers over the batch entry
// points: every field operation on a share runs in the CUDA kernels (B = 1 batches); the host only draws the random polynomial,
// validates arguments in the reference's order and -- for the x = id Shamir scheme -- builds the small Vandermonde / Lagrange
// matrices of the call's id set (table set-up, like the domain tables of hbmpc.cu).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <set>
#include <vector>

#include "../../include/hbmpc_b200.h"
#include "../../include/hbmpc_compat_share.h"
#include "tables.hpp"

using namespace hb;

namespace {

thread_local int g_last_status = 0;
std::mutex g_mu;          // the process-wide context is thread-compatible, not thread-safe
hbmpc_ctx *g_ctx = nullptr;

hbmpc_ctx *context() {
    if (!g_ctx) {
        int rc = hbmpc_ctx_create(0, &g_ctx);
        if (rc != HBMPC_SUCCESS) {
            g_last_status = rc;
            g_ctx = nullptr;
        }
    }
    return g_ctx;
}
ShareErrorCode status_to_code(int rc) {
    g_last_status = rc;
    if (rc >= 0 && rc <= 8) return (ShareErrorCode)rc;
    return PolynomialOperationError;  // NO_DEVICE / CUDA_ERROR: the reference has no such variant
}

struct Boxed { uint64_t v[4]; };  // FieldOpaque: canonical value
FieldOpaque *box(const uint64_t *v) {
    Boxed *b = (Boxed *)std::malloc(sizeof(Boxed));
    std::memcpy(b->v, v, 32);
    return (FieldOpaque *)b;
}
const uint64_t *unbox(const FieldOpaque *p) { return ((const Boxed *)p)->v; }

// Fp::rand as arkworks draws it: four 64-bit words, top bit cleared, rejected while >= r.  The accepted words are a
// uniform residue; which representation they are read in does not change the distribution.
void random_fr(std::mt19937_64 &g, uint64_t *out) {
    for (;;) {
        for (int i = 0; i < 4; ++i) out[i] = g();
        out[3] &= ~0ULL >> 1;
        if (!hfr::geq_mod(out)) return;
    }
}
std::mt19937_64 &thread_rng() {
    thread_local std::mt19937_64 g{std::random_device{}()};
    return g;
}

template <typename Share>
void free_share(Share s) {
    if (s.share) std::free(s.share);
}
template <typename Slice>
void free_share_slice(Slice s) {
    if (!s.pointer) return;
    for (uintptr_t i = 0; i < s.len; ++i) std::free(s.pointer[i].share);
    std::free(s.pointer);
}
ByteSlice bytes_of(const uint64_t *limbs, bool be) {
    uint8_t *p = (uint8_t *)std::malloc(32);
    for (int i = 0; i < 32; ++i) {
        uint8_t byte = (uint8_t)(limbs[i / 8] >> (8 * (i % 8)));
        p[be ? 31 - i : i] = byte;
    }
    return ByteSlice{p, 32};
}
// normalised coefficient vector (DensePolynomial drops trailing zero coefficients: robust_interpolate.rs:149, shamir.rs:238)
void emit_coeffs(const std::vector<uint64_t> &c, size_t m, U256Slice *out) {
    size_t len = m;
    while (len > 0 && !(c[4 * (len - 1)] | c[4 * (len - 1) + 1] | c[4 * (len - 1) + 2] | c[4 * (len - 1) + 3])) --len;
    out->pointer = (U256 *)std::malloc(std::max<size_t>(len, 1) * sizeof(U256));
    out->len = len;
    std::memcpy(out->pointer, c.data(), len * 32);
}

// shares of one secret on the domain (RobustShare / NonRobustShare::compute_shares): K1 with B = 1
template <typename Share, typename Slice>
ShareErrorCode domain_shares(U256 secret, uintptr_t degree, uintptr_t n, Slice *out) {
    if (!out) return status_to_code(HBMPC_INVALID_INPUT);
    if (hfr::geq_mod(secret.data)) return status_to_code(HBMPC_INVALID_INPUT);   // from_bigint(..).unwrap() panics in the reference
    if (n <= degree) return status_to_code(HBMPC_INVALID_INPUT);
    std::lock_guard<std::mutex> lk(g_mu);
    hbmpc_ctx *ctx = context();
    if (!ctx) return status_to_code(g_last_status);
    std::vector<uint64_t> coeffs(4 * (degree + 1)), shares(4 * n);
    for (uintptr_t k = 0; k <= degree; ++k) random_fr(thread_rng(), &coeffs[4 * k]);   // d+1 draws, the first overwritten (:68-69)
    std::memcpy(coeffs.data(), secret.data, 32);
    int rc = hbmpc_compute_shares_batch(ctx, n, degree, 1, coeffs.data(), shares.data());
    if (rc) return status_to_code(rc);
    out->pointer = (Share *)std::malloc(n * sizeof(Share));
    out->len = n;
    for (uintptr_t j = 0; j < n; ++j) out->pointer[j] = Share{box(&shares[4 * j]), j, degree};
    return status_to_code(HBMPC_SUCCESS);
}

}  // namespace

extern "C" {

int hbmpc_compat_last_status(void) { return g_last_status; }

void free_u256_slice(U256Slice s) { std::free(s.pointer); }
void free_bytes_slice(ByteSlice s) { std::free(s.pointer); }
U256 be_bytes_to_u256(ByteSlice b) {
    U256 u{};
    for (uintptr_t i = 0; i < b.len && i < 32; ++i) u.data[i / 8] |= (uint64_t)b.pointer[b.len - 1 - i] << (8 * (i % 8));
    return u;
}
U256 le_bytes_to_u256(ByteSlice b) {
    U256 u{};
    for (uintptr_t i = 0; i < b.len && i < 32; ++i) u.data[i / 8] |= (uint64_t)b.pointer[i] << (8 * (i % 8));
    return u;
}
ByteSlice u256_to_be_bytes(U256 n) { return bytes_of(n.data, true); }
ByteSlice u256_to_le_bytes(U256 n) { return bytes_of(n.data, false); }
ByteSlice field_ptr_to_bytes(FieldOpaque *f, bool be) { return bytes_of(unbox(f), be); }

void free_shamir_share(ShamirShare s) { free_share(s); }
void free_robust_share(RobustShare s) { free_share(s); }
void free_non_robust_share(NonRobustShare s) { free_share(s); }
void free_shamir_share_slice(ShamirShareSlice s) { free_share_slice(s); }
void free_robust_share_slice(RobustShareSlice s) { free_share_slice(s); }
void free_non_robust_share_slice(NonRobustShareSlice s) { free_share_slice(s); }

ShamirShare shamir_share_new(U256 secret, uintptr_t id, uintptr_t degree, FieldKind) { return ShamirShare{box(secret.data), id, degree}; }
RobustShare robust_share_new(U256 secret, uintptr_t id, uintptr_t degree, FieldKind) { return RobustShare{box(secret.data), id, degree}; }
NonRobustShare non_robust_share_new(U256 secret, uintptr_t id, uintptr_t degree, FieldKind) { return NonRobustShare{box(secret.data), id, degree}; }

// Shamirshare::compute_shares (shamir.rs:43-89): P(x = id), ids unique and non-zero; the evaluation is the Vandermonde matrix of the
// ids applied on the device (hbmpc_apply_matrix_batch, B = 1)
ShareErrorCode shamir_share_compute_shares(U256 secret, uintptr_t degree, const UsizeSlice *ids, FieldKind, ShamirShareSlice *out) {
    if (!ids || !out) return status_to_code(HBMPC_INVALID_INPUT);
    if (hfr::geq_mod(secret.data)) return status_to_code(HBMPC_INVALID_INPUT);
    const uintptr_t k = ids->len;
    if (k < degree + 1) return status_to_code(HBMPC_INSUFFICIENT_SHARES);
    std::set<uintptr_t> seen;
    for (uintptr_t i = 0; i < k; ++i)
        if (ids->pointer[i] == 0) return status_to_code(HBMPC_INVALID_INPUT);
    for (uintptr_t i = 0; i < k; ++i)
        if (!seen.insert(ids->pointer[i]).second) return status_to_code(HBMPC_INVALID_INPUT);
    if (k > 256 || degree + 1 > 256) return status_to_code(HBMPC_INVALID_INPUT);
    std::lock_guard<std::mutex> lk(g_mu);
    hbmpc_ctx *ctx = context();
    if (!ctx) return status_to_code(g_last_status);
    const size_t m = degree + 1;
    std::vector<uint64_t> coeffs(4 * m), V(4 * k * m), shares(4 * k);
    for (size_t c = 0; c < m; ++c) random_fr(thread_rng(), &coeffs[4 * c]);
    std::memcpy(coeffs.data(), secret.data, 32);
    for (uintptr_t i = 0; i < k; ++i) {   // row i = powers of x_i = id_i (table set-up)
        HFr x = hfr::from_u64((uint64_t)ids->pointer[i]), p = hfr::ONE;
        for (size_t c = 0; c < m; ++c) {
            hfr::to_canon(p, &V[4 * (i * m + c)]);
            p = hfr::mul(p, x);
        }
    }
    int rc = hbmpc_apply_matrix_batch(ctx, k, m, V.data(), 1, coeffs.data(), shares.data(), 0);
    if (rc) return status_to_code(rc);
    out->pointer = (ShamirShare *)std::malloc(k * sizeof(ShamirShare));
    out->len = k;
    for (uintptr_t i = 0; i < k; ++i) out->pointer[i] = ShamirShare{box(&shares[4 * i]), ids->pointer[i], degree};
    return status_to_code(HBMPC_SUCCESS);
}

// Shamirshare::recover_secret (shamir.rs:92-126): Lagrange interpolation through ALL supplied points, DegreeMismatch when the
// interpolant exceeds the claimed degree.  The coefficient-form basis of the id set is built here, applied on the device.
ShareErrorCode shamir_share_recover_secret(ShamirShareSlice shares, U256 *output_secret, U256Slice *output_coeffs, FieldKind) {
    if (!output_secret || !output_coeffs) return status_to_code(HBMPC_INVALID_INPUT);
    const uintptr_t k = shares.len;
    if (k == 0 || !shares.pointer) return status_to_code(HBMPC_INVALID_INPUT);
    std::set<uintptr_t> seen;
    for (uintptr_t i = 0; i < k; ++i)
        if (!seen.insert(shares.pointer[i].id).second) return status_to_code(HBMPC_INVALID_INPUT);
    const uintptr_t deg = shares.pointer[0].degree;
    for (uintptr_t i = 0; i < k; ++i)
        if (shares.pointer[i].degree != deg) return status_to_code(HBMPC_DEGREE_MISMATCH);
    if (k < deg + 1) return status_to_code(HBMPC_INSUFFICIENT_SHARES);
    for (uintptr_t i = 0; i < k; ++i)
        if (shares.pointer[i].id == 0) return status_to_code(HBMPC_INVALID_INPUT);
    if (k > 256) return status_to_code(HBMPC_INVALID_INPUT);
    std::lock_guard<std::mutex> lk(g_mu);
    hbmpc_ctx *ctx = context();
    if (!ctx) return status_to_code(g_last_status);
    std::vector<HFr> xs(k);
    for (uintptr_t i = 0; i < k; ++i) xs[i] = hfr::from_u64((uint64_t)shares.pointer[i].id);
    Lagrange L = lagrange_basis(xs);                 // Lc[c*k + i]: coefficient c of the basis polynomial of point i
    std::vector<uint64_t> M(4 * k * k), y(4 * k), c(4 * k);
    for (size_t e = 0; e < (size_t)k * k; ++e) hfr::to_canon(L.Lc[e], &M[4 * e]);
    for (uintptr_t i = 0; i < k; ++i) std::memcpy(&y[4 * i], unbox(shares.pointer[i].share), 32);
    int rc = hbmpc_apply_matrix_batch(ctx, k, k, M.data(), 1, y.data(), c.data(), 0);
    if (rc) return status_to_code(rc);
    for (size_t q = deg + 1; q < k; ++q)
        if (c[4 * q] | c[4 * q + 1] | c[4 * q + 2] | c[4 * q + 3]) return status_to_code(HBMPC_DEGREE_MISMATCH);
    std::memcpy(output_secret->data, c.data(), 32);
    emit_coeffs(c, deg + 1, output_coeffs);
    return status_to_code(HBMPC_SUCCESS);
}

ShareErrorCode robust_share_compute_shares(U256 secret, uintptr_t degree, uintptr_t n, RobustShareSlice *out, FieldKind) {
    return domain_shares<RobustShare>(secret, degree, n, out);
}
ShareErrorCode non_robust_share_compute_shares(U256 secret, uintptr_t degree, uintptr_t n, NonRobustShareSlice *out, FieldKind) {
    return domain_shares<NonRobustShare>(secret, degree, n, out);
}

// RobustShare::recover_secret (robust_interpolate.rs:94-157) through K4 with B = 1
ShareErrorCode robust_share_recover_secret(RobustShareSlice shares, uintptr_t n, uintptr_t t, U256 *output_secret, U256Slice *output_coeffs, FieldKind) {
    if (!output_secret || !output_coeffs) return status_to_code(HBMPC_INVALID_INPUT);
    const uintptr_t S = shares.len;
    if (n < 3 * t + 1 || S == 0 || !shares.pointer) return status_to_code(HBMPC_INVALID_INPUT);   // :100-110
    const uintptr_t deg = shares.pointer[0].degree;
    for (uintptr_t i = 0; i < S; ++i)
        if (shares.pointer[i].degree != deg) return status_to_code(HBMPC_DEGREE_MISMATCH);        // :116-120
    std::vector<size_t> ids(S);
    std::vector<uint64_t> y(4 * S), c(4 * (deg + 1)), sec(4);
    for (uintptr_t i = 0; i < S; ++i) {
        ids[i] = shares.pointer[i].id;
        std::memcpy(&y[4 * i], unbox(shares.pointer[i].share), 32);
    }
    std::lock_guard<std::mutex> lk(g_mu);
    hbmpc_ctx *ctx = context();
    if (!ctx) return status_to_code(g_last_status);
    int32_t path = 0;
    int rc = hbmpc_robust_interpolate_batch(ctx, n, deg, t, S, ids.data(), 1, y.data(), c.data(), sec.data(), &path, nullptr);
    if (rc) return status_to_code(rc);
    if (path < 0) return status_to_code(-path);
    std::memcpy(output_secret->data, sec.data(), 32);
    emit_coeffs(c, deg + 1, output_coeffs);
    return status_to_code(HBMPC_SUCCESS);
}

// NonRobustShare::recover_secret (shamir.rs:199-239) through the a10 entry point with B = 1
ShareErrorCode non_robust_share_recover_secret(NonRobustShareSlice shares, uintptr_t n, U256 *output_secret, U256Slice *output_coeffs, FieldKind) {
    if (!output_secret || !output_coeffs) return status_to_code(HBMPC_INVALID_INPUT);
    const uintptr_t S = shares.len;
    if (S == 0 || !shares.pointer) return status_to_code(HBMPC_INVALID_INPUT);
    {
        std::set<uintptr_t> seen;
        for (uintptr_t i = 0; i < S; ++i)
            if (!seen.insert(shares.pointer[i].id).second) return status_to_code(HBMPC_INVALID_INPUT);
    }
    const uintptr_t deg = shares.pointer[0].degree;
    for (uintptr_t i = 0; i < S; ++i)
        if (shares.pointer[i].degree != deg) return status_to_code(HBMPC_DEGREE_MISMATCH);
    std::vector<size_t> ids(S);
    std::vector<uint64_t> y(4 * S), c(4 * (deg + 1)), sec(4);
    for (uintptr_t i = 0; i < S; ++i) {
        ids[i] = shares.pointer[i].id;
        std::memcpy(&y[4 * i], unbox(shares.pointer[i].share), 32);
    }
    std::lock_guard<std::mutex> lk(g_mu);
    hbmpc_ctx *ctx = context();
    if (!ctx) return status_to_code(g_last_status);
    int32_t st = 0;
    int rc = hbmpc_nonrobust_recover_batch(ctx, n, deg, S, ids.data(), 1, y.data(), 0, c.data(), sec.data(), &st);
    if (rc) return status_to_code(rc);
    if (st < 0) return status_to_code(-st);
    std::memcpy(output_secret->data, sec.data(), 32);
    emit_coeffs(c, deg + 1, output_coeffs);
    return status_to_code(HBMPC_SUCCESS);
}

}  // extern "C"
