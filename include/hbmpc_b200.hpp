// hbmpc_b200.hpp -- C++17 host-side mirror of the reference's share API over the C ABI (hbmpc_b200.h).
//
// The reference host is Rust (not available in this image), so the operator interface of the hot path is mirrored in
// C++ with the reference's names, argument meaning and error behaviour, so that the tests in tests/host/ read like the
// reference's own unit tests:
//   SecretSharingScheme::{compute_shares, recover_secret}      common/mod.rs:101-128
//   RobustShare                                                honeybadger/robust_interpolate/robust_interpolate.rs:16-158
//   NonRobustShare                                             common/share/shamir.rs:127-240
//   batch_recover_secret                                       robust_interpolate.rs:284-443
//   make_vandermonde / apply_vandermonde                       common/share/mod.rs:31-76
// Every function is a thin wrapper: all field arithmetic runs in the CUDA kernels behind the C ABI (no CPU fallback).
// Errors are the reference's ShareErrorCode numbers (ffi/c_bindings/share/mod.rs:18-37) carried by `ShareError`.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "hbmpc_b200.h"

namespace hbmpc {

using U256 = std::array<uint64_t, 4>;  // canonical value, little-endian limbs (ffi/c_bindings/mod.rs:17-49)

inline constexpr U256 FR_MODULUS = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};

struct ShareError : std::runtime_error {
    int code;
    explicit ShareError(int c) : std::runtime_error("ShareErrorCode " + std::to_string(c)), code(c) {}
};
inline void check(int rc) {
    if (rc != HBMPC_SUCCESS) throw ShareError(rc);
}

inline U256 fr_from_u64(uint64_t v) { return U256{v, 0, 0, 0}; }
inline bool fr_is_canonical(const U256 &x) {
    for (int i = 3; i >= 0; --i) {
        if (x[i] < FR_MODULUS[i]) return true;
        if (x[i] > FR_MODULUS[i]) return false;
    }
    return false;
}
// F::rand analogue: uniform canonical element by rejection from a 64-bit generator.  (arkworks maps the accepted limbs
// through the Montgomery representation; the distribution is the same, the stream-to-value mapping is not reproduced.)
template <class Rng>
inline U256 fr_rand(Rng &next_u64) {
    for (;;) {
        U256 x = {next_u64(), next_u64(), next_u64(), next_u64() >> 1};
        if (fr_is_canonical(x)) return x;
    }
}

class Context {
   public:
    explicit Context(int device = 0) { check(hbmpc_ctx_create(device, &ctx_)); }
    ~Context() { hbmpc_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    hbmpc_ctx *get() const { return ctx_; }

   private:
    hbmpc_ctx *ctx_ = nullptr;
};

// ShamirShare<F, 1, P> (common/mod.rs:92-99): one field value, the evaluation-point index and the polynomial degree
struct Share {
    U256 share{};
    size_t id = 0;
    size_t degree = 0;
};

namespace detail {
inline void same_degree(const std::vector<Share> &shares) {
    for (const Share &s : shares)
        if (s.degree != shares[0].degree) throw ShareError(HBMPC_DEGREE_MISMATCH);
}
inline std::vector<Share> compute_shares(Context &ctx, const U256 &secret, size_t n, size_t degree, const std::function<uint64_t()> &rng) {
    // DensePolynomial::rand(degree) then coeffs[0] = secret (robust_interpolate.rs:68-69, shamir.rs:181-182)
    std::vector<U256> coeffs(degree + 1);
    for (auto &c : coeffs) c = fr_rand(rng);
    coeffs[0] = secret;
    std::vector<U256> out(n);
    check(hbmpc_compute_shares_batch(ctx.get(), n, degree, 1, coeffs[0].data(), n ? out[0].data() : nullptr));
    std::vector<Share> shares(n);
    for (size_t i = 0; i < n; ++i) shares[i] = Share{out[i], i, degree};
    return shares;
}
inline void trim(std::vector<U256> &c) {  // DensePolynomial keeps no trailing zero coefficients
    while (!c.empty() && c.back() == U256{0, 0, 0, 0}) c.pop_back();
}
}  // namespace detail

// RobustShare<F> : SecretSharingScheme<F>
struct RobustShare {
    static std::vector<Share> compute_shares(Context &ctx, const U256 &secret, size_t n, size_t degree, const std::function<uint64_t()> &rng) {
        return detail::compute_shares(ctx, secret, n, degree, rng);
    }
    // -> (poly.coeffs (trimmed), poly(0));  throws ShareError(InvalidInput / DegreeMismatch / DecodingError)
    static std::pair<std::vector<U256>, U256> recover_secret(Context &ctx, const std::vector<Share> &shares, size_t n, size_t t) {
        if (n < 3 * t + 1) throw ShareError(HBMPC_INVALID_INPUT);
        if (shares.empty()) throw ShareError(HBMPC_INVALID_INPUT);
        detail::same_degree(shares);
        const size_t d = shares[0].degree, S = shares.size();
        std::vector<size_t> ids(S);
        std::vector<U256> vals(S), coeffs(d + 1);
        for (size_t i = 0; i < S; ++i) { ids[i] = shares[i].id; vals[i] = shares[i].share; }
        U256 secret{};
        int32_t path = 0;
        check(hbmpc_robust_interpolate_batch(ctx.get(), n, d, t, S, ids.data(), 1, vals[0].data(), coeffs[0].data(), secret.data(), &path, nullptr));
        detail::trim(coeffs);
        return {coeffs, secret};
    }
};

// NonRobustShare<F> : SecretSharingScheme<F>
struct NonRobustShare {
    static std::vector<Share> compute_shares(Context &ctx, const U256 &secret, size_t n, size_t degree, const std::function<uint64_t()> &rng) {
        if (n <= degree) throw ShareError(HBMPC_INVALID_INPUT);
        return detail::compute_shares(ctx, secret, n, degree, rng);
    }
    static std::pair<std::vector<U256>, U256> recover_secret(Context &ctx, const std::vector<Share> &shares, size_t n) {
        if (shares.empty()) throw ShareError(HBMPC_INVALID_INPUT);
        const size_t S = shares.size();
        std::vector<size_t> ids(S);
        std::vector<U256> vals(S);
        for (size_t i = 0; i < S; ++i) { ids[i] = shares[i].id; vals[i] = shares[i].share; }
        for (size_t i = 0; i < S; ++i)
            for (size_t j = i + 1; j < S; ++j)
                if (ids[i] == ids[j]) throw ShareError(HBMPC_INVALID_INPUT);
        detail::same_degree(shares);
        const size_t d = shares[0].degree;
        std::vector<U256> coeffs(d + 1);
        U256 secret{};
        int32_t status = 0;
        check(hbmpc_nonrobust_recover_batch(ctx.get(), n, d, S, ids.data(), 1, vals[0].data(), 0, coeffs[0].data(), secret.data(), &status));
        if (status < 0) throw ShareError(-status);
        detail::trim(coeffs);
        return {coeffs, secret};
    }
};

// batch_recover_secret(evals_by_sender, n, degree, t): one coefficient vector (length degree+1) per chunk
inline std::vector<std::vector<U256>> batch_recover_secret(Context &ctx, const std::vector<std::pair<size_t, std::vector<U256>>> &evals_by_sender,
                                                           size_t n, size_t degree, size_t t) {
    if (n < 3 * t + 1) throw ShareError(HBMPC_INVALID_INPUT);
    if (evals_by_sender.empty()) throw ShareError(HBMPC_INVALID_INPUT);
    const size_t S = evals_by_sender.size(), B = evals_by_sender[0].second.size();
    if (B == 0) throw ShareError(HBMPC_INVALID_INPUT);
    std::vector<size_t> ids(S);
    std::vector<U256> evals(S * B);
    for (size_t i = 0; i < S; ++i) {
        if (evals_by_sender[i].second.size() != B) throw ShareError(HBMPC_INVALID_INPUT);  // "Inconsistent batch widths"
        ids[i] = evals_by_sender[i].first;
        for (size_t c = 0; c < B; ++c) evals[i * B + c] = evals_by_sender[i].second[c];
    }
    std::vector<U256> coeffs(B * (degree + 1));
    std::vector<int32_t> path(B);
    check(hbmpc_batch_recover(ctx.get(), n, degree, t, S, ids.data(), B, evals[0].data(), coeffs[0].data(), path.data(), nullptr));
    std::vector<std::vector<U256>> out(B);
    for (size_t c = 0; c < B; ++c) out[c].assign(coeffs.begin() + c * (degree + 1), coeffs.begin() + (c + 1) * (degree + 1));
    return out;
}

// apply_vandermonde(make_vandermonde(n, t), shares): out_j = sum_k V[j][k] * shares[k]; ids/degrees of the inputs must agree
inline std::vector<Share> apply_vandermonde(Context &ctx, size_t n, const std::vector<Share> &shares) {
    if (shares.empty()) throw ShareError(HBMPC_INVALID_INPUT);
    for (const Share &s : shares) {
        if (s.degree != shares[0].degree) throw ShareError(HBMPC_DEGREE_MISMATCH);  // Add: common/mod.rs:170-177
        if (s.id != shares[0].id) throw ShareError(HBMPC_ID_MISMATCH);
    }
    std::vector<U256> in(shares.size()), out(n);
    for (size_t k = 0; k < shares.size(); ++k) in[k] = shares[k].share;
    check(hbmpc_apply_vandermonde_batch(ctx.get(), n, shares.size(), 1, in[0].data(), out[0].data(), 0));
    std::vector<Share> res(n);
    for (size_t j = 0; j < n; ++j) res[j] = Share{out[j], shares[0].id, shares[0].degree};
    return res;
}

}  // namespace hbmpc
