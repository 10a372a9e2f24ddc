// tables.hpp -- host-side construction of the constant matrices the kernels consume (Montgomery form).
// Set-up code only: O(n^2) field operations per (n, d, t, id-set), cached per context.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "host_fr.hpp"

namespace hb {

// GeneralEvaluationDomain::new(n) of the reference (common/mod.rs:51-68): radix-2, size N = next_pow2(n),
// element(j) = w_N^j with w_N = root32^(2^32/N).
inline int domain_size(size_t n) {
    if (n == 0 || n > 256) return 0;
    int N = 1;
    while ((size_t)N < n) N <<= 1;
    return N;
}
inline std::vector<HFr> domain_elements(size_t n, size_t count) {
    int N = domain_size(n);
    HFr w = hfr::from_canon(hfr::ROOT32_CANON);
    for (uint64_t k = (1ULL << 32) / (uint64_t)N; k > 1; k >>= 1) w = hfr::mul(w, w);
    std::vector<HFr> x(count);
    HFr p = hfr::ONE;
    for (size_t j = 0; j < count; ++j) {
        x[j] = p;
        p = hfr::mul(p, w);
    }
    return x;
}

// make_vandermonde (common/share/mod.rs:31-45): V[j][k] = element(j)^k, rows x cols, row-major
inline std::vector<HFr> vandermonde_on_points(const std::vector<HFr> &pts, size_t cols) {
    std::vector<HFr> V(pts.size() * cols);
    for (size_t j = 0; j < pts.size(); ++j) {
        HFr p = hfr::ONE;
        for (size_t k = 0; k < cols; ++k) {
            V[j * cols + k] = p;
            p = hfr::mul(p, pts[j]);
        }
    }
    return V;
}

// Lagrange data for interpolation points xs[0..m):
//   Lc[k][i] = coefficient k of L_i(x)  (m x m, row-major)      -- basis_coeffs of robust_interpolate.rs:353-391
//   w[i]     = 1 / prod_{j != i} (xs[i] - xs[j])
//   A[0..m]  = coefficients of prod_i (x - xs[i])
struct Lagrange {
    size_t m;
    std::vector<HFr> Lc, w, A;
};
inline Lagrange lagrange_basis(const std::vector<HFr> &xs) {
    Lagrange L;
    size_t m = xs.size();
    L.m = m;
    L.A.assign(m + 1, hfr::ZERO);
    L.A[0] = hfr::ONE;
    size_t deg = 0;
    for (size_t i = 0; i < m; ++i) {  // A *= (x - xs[i])
        HFr nx = hfr::neg(xs[i]);
        L.A[deg + 1] = L.A[deg];
        for (size_t k = deg; k >= 1; --k) L.A[k] = hfr::add(L.A[k - 1], hfr::mul(L.A[k], nx));
        L.A[0] = hfr::mul(L.A[0], nx);
        ++deg;
    }
    L.w.resize(m);
    for (size_t i = 0; i < m; ++i) {
        HFr p = hfr::ONE;
        for (size_t j = 0; j < m; ++j)
            if (j != i) p = hfr::mul(p, hfr::sub(xs[i], xs[j]));
        L.w[i] = p;
    }
    hfr::batch_inv(L.w);
    L.Lc.assign(m * m, hfr::ZERO);
    std::vector<HFr> q(m);
    for (size_t i = 0; i < m; ++i) {  // q = A / (x - xs[i]) by synthetic division
        q[m - 1] = L.A[m];
        for (size_t k = m - 1; k >= 1; --k) q[k - 1] = hfr::add(L.A[k], hfr::mul(xs[i], q[k]));
        for (size_t k = 0; k < m; ++k) L.Lc[k * m + i] = hfr::mul(q[k], L.w[i]);
    }
    return L;
}
// L_i(x) for all i at each extra point:  rows x m, row-major   (verify_matrix rows s >= m, :392-399)
inline std::vector<HFr> lagrange_eval_rows(const std::vector<HFr> &xs, const Lagrange &L, const std::vector<HFr> &pts) {
    size_t m = xs.size();
    std::vector<HFr> out(pts.size() * m), diffs(pts.size() * m);
    for (size_t s = 0; s < pts.size(); ++s)
        for (size_t i = 0; i < m; ++i) diffs[s * m + i] = hfr::sub(pts[s], xs[i]);
    std::vector<HFr> inv = diffs;
    hfr::batch_inv(inv);  // extra points are distinct from the interpolation points (ids are unique)
    for (size_t s = 0; s < pts.size(); ++s) {
        HFr Ax = hfr::ONE;
        for (size_t i = 0; i < m; ++i) Ax = hfr::mul(Ax, diffs[s * m + i]);
        for (size_t i = 0; i < m; ++i) out[s * m + i] = hfr::mul(hfr::mul(Ax, inv[s * m + i]), L.w[i]);
    }
    return out;
}

inline void to_u32(const std::vector<HFr> &v, std::vector<uint32_t> &out) {
    out.resize(v.size() * 8);
    for (size_t i = 0; i < v.size(); ++i)
        for (int k = 0; k < 4; ++k) {
            out[i * 8 + 2 * k] = (uint32_t)v[i].l[k];
            out[i * 8 + 2 * k + 1] = (uint32_t)(v[i].l[k] >> 32);
        }
}

}  // namespace hb
