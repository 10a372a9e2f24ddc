// mirror_test.cpp -- the reference's own unit tests for the hot path, re-stated against the C++ mirror
// (include/hbmpc_b200.hpp) and the plain C ABI.  Built with g++ and run on the GPU box by tests/test_gpu_host_mirror.py.
//   robust_interpolate.rs:646-680  test_robust_interpolate_fnt        (f = 7 + 3x + 5x^2, n=16, t=2, first 2t+1 shares)
//   robust_interpolate.rs:791-826  full robust interpolation with 2 errors (n=10, t=3)
//   robust_interpolate.rs:880-927  batch_recover_secret == per-chunk recover_secret, reversed arrival order
//   shamir.rs:250-451              NonRobustShare round trip and error variants
//   ffi/tests/secret_share.c:64-118  {3,3,22,22} U256 secret, n=6, degree 2, t=1 through the C ABI
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "hbmpc_b200.hpp"

using namespace hbmpc;

static uint64_t sm = 0x1234;
static uint64_t next_u64() {
    uint64_t z = (sm += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
#define REQUIRE(c)                                                       \
    do {                                                                 \
        if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)
template <class F>
static int throws_code(F f) {
    try { f(); } catch (const ShareError &e) { return e.code; }
    return 0;
}

int main() {
    Context ctx(0);
    {   // test_robust_interpolate_fnt: fixed polynomial through the C ABI, then optimistic recovery from 2t+1 shares
        const size_t n = 16, t = 2;
        U256 coeffs[3] = {fr_from_u64(7), fr_from_u64(3), fr_from_u64(5)};
        U256 out[16];
        REQUIRE(hbmpc_compute_shares_batch(ctx.get(), n, 2, 1, coeffs[0].data(), out[0].data()) == HBMPC_SUCCESS);
        REQUIRE(out[0] == fr_from_u64(15));
        std::vector<Share> shares;
        for (size_t i = 0; i < 2 * t + 1; ++i) shares.push_back(Share{out[i], i, 2});
        auto [poly, secret] = RobustShare::recover_secret(ctx, shares, n, t);
        REQUIRE(poly.size() == 3 && poly[0] == coeffs[0] && poly[1] == coeffs[1] && poly[2] == coeffs[2] && secret == fr_from_u64(7));
    }
    {   // full robust interpolation, n=10, t=3, secret 42, two corrupted shares
        const size_t n = 10, t = 3;
        auto shares = RobustShare::compute_shares(ctx, fr_from_u64(42), n, t, next_u64);
        REQUIRE(shares.size() == n && shares[3].id == 3 && shares[3].degree == t);
        shares[1].share[0] ^= 5;
        shares[6].share[0] ^= 9;
        std::reverse(shares.begin(), shares.end());
        auto [poly, secret] = RobustShare::recover_secret(ctx, shares, n, t);
        REQUIRE(secret == fr_from_u64(42) && poly.size() == t + 1 && poly[0] == secret);
        // > t errors with exactly d+t+1 shares cannot be decoded
        std::vector<Share> few(shares.begin(), shares.begin() + 2 * t + 1);
        few[0].share[1] ^= 1;
        REQUIRE(throws_code([&] { RobustShare::recover_secret(ctx, few, n, t); }) == HBMPC_DECODING_ERROR);
        REQUIRE(throws_code([&] { RobustShare::recover_secret(ctx, shares, 9, t); }) == HBMPC_INVALID_INPUT);  // n < 3t+1
        auto dup = shares; dup[1].id = dup[0].id;
        REQUIRE(throws_code([&] { RobustShare::recover_secret(ctx, dup, n, t); }) == HBMPC_INVALID_INPUT);
        auto mixed = shares; mixed[2].degree = t + 1;
        REQUIRE(throws_code([&] { RobustShare::recover_secret(ctx, mixed, n, t); }) == HBMPC_DEGREE_MISMATCH);
    }
    {   // batch_recover_secret == per-chunk recover_secret for reversed arrival order
        const size_t n = 10, t = 3, B = 8;
        std::vector<std::vector<Share>> per_chunk(B);
        for (size_t c = 0; c < B; ++c) per_chunk[c] = RobustShare::compute_shares(ctx, fr_rand(next_u64), n, t, next_u64);
        std::vector<std::pair<size_t, std::vector<U256>>> by_sender;
        for (size_t i = n; i-- > 0;) {
            std::vector<U256> v(B);
            for (size_t c = 0; c < B; ++c) v[c] = per_chunk[c][i].share;
            by_sender.push_back({i, v});
        }
        auto batch = batch_recover_secret(ctx, by_sender, n, t, t);
        REQUIRE(batch.size() == B);
        for (size_t c = 0; c < B; ++c) {
            auto [poly, secret] = RobustShare::recover_secret(ctx, per_chunk[c], n, t);
            poly.resize(t + 1, U256{0, 0, 0, 0});
            REQUIRE(batch[c] == poly && batch[c][0] == secret);
        }
    }
    {   // NonRobustShare: round trip, degree mismatch, insufficient shares, duplicate ids; Vandermonde apply
        const size_t n = 7, d = 2;
        auto shares = NonRobustShare::compute_shares(ctx, fr_from_u64(99), n, d, next_u64);
        auto [poly, secret] = NonRobustShare::recover_secret(ctx, shares, n);
        REQUIRE(secret == fr_from_u64(99) && poly.size() == d + 1);
        auto wrong = shares;
        for (auto &s : wrong) s.degree = 1;  // claimed degree lower than the real one -> DegreeMismatch (shamir.rs:234-237)
        REQUIRE(throws_code([&] { NonRobustShare::recover_secret(ctx, wrong, n); }) == HBMPC_DEGREE_MISMATCH);
        std::vector<Share> two(shares.begin(), shares.begin() + 2);
        REQUIRE(throws_code([&] { NonRobustShare::recover_secret(ctx, two, n); }) == HBMPC_INSUFFICIENT_SHARES);
        auto dup = shares; dup[1].id = 0;
        REQUIRE(throws_code([&] { NonRobustShare::recover_secret(ctx, dup, n); }) == HBMPC_INVALID_INPUT);
        REQUIRE(throws_code([&] { NonRobustShare::compute_shares(ctx, secret, 2, 2, next_u64); }) == HBMPC_INVALID_INPUT);
        std::vector<Share> col = {Share{fr_from_u64(1), 4, 1}, Share{fr_from_u64(2), 4, 1}, Share{fr_from_u64(3), 4, 1}};
        auto y = apply_vandermonde(ctx, 4, col);  // y_0 = 1 + 2 + 3
        REQUIRE(y.size() == 4 && y[0].share == fr_from_u64(6) && y[2].share == fr_from_u64(2) && y[1].id == 4 && y[1].degree == 1);
        col[1].id = 5;
        REQUIRE(throws_code([&] { apply_vandermonde(ctx, 4, col); }) == HBMPC_ID_MISMATCH);
    }
    {   // ffi/tests/secret_share.c: {3,3,22,22}, n = 6, degree 2, t = 1, plain C ABI, memcmp on the limbs
        const size_t n = 6, d = 2, t = 1;
        uint64_t coeffs[3][4] = {{3, 3, 22, 22}, {11, 0, 0, 1}, {5, 6, 7, 8}};
        uint64_t shares[6][4], rec[3][4], secret[4];
        size_t ids[6] = {0, 1, 2, 3, 4, 5};
        int32_t path = -1;
        REQUIRE(hbmpc_compute_shares_batch(ctx.get(), n, d, 1, &coeffs[0][0], &shares[0][0]) == HBMPC_SUCCESS);
        REQUIRE(hbmpc_robust_interpolate_batch(ctx.get(), n, d, t, n, ids, 1, &shares[0][0], &rec[0][0], secret, &path, nullptr) == HBMPC_SUCCESS);
        REQUIRE(path == 0 && memcmp(secret, coeffs[0], 32) == 0 && memcmp(rec, coeffs, sizeof coeffs) == 0);
    }
    printf("mirror_test: all reference-shaped tests passed\n");
    return 0;
}
