// host_fr.hpp -- host-side Fr arithmetic used ONLY to build the small per-(n, d, t, id-set) constant tables
// (evaluation points, Vandermonde rows, Lagrange recover/verify matrices, syndrome matrices) that the CUDA
// kernels consume.  It is set-up code (O(n^2) field operations per table, cached in the context), not a data path:
// no share, coefficient or codeword of a batch is ever processed on the CPU.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace hb {

struct HFr {
    uint64_t l[4];  // Montgomery form, R = 2^256
};

namespace hfr {

typedef unsigned __int128 u128;
static const uint64_t MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const uint64_t NINV = 0xfffffffeffffffffULL;  // -r^{-1} mod 2^64
static const HFr ONE = {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}};
static const HFr R2 = {{0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}};
static const HFr ZERO = {{0, 0, 0, 0}};
// canonical 2^32-th root of unity of Fr: 7^((r-1)/2^32)
static const uint64_t ROOT32_CANON[4] = {0x3829971f439f0d2bULL, 0xb63683508c2280b9ULL, 0xd09b681922c813b4ULL, 0x16a2a19edfe81f20ULL};

inline bool geq_mod(const uint64_t *a) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > MOD[i]) return true;
        if (a[i] < MOD[i]) return false;
    }
    return true;
}
inline void sub_mod(uint64_t *a) {
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - MOD[i] - br;
        a[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}
inline bool is_zero(const HFr &a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
inline bool eq(const HFr &a, const HFr &b) { return std::memcmp(a.l, b.l, 32) == 0; }
inline HFr add(const HFr &a, const HFr &b) {
    HFr c;
    u128 cy = 0;
    for (int i = 0; i < 4; ++i) {
        u128 s = (u128)a.l[i] + b.l[i] + cy;
        c.l[i] = (uint64_t)s;
        cy = s >> 64;
    }
    if (cy || geq_mod(c.l)) sub_mod(c.l);
    return c;
}
inline HFr sub(const HFr &a, const HFr &b) {
    HFr c;
    u128 br = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a.l[i] - b.l[i] - br;
        c.l[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        u128 cy = 0;
        for (int i = 0; i < 4; ++i) {
            u128 s = (u128)c.l[i] + MOD[i] + cy;
            c.l[i] = (uint64_t)s;
            cy = s >> 64;
        }
    }
    return c;
}
inline HFr neg(const HFr &a) { return is_zero(a) ? a : sub(ZERO, a); }
// separated operand scanning: full 512-bit product, then 4 Montgomery rounds
inline HFr mul(const HFr &a, const HFr &b) {
    uint64_t t[9] = {0};
    for (int i = 0; i < 4; ++i) {
        u128 cy = 0;
        for (int j = 0; j < 4; ++j) {
            u128 s = (u128)a.l[i] * b.l[j] + t[i + j] + cy;
            t[i + j] = (uint64_t)s;
            cy = s >> 64;
        }
        t[i + 4] = (uint64_t)cy;
    }
    for (int i = 0; i < 4; ++i) {
        uint64_t m = t[i] * NINV;
        u128 cy = 0;
        for (int j = 0; j < 4; ++j) {
            u128 s = (u128)m * MOD[j] + t[i + j] + cy;
            t[i + j] = (uint64_t)s;
            cy = s >> 64;
        }
        for (int k = i + 4; k < 9 && cy; ++k) {
            u128 s = (u128)t[k] + cy;
            t[k] = (uint64_t)s;
            cy = s >> 64;
        }
    }
    HFr c = {{t[4], t[5], t[6], t[7]}};
    if (t[8] || geq_mod(c.l)) sub_mod(c.l);
    return c;
}
inline HFr inv(const HFr &a) {  // a^(r-2)
    uint64_t e[4] = {MOD[0] - 2, MOD[1], MOD[2], MOD[3]};
    HFr acc = ONE;
    for (int i = 255; i >= 0; --i) {
        acc = mul(acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a);
    }
    return acc;
}
inline HFr from_canon(const uint64_t *u) {
    HFr a = {{u[0], u[1], u[2], u[3]}};
    return mul(a, R2);
}
inline void to_canon(const HFr &a, uint64_t *u) {
    HFr one = {{1, 0, 0, 0}};
    HFr c = mul(a, one);
    std::memcpy(u, c.l, 32);
}
inline HFr from_u64(uint64_t v) {
    uint64_t u[4] = {v, 0, 0, 0};
    return from_canon(u);
}
// Montgomery's trick: invert all (non-zero) entries with one field inversion
inline void batch_inv(std::vector<HFr> &v) {
    if (v.empty()) return;
    std::vector<HFr> pre(v.size());
    HFr acc = ONE;
    for (size_t i = 0; i < v.size(); ++i) {
        pre[i] = acc;
        acc = mul(acc, v[i]);
    }
    HFr ia = inv(acc);
    for (size_t i = v.size(); i-- > 0;) {
        HFr t = mul(ia, pre[i]);
        ia = mul(ia, v[i]);
        v[i] = t;
    }
}

}  // namespace hfr
}  // namespace hb
