// fr.cuh -- BLS12-381 scalar field (ark_bls12_381::Fr) arithmetic for sm_100a integer pipes.
//
// Representation: 8 x 32-bit little-endian limbs.  Data crossing the C ABI is the CANONICAL value
// (reference: U256, /root/reference/mpc/src/ffi/c_bindings/mod.rs:17-49).  Constant operands
// (evaluation points, Vandermonde / Lagrange matrices) are stored in Montgomery form c*R mod r,
// R = 2^256, so that  REDC(sum_k data_k * constR_k) = sum_k data_k * const_k mod r  is canonical again:
// the hot kernels run directly on the boundary format with no conversion pass.
//
// Multiply-accumulate uses the even/odd column split: each 32x32->64 product is one IMAD.WIDE.U32(.X)
// whose carry travels in a predicate along a chain of four 64-bit lanes; the carry leaving a chain is
// counted in a small per-position counter (IADD3.X on the ALU pipe) instead of rippling.  A sum of up
// to 256 products is accumulated lazily (no reduction per term) and Montgomery-reduced once.
#pragma once
#include <cstdint>

// HB_HOST_EMULATION (tests only): lets g++ compile this header and run the limb logic on the CPU so the index
// bookkeeping of the even/odd accumulator can be unit-tested without a GPU.  The shipped library never defines it.
#if defined(__CUDACC__)
#define HB_DEV __device__ __forceinline__
#else
#define HB_DEV static inline
#ifndef HB_HOST_EMULATION
#error "fr.cuh is device code; define HB_HOST_EMULATION only in the host unit test"
#endif
struct uint4 { uint32_t x, y, z, w; };
#endif

namespace hb {

// r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
#define HB_R0 0x00000001u
#define HB_R1 0xffffffffu
#define HB_R2 0xfffe5bfeu
#define HB_R3 0x53bda402u
#define HB_R4 0x09a1d805u
#define HB_R5 0x3339d808u
#define HB_R6 0x299d7d48u
#define HB_R7 0x73eda753u
// modulus limbs as PTX immediates (PTX does not accept the C "u" suffix)
#define HB_PR0 "0x00000001"
#define HB_PR1 "0xffffffff"
#define HB_PR2 "0xfffe5bfe"
#define HB_PR3 "0x53bda402"
#define HB_PR4 "0x09a1d805"
#define HB_PR5 "0x3339d808"
#define HB_PR6 "0x299d7d48"
#define HB_PR7 "0x73eda753"
// -r^{-1} mod 2^32 == 0xffffffff  =>  Montgomery quotient digit m = -T[i] mod 2^32 (no multiply)

struct fr_t {
    uint32_t v[8];
};

// ----------------------------------------------------------------------------------------------
// carry-chain primitives
// ----------------------------------------------------------------------------------------------
// (l0,l1,l2,l3) += (a0 + a1*2^64 + a2*2^128 + a3*2^192) * b   (four 64-bit lanes), k += carry-out.
// The lanes are 64-bit variables so that ptxas keeps each (lo,hi) in the aligned register pair IMAD.WIDE needs.
HB_DEV void chain4w(unsigned long long &l0, unsigned long long &l1, unsigned long long &l2, unsigned long long &l3, uint32_t &k,
                    uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t"
        ".reg .u32 x0, x1, x2, x3, x4, x5, x6, x7;\n\t"
        "mov.b64 {x0, x1}, %0;\n\t"
        "mov.b64 {x2, x3}, %1;\n\t"
        "mov.b64 {x4, x5}, %2;\n\t"
        "mov.b64 {x6, x7}, %3;\n\t"
        "mad.lo.cc.u32 x0, %5, %9, x0;\n\t"
        "madc.hi.cc.u32 x1, %5, %9, x1;\n\t"
        "madc.lo.cc.u32 x2, %6, %9, x2;\n\t"
        "madc.hi.cc.u32 x3, %6, %9, x3;\n\t"
        "madc.lo.cc.u32 x4, %7, %9, x4;\n\t"
        "madc.hi.cc.u32 x5, %7, %9, x5;\n\t"
        "madc.lo.cc.u32 x6, %8, %9, x6;\n\t"
        "madc.hi.cc.u32 x7, %8, %9, x7;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mov.b64 %0, {x0, x1};\n\t"
        "mov.b64 %1, {x2, x3};\n\t"
        "mov.b64 %2, {x4, x5};\n\t"
        "mov.b64 %3, {x6, x7};\n\t"
        "}"
        : "+l"(l0), "+l"(l1), "+l"(l2), "+l"(l3), "+r"(k)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
    unsigned long long *x[4] = {&l0, &l1, &l2, &l3};
    const uint32_t a[4] = {a0, a1, a2, a3};
    unsigned cy = 0;
    for (int l = 0; l < 4; ++l) {
        unsigned __int128 s = (unsigned __int128)a[l] * b + *x[l] + cy;
        *x[l] = (unsigned long long)s;
        cy = (unsigned)(s >> 64);
    }
    k += cy;
#endif
}
// 32-bit view kept for the Montgomery rows of acc_reduce
HB_DEV void chain4(uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3, uint32_t &x4, uint32_t &x5,
                                       uint32_t &x6, uint32_t &x7, uint32_t &k, uint32_t a0, uint32_t a1, uint32_t a2,
                                       uint32_t a3, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7), "+r"(k)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
    uint32_t *x[8] = {&x0, &x1, &x2, &x3, &x4, &x5, &x6, &x7};
    const uint32_t a[4] = {a0, a1, a2, a3};
    unsigned cy = 0;
    for (int l = 0; l < 4; ++l) {
        unsigned __int128 s = (unsigned __int128)a[l] * b + (((uint64_t)*x[2 * l + 1] << 32) | *x[2 * l]) + cy;
        *x[2 * l] = (uint32_t)s;
        *x[2 * l + 1] = (uint32_t)(s >> 32);
        cy = (unsigned)(s >> 64);
    }
    k += cy;
#endif
}

// Lazy accumulator for sum_k a_k * b_k (each factor < 2^256).  E[j] holds the even-aligned 64-bit lane at limb positions
// (2j, 2j+1), O[j] the odd-aligned lane at (2j+1, 2j+2), K[p] counts carries into limb position p.
struct acc_t {
    unsigned long long E[8];
    unsigned long long O[7];
    uint32_t K[17];  // positions 8..16 used
};

HB_DEV void acc_zero(acc_t &A) {
#pragma unroll
    for (int i = 0; i < 8; ++i) A.E[i] = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) A.O[i] = 0;
#pragma unroll
    for (int i = 0; i < 17; ++i) A.K[i] = 0;
}

// A += a * b  (64 IMAD.WIDE + 16 carry-counter adds)
HB_DEV void acc_mac(acc_t &A, const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        const int i = 2 * h;
        chain4w(A.E[h], A.E[h + 1], A.E[h + 2], A.E[h + 3], A.K[i + 8], a[0], a[2], a[4], a[6], b[i]);
        chain4w(A.O[h], A.O[h + 1], A.O[h + 2], A.O[h + 3], A.K[i + 9], a[1], a[3], a[5], a[7], b[i]);
        chain4w(A.O[h], A.O[h + 1], A.O[h + 2], A.O[h + 3], A.K[i + 9], a[0], a[2], a[4], a[6], b[i + 1]);
        chain4w(A.E[h + 1], A.E[h + 2], A.E[h + 3], A.E[h + 4], A.K[i + 10], a[1], a[3], a[5], a[7], b[i + 1]);
    }
}

// one Montgomery row: T[i..i+8] += m * r, carries leaving the two chains counted in k8 / k9
HB_DEV void redc_row(uint32_t &t0, uint32_t &t1, uint32_t &t2, uint32_t &t3, uint32_t &t4, uint32_t &t5,
                                         uint32_t &t6, uint32_t &t7, uint32_t &t8, uint32_t &k8, uint32_t &k9) {
    uint32_t m = 0u - t0;
    chain4(t0, t1, t2, t3, t4, t5, t6, t7, k8, HB_R0, HB_R2, HB_R4, HB_R6, m);
    chain4(t1, t2, t3, t4, t5, t6, t7, t8, k9, HB_R1, HB_R3, HB_R5, HB_R7, m);
}

// ----------------------------------------------------------------------------------------------
// plain 256-bit helpers
// ----------------------------------------------------------------------------------------------
// d = a - b, returns borrow (1 if a < b)
HB_DEV uint32_t sub8(uint32_t (&d)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#if defined(__CUDA_ARCH__)
    uint32_t br;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(d[0]), "=&r"(d[1]), "=&r"(d[2]), "=&r"(d[3]), "=&r"(d[4]), "=&r"(d[5]), "=&r"(d[6]), "=&r"(d[7]), "=&r"(br)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]),
          "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return br & 1u;
#else
    uint32_t br = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t t = (uint64_t)a[i] - b[i] - br;
        d[i] = (uint32_t)t;
        br = (uint32_t)(t >> 63);
    }
    return br;
#endif
}
// d = a + b + cin (cin in {0,1}), returns carry
HB_DEV uint32_t add8c(uint32_t (&d)[8], const uint32_t (&a)[8], const uint32_t (&b)[8], uint32_t cin) {
#if defined(__CUDA_ARCH__)
    uint32_t cy;
    asm("add.cc.u32 %8, %25, 0xffffffff;\n\t"
        "addc.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(d[0]), "=&r"(d[1]), "=&r"(d[2]), "=&r"(d[3]), "=&r"(d[4]), "=&r"(d[5]), "=&r"(d[6]), "=&r"(d[7]), "=&r"(cy)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]),
          "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(cin));
    return cy;
#else
    uint32_t cy = cin;
    for (int i = 0; i < 8; ++i) {
        uint64_t t = (uint64_t)a[i] + b[i] + cy;
        d[i] = (uint32_t)t;
        cy = (uint32_t)(t >> 32);
    }
    return cy;
#endif
}
HB_DEV uint32_t add8(uint32_t (&d)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { return add8c(d, a, b, 0); }

HB_DEV void mod_limbs(uint32_t (&r)[8]) {
    r[0] = HB_R0; r[1] = HB_R1; r[2] = HB_R2; r[3] = HB_R3; r[4] = HB_R4; r[5] = HB_R5; r[6] = HB_R6; r[7] = HB_R7;
}

// x >= r ?   The top limb decides in all but one case in 2^31 for values that are uniform below r (x_7 < r_7: no; x_7 > r_7: yes);
// only x_7 == r_7 needs the full-width comparison (one compare instead of an 8-limb borrow chain per validated input).
HB_DEV bool geq_mod(const uint32_t (&x)[8]) {
#if defined(__CUDA_ARCH__)
    if (x[7] != HB_R7) return x[7] > HB_R7;
#endif
    uint32_t r[8], d[8];
    mod_limbs(r);
    return sub8(d, x, r) == 0;
}
// x in [0, 2r) -> x mod r, exact.  The decision is taken on the top limb alone: x_7 > r_7 means x >= (r_7 + 1)*2^224 > r (subtract),
// x_7 < r_7 means x < r (keep); only x_7 == r_7 (one value in 2^31 for the pseudo-random values of this path) needs the full-width
// comparison, behind a branch that is practically never taken.  One compare + 8 predicated subtractions on the ALU pipe instead of
// 8 subtractions, a borrow capture and 8 selects (which ptxas lowers to MOV / IMAD.MOV, the latter on the multiplier pipe).
#if defined(__CUDACC__)
#define HB_DEV_COLD __device__ __noinline__
#else
#define HB_DEV_COLD static
#endif
HB_DEV_COLD void cond_sub_mod_slow(uint32_t *x) {  // out of line: one copy of the rare path per kernel
    uint32_t r[8], d[8], v[8];
    mod_limbs(r);
    for (int i = 0; i < 8; ++i) v[i] = x[i];
    uint32_t br = sub8(d, v, r);
    if (!br) {
        for (int i = 0; i < 8; ++i) x[i] = d[i];
    }
}
HB_DEV void cond_sub_mod(uint32_t (&x)[8]) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.gt.u32 p, %7, " HB_PR7 ";\n\t"
        "@p sub.cc.u32 %0, %0, " HB_PR0 ";\n\t"
        "@p subc.cc.u32 %1, %1, " HB_PR1 ";\n\t"
        "@p subc.cc.u32 %2, %2, " HB_PR2 ";\n\t"
        "@p subc.cc.u32 %3, %3, " HB_PR3 ";\n\t"
        "@p subc.cc.u32 %4, %4, " HB_PR4 ";\n\t"
        "@p subc.cc.u32 %5, %5, " HB_PR5 ";\n\t"
        "@p subc.cc.u32 %6, %6, " HB_PR6 ";\n\t"
        "@p subc.u32 %7, %7, " HB_PR7 ";\n\t"
        "}"
        : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]));
    // after a subtraction x_7 < r_7 (x < 2r), so this test only fires for inputs with x_7 == r_7
    if (__builtin_expect(x[7] == HB_R7, 0)) {
        uint32_t tmp[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) tmp[i] = x[i];
        cond_sub_mod_slow(tmp);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = tmp[i];
    }
#else
    if (x[7] > HB_R7) {
        uint32_t r[8], d[8];
        mod_limbs(r);
        sub8(d, x, r);
        for (int i = 0; i < 8; ++i) x[i] = d[i];
    }
    if (x[7] == HB_R7) cond_sub_mod_slow(x);
#endif
}
// EXACT = true: branch-free variant (full-width borrow chain + selects, no rare path).  A few more ALU instructions, no branch and no
// convergence barrier: slower in the transforms (9.53 against 9.42 ms per 2^22, their lanes run in lockstep anyway), faster in the
// decoder's one-thread-per-codeword kernels, whose lanes diverge (Berlekamp-Massey: +4.5 % on the n = 128 leg).  robust.cuh selects it.
template <bool EXACT>
HB_DEV void cond_sub_mod_t(uint32_t (&x)[8]) {
#if defined(__CUDA_ARCH__)
    if (EXACT) {
        uint32_t d[8], br;
        asm("sub.cc.u32 %0, %9, " HB_PR0 ";\n\t"
            "subc.cc.u32 %1, %10, " HB_PR1 ";\n\t"
            "subc.cc.u32 %2, %11, " HB_PR2 ";\n\t"
            "subc.cc.u32 %3, %12, " HB_PR3 ";\n\t"
            "subc.cc.u32 %4, %13, " HB_PR4 ";\n\t"
            "subc.cc.u32 %5, %14, " HB_PR5 ";\n\t"
            "subc.cc.u32 %6, %15, " HB_PR6 ";\n\t"
            "subc.cc.u32 %7, %16, " HB_PR7 ";\n\t"
            "subc.u32 %8, 0, 0;"
            : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(br)
            : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]));
        const bool ge = br == 0;   // no borrow: x >= r
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = ge ? d[i] : x[i];
        return;
    }
#endif
    cond_sub_mod(x);
}
template <bool EXACT>
HB_DEV void fr_add_t(uint32_t (&d)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    add8(d, a, b);  // a,b < r < 2^255: no carry out
    cond_sub_mod_t<EXACT>(d);
}
HB_DEV void fr_add(uint32_t (&d)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { fr_add_t<false>(d, a, b); }
// a, b < r < 2^255: the 256-bit difference is negative exactly when its top bit is set (no borrow capture needed)
HB_DEV void fr_sub(uint32_t (&d)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "sub.cc.u32 %0, %8, %16;\n\t"
        "subc.cc.u32 %1, %9, %17;\n\t"
        "subc.cc.u32 %2, %10, %18;\n\t"
        "subc.cc.u32 %3, %11, %19;\n\t"
        "subc.cc.u32 %4, %12, %20;\n\t"
        "subc.cc.u32 %5, %13, %21;\n\t"
        "subc.cc.u32 %6, %14, %22;\n\t"
        "subc.u32 %7, %15, %23;\n\t"
        "setp.lt.s32 p, %7, 0;\n\t"
        "@p add.cc.u32 %0, %0, " HB_PR0 ";\n\t"
        "@p addc.cc.u32 %1, %1, " HB_PR1 ";\n\t"
        "@p addc.cc.u32 %2, %2, " HB_PR2 ";\n\t"
        "@p addc.cc.u32 %3, %3, " HB_PR3 ";\n\t"
        "@p addc.cc.u32 %4, %4, " HB_PR4 ";\n\t"
        "@p addc.cc.u32 %5, %5, " HB_PR5 ";\n\t"
        "@p addc.cc.u32 %6, %6, " HB_PR6 ";\n\t"
        "@p addc.u32 %7, %7, " HB_PR7 ";\n\t"
        "}"
        : "=&r"(d[0]), "=&r"(d[1]), "=&r"(d[2]), "=&r"(d[3]), "=&r"(d[4]), "=&r"(d[5]), "=&r"(d[6]), "=&r"(d[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]),
          "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint32_t br = sub8(d, a, b);
    if (br) {
        uint32_t r[8], e[8];
        mod_limbs(r);
        add8(e, d, r);
        for (int i = 0; i < 8; ++i) d[i] = e[i];
    }
#endif
}
HB_DEV bool fr_eq(const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x |= a[i] ^ b[i];
    return x == 0;
}
HB_DEV bool fr_is_zero(const uint32_t (&a)[8]) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x |= a[i];
    return x == 0;
}

// ----------------------------------------------------------------------------------------------
// acc_reduce: out = (sum of accumulated products) * R^{-1} mod r, fully reduced (< r).
// Valid for up to 256 accumulated products of factors < r (T < 2^518).
// ----------------------------------------------------------------------------------------------
template <bool EXACT>
HB_DEV void acc_reduce_t(const acc_t &A, uint32_t (&out)[8]) {
    uint32_t T[17];
    uint32_t AE[16], AO[16];  // 32-bit views: AE[p] / AO[p] = limb at position p
#pragma unroll
    for (int j = 0; j < 8; ++j) { AE[2 * j] = (uint32_t)A.E[j]; AE[2 * j + 1] = (uint32_t)(A.E[j] >> 32); }
    AO[0] = 0; AO[15] = 0;
#pragma unroll
    for (int j = 0; j < 7; ++j) { AO[2 * j + 1] = (uint32_t)A.O[j]; AO[2 * j + 2] = (uint32_t)(A.O[j] >> 32); }
    // merge E + O (O[p] sits at position p), then the carry counters K[8..16]
    T[0] = AE[0];
    uint32_t c;
    {
        uint32_t x[8], y[8], d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = AE[1 + i]; y[i] = AO[1 + i]; }
        c = add8c(d, x, y, 0);
#pragma unroll
        for (int i = 0; i < 8; ++i) T[1 + i] = d[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = (9 + i <= 15) ? AE[9 + i] : 0u; y[i] = (9 + i <= 14) ? AO[9 + i] : 0u; }
        add8c(d, x, y, c);
#pragma unroll
        for (int i = 0; i < 8; ++i) T[9 + i] = d[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = T[8 + i]; y[i] = A.K[8 + i]; }
        c = add8c(d, x, y, 0);
#pragma unroll
        for (int i = 0; i < 8; ++i) T[8 + i] = d[i];
        T[16] += A.K[16] + c;
    }
    // Montgomery reduction, 8 rows; carries into positions >= 8 are deferred (they never feed a quotient digit)
    uint32_t K2[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) K2[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        redc_row(T[i], T[i + 1], T[i + 2], T[i + 3], T[i + 4], T[i + 5], T[i + 6], T[i + 7], T[i + 8], K2[i + 8], K2[i + 9]);
    uint32_t U[8], U8;
    {
        uint32_t x[8], y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { x[i] = T[8 + i]; y[i] = K2[8 + i]; }
        c = add8c(U, x, y, 0);
        U8 = T[16] + K2[16] + c;
    }
    // U < (0.4529*terms + 1) * r < 2^262.  Quotient estimate from the 32 bits above 2^230:
    // D = floor(r / 2^230) + 1  =>  q_est in {q-1, q}; after U -= q_est*r one conditional subtraction remains.
    uint32_t u_top = (U[7] >> 6) | (U8 << 26);
    uint32_t q = u_top / 30389918u;
    uint64_t cc = 0;
    uint32_t P[8];
    const uint32_t rl[8] = {HB_R0, HB_R1, HB_R2, HB_R3, HB_R4, HB_R5, HB_R6, HB_R7};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        cc += (uint64_t)q * rl[i];
        P[i] = (uint32_t)cc;
        cc >>= 32;
    }
    // U - q*r < 2r < 2^256: arithmetic mod 2^256 is exact
    sub8(out, U, P);
    cond_sub_mod_t<EXACT>(out);
}
HB_DEV void acc_reduce(const acc_t &A, uint32_t (&out)[8]) { acc_reduce_t<false>(A, out); }

// Montgomery product through the lazy accumulator (kept for the unit test of acc_mac / acc_reduce)
HB_DEV void mont_mul_acc(uint32_t (&d)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    acc_t A;
    acc_zero(A);
    acc_mac(A, a, b);
    acc_reduce(A, d);
}

// ----------------------------------------------------------------------------------------------
// mont_mul: a*b*R^{-1} mod r for ONE product (butterflies, decoder), interleaved multiply/reduce rows.
// The running value T is kept as an even-aligned and an odd-aligned array of four 64-bit lanes (T = E + O*2^32); each row
// adds a*b_i with two IMAD.WIDE carry chains, cancels the low limb with m = -T_0 (again two chains with the modulus)
// and shifts by one limb by swapping the roles of the two arrays.  ~30 live registers instead of the ~50 of the lazy
// 512-bit accumulator, so kernels built on it keep twice as many warps resident.  Inputs < 2^256 with a*b < r*2^256.
// ----------------------------------------------------------------------------------------------
// row i >= 1:  (E: current even lanes with E0 == 0, O: current odd lanes)  ->  roles swapped (E holds the new odd lanes)
HB_DEV void cios_row(unsigned long long &e0, unsigned long long &e1, unsigned long long &e2, unsigned long long &e3,
                     unsigned long long &o0, unsigned long long &o1, unsigned long long &o2, unsigned long long &o3,
                     const uint32_t (&a)[8], uint32_t bi) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t"
        ".reg .u32 E0, E1, E2, E3, E4, E5, E6, E7, O0, O1, O2, O3, O4, O5, O6, O7, mi, mh, t0;\n\t"
        "mov.b64 {E0, E1}, %0;\n\t"
        "mov.b64 {E2, E3}, %1;\n\t"
        "mov.b64 {E4, E5}, %2;\n\t"
        "mov.b64 {E6, E7}, %3;\n\t"
        "mov.b64 {O0, O1}, %4;\n\t"
        "mov.b64 {O2, O3}, %5;\n\t"
        "mov.b64 {O4, O5}, %6;\n\t"
        "mov.b64 {O6, O7}, %7;\n\t"
        // shift by one limb: new even = old odd (+ old E1 at position 0), new odd[j] = old E[j+2] + a[2j+1]*bi
        "add.cc.u32 O0, O0, E1;\n\t"
        "madc.lo.cc.u32 E0, %9, %16, E2;\n\t"
        "madc.hi.cc.u32 E1, %9, %16, E3;\n\t"
        "madc.lo.cc.u32 E2, %11, %16, E4;\n\t"
        "madc.hi.cc.u32 E3, %11, %16, E5;\n\t"
        "madc.lo.cc.u32 E4, %13, %16, E6;\n\t"
        "madc.hi.cc.u32 E5, %13, %16, E7;\n\t"
        "madc.lo.cc.u32 E6, %15, %16, 0;\n\t"
        "madc.hi.u32 E7, %15, %16, 0;\n\t"
        "mad.lo.cc.u32 O0, %8, %16, O0;\n\t"
        "madc.hi.cc.u32 O1, %8, %16, O1;\n\t"
        "madc.lo.cc.u32 O2, %10, %16, O2;\n\t"
        "madc.hi.cc.u32 O3, %10, %16, O3;\n\t"
        "madc.lo.cc.u32 O4, %12, %16, O4;\n\t"
        "madc.hi.cc.u32 O5, %12, %16, O5;\n\t"
        "madc.lo.cc.u32 O6, %14, %16, O6;\n\t"
        "madc.hi.cc.u32 O7, %14, %16, O7;\n\t"
        "addc.u32 E7, E7, 0;\n\t"
        // Montgomery step: m = -T_0, T += m * r.  r_0 = 1 and r_1 = 2^32 - 1, so m*r_0 = m and
        // m*r_1 = (m - [m != 0]) * 2^32 + T_0: those two products are plain additions (ALU pipe), 6 wide multiplies remain
        "mov.u32 t0, O0;\n\t"
        "sub.u32 mi, 0, O0;\n\t"
        "min.u32 mh, mi, 1;\n\t"
        "sub.u32 mh, mi, mh;\n\t"
        "add.cc.u32 E0, E0, t0;\n\t"
        "addc.cc.u32 E1, E1, mh;\n\t"
        "madc.lo.cc.u32 E2, mi, " HB_PR3 ", E2;\n\t"
        "madc.hi.cc.u32 E3, mi, " HB_PR3 ", E3;\n\t"
        "madc.lo.cc.u32 E4, mi, " HB_PR5 ", E4;\n\t"
        "madc.hi.cc.u32 E5, mi, " HB_PR5 ", E5;\n\t"
        "madc.lo.cc.u32 E6, mi, " HB_PR7 ", E6;\n\t"
        "madc.hi.u32 E7, mi, " HB_PR7 ", E7;\n\t"
        "add.cc.u32 O0, O0, mi;\n\t"
        "addc.cc.u32 O1, O1, 0;\n\t"
        "madc.lo.cc.u32 O2, mi, " HB_PR2 ", O2;\n\t"
        "madc.hi.cc.u32 O3, mi, " HB_PR2 ", O3;\n\t"
        "madc.lo.cc.u32 O4, mi, " HB_PR4 ", O4;\n\t"
        "madc.hi.cc.u32 O5, mi, " HB_PR4 ", O5;\n\t"
        "madc.lo.cc.u32 O6, mi, " HB_PR6 ", O6;\n\t"
        "madc.hi.cc.u32 O7, mi, " HB_PR6 ", O7;\n\t"
        "addc.u32 E7, E7, 0;\n\t"
        "mov.b64 %0, {E0, E1};\n\t"
        "mov.b64 %1, {E2, E3};\n\t"
        "mov.b64 %2, {E4, E5};\n\t"
        "mov.b64 %3, {E6, E7};\n\t"
        "mov.b64 %4, {O0, O1};\n\t"
        "mov.b64 %5, {O2, O3};\n\t"
        "mov.b64 %6, {O4, O5};\n\t"
        "mov.b64 %7, {O6, O7};\n\t"
        "}"
        : "+l"(e0), "+l"(e1), "+l"(e2), "+l"(e3), "+l"(o0), "+l"(o1), "+l"(o2), "+l"(o3)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(bi));
#else
    // literal emulation of the instruction sequence above (32-bit limbs, explicit carry flag)
    uint32_t E[8] = {(uint32_t)e0, (uint32_t)(e0 >> 32), (uint32_t)e1, (uint32_t)(e1 >> 32), (uint32_t)e2, (uint32_t)(e2 >> 32), (uint32_t)e3, (uint32_t)(e3 >> 32)};
    uint32_t O[8] = {(uint32_t)o0, (uint32_t)(o0 >> 32), (uint32_t)o1, (uint32_t)(o1 >> 32), (uint32_t)o2, (uint32_t)(o2 >> 32), (uint32_t)o3, (uint32_t)(o3 >> 32)};
    const uint32_t rl[8] = {HB_R0, HB_R1, HB_R2, HB_R3, HB_R4, HB_R5, HB_R6, HB_R7};
    unsigned cf = 0;
    auto addc = [&](uint32_t x, uint64_t y, bool use_cf) { uint64_t t = (uint64_t)x + y + (use_cf ? cf : 0); cf = (unsigned)(t >> 32); return (uint32_t)t; };
    auto lo = [](uint32_t x, uint32_t y) { return (uint64_t)(uint32_t)((uint64_t)x * y); };
    auto hi = [](uint32_t x, uint32_t y) { return (uint64_t)(((uint64_t)x * y) >> 32); };
    O[0] = addc(O[0], E[1], false);
    for (int j = 0; j < 3; ++j) { uint32_t x = a[2 * j + 1]; E[2 * j] = addc(E[2 * j + 2], lo(x, bi), true); E[2 * j + 1] = addc(E[2 * j + 3], hi(x, bi), true); }
    E[6] = addc(0, lo(a[7], bi), true);
    E[7] = addc(0, hi(a[7], bi), true);
    for (int j = 0; j < 4; ++j) { uint32_t x = a[2 * j]; O[2 * j] = addc(O[2 * j], lo(x, bi), j > 0); O[2 * j + 1] = addc(O[2 * j + 1], hi(x, bi), true); }
    E[7] = addc(E[7], 0, true);
    const uint32_t t0 = O[0], mi = 0u - O[0], mh = mi - (mi < 1u ? mi : 1u);
    E[0] = addc(E[0], t0, false);
    E[1] = addc(E[1], mh, true);
    for (int j = 1; j < 4; ++j) { uint32_t x = rl[2 * j + 1]; E[2 * j] = addc(E[2 * j], lo(mi, x), true); E[2 * j + 1] = addc(E[2 * j + 1], hi(mi, x), true); }
    O[0] = addc(O[0], mi, false);
    O[1] = addc(O[1], 0, true);
    for (int j = 1; j < 4; ++j) { uint32_t x = rl[2 * j]; O[2 * j] = addc(O[2 * j], lo(mi, x), true); O[2 * j + 1] = addc(O[2 * j + 1], hi(mi, x), true); }
    E[7] = addc(E[7], 0, true);
    e0 = E[0] | ((unsigned long long)E[1] << 32); e1 = E[2] | ((unsigned long long)E[3] << 32);
    e2 = E[4] | ((unsigned long long)E[5] << 32); e3 = E[6] | ((unsigned long long)E[7] << 32);
    o0 = O[0] | ((unsigned long long)O[1] << 32); o1 = O[2] | ((unsigned long long)O[3] << 32);
    o2 = O[4] | ((unsigned long long)O[5] << 32); o3 = O[6] | ((unsigned long long)O[7] << 32);
#endif
}

template <bool EXACT>
HB_DEV void mont_mul_t(uint32_t (&d)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
    // row 0 is the general row on a zero accumulator (ptxas folds the zero addends into plain IMAD.WIDE): 14 wide multiply-adds.
    // (A hand-specialised row 0 -- products first, then m*r over 32-bit views -- was lowered to 8 IMAD.HI + 6 IMAD.X pairs.)
    unsigned long long E[4] = {0ull, 0ull, 0ull, 0ull}, O[4] = {0ull, 0ull, 0ull, 0ull};
    cios_row(E[0], E[1], E[2], E[3], O[0], O[1], O[2], O[3], a, b[0]);  // now O = even, E = odd
    cios_row(O[0], O[1], O[2], O[3], E[0], E[1], E[2], E[3], a, b[1]);  // E = even
    cios_row(E[0], E[1], E[2], E[3], O[0], O[1], O[2], O[3], a, b[2]);
    cios_row(O[0], O[1], O[2], O[3], E[0], E[1], E[2], E[3], a, b[3]);
    cios_row(E[0], E[1], E[2], E[3], O[0], O[1], O[2], O[3], a, b[4]);
    cios_row(O[0], O[1], O[2], O[3], E[0], E[1], E[2], E[3], a, b[5]);
    cios_row(E[0], E[1], E[2], E[3], O[0], O[1], O[2], O[3], a, b[6]);
    cios_row(O[0], O[1], O[2], O[3], E[0], E[1], E[2], E[3], a, b[7]);  // E = even (E0 low limb == 0), O = odd
    // result = (even >> 32) + odd  (< 2r), then one conditional subtraction
    uint32_t x[8], y[8], sres[8];
    x[0] = (uint32_t)(E[0] >> 32); x[1] = (uint32_t)E[1]; x[2] = (uint32_t)(E[1] >> 32); x[3] = (uint32_t)E[2];
    x[4] = (uint32_t)(E[2] >> 32); x[5] = (uint32_t)E[3]; x[6] = (uint32_t)(E[3] >> 32); x[7] = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { y[2 * j] = (uint32_t)O[j]; y[2 * j + 1] = (uint32_t)(O[j] >> 32); }
    add8(sres, x, y);
    cond_sub_mod_t<EXACT>(sres);
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = sres[i];
}
HB_DEV void mont_mul(uint32_t (&d)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) { mont_mul_t<false>(d, a, b); }

// Two independent products with their rows issued alternately: twice the independent carry chains in flight per thread (a warp
// issues one IMAD.WIDE of a single chain every ~8 cycles; the multiplier pipe accepts one every ~4).  Bit-identical to two mont_mul.
HB_DEV void mont_mul2(uint32_t (&d0)[8], const uint32_t (&a0)[8], const uint32_t (&b0)[8], uint32_t (&d1)[8], const uint32_t (&a1)[8],
                      const uint32_t (&b1)[8]) {
    unsigned long long E0[4], O0[4], E1[4], O1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        E0[j] = (unsigned long long)a0[2 * j] * b0[0];
        O0[j] = (unsigned long long)a0[2 * j + 1] * b0[0];
        E1[j] = (unsigned long long)a1[2 * j] * b1[0];
        O1[j] = (unsigned long long)a1[2 * j + 1] * b1[0];
    }
    {
        const uint32_t m0 = 0u - (uint32_t)E0[0], m1 = 0u - (uint32_t)E1[0];
        uint32_t k0 = 0, k1 = 0, k2 = 0, k3 = 0;
        chain4w(O0[0], O0[1], O0[2], O0[3], k0, HB_R1, HB_R3, HB_R5, HB_R7, m0);
        chain4w(O1[0], O1[1], O1[2], O1[3], k2, HB_R1, HB_R3, HB_R5, HB_R7, m1);
        chain4w(E0[0], E0[1], E0[2], E0[3], k1, HB_R0, HB_R2, HB_R4, HB_R6, m0);
        chain4w(E1[0], E1[1], E1[2], E1[3], k3, HB_R0, HB_R2, HB_R4, HB_R6, m1);
        O0[3] += (unsigned long long)k1 << 32;
        O1[3] += (unsigned long long)k3 << 32;
        (void)k0; (void)k2;
    }
    cios_row(E0[0], E0[1], E0[2], E0[3], O0[0], O0[1], O0[2], O0[3], a0, b0[1]);
    cios_row(E1[0], E1[1], E1[2], E1[3], O1[0], O1[1], O1[2], O1[3], a1, b1[1]);
    cios_row(O0[0], O0[1], O0[2], O0[3], E0[0], E0[1], E0[2], E0[3], a0, b0[2]);
    cios_row(O1[0], O1[1], O1[2], O1[3], E1[0], E1[1], E1[2], E1[3], a1, b1[2]);
    cios_row(E0[0], E0[1], E0[2], E0[3], O0[0], O0[1], O0[2], O0[3], a0, b0[3]);
    cios_row(E1[0], E1[1], E1[2], E1[3], O1[0], O1[1], O1[2], O1[3], a1, b1[3]);
    cios_row(O0[0], O0[1], O0[2], O0[3], E0[0], E0[1], E0[2], E0[3], a0, b0[4]);
    cios_row(O1[0], O1[1], O1[2], O1[3], E1[0], E1[1], E1[2], E1[3], a1, b1[4]);
    cios_row(E0[0], E0[1], E0[2], E0[3], O0[0], O0[1], O0[2], O0[3], a0, b0[5]);
    cios_row(E1[0], E1[1], E1[2], E1[3], O1[0], O1[1], O1[2], O1[3], a1, b1[5]);
    cios_row(O0[0], O0[1], O0[2], O0[3], E0[0], E0[1], E0[2], E0[3], a0, b0[6]);
    cios_row(O1[0], O1[1], O1[2], O1[3], E1[0], E1[1], E1[2], E1[3], a1, b1[6]);
    cios_row(E0[0], E0[1], E0[2], E0[3], O0[0], O0[1], O0[2], O0[3], a0, b0[7]);
    cios_row(E1[0], E1[1], E1[2], E1[3], O1[0], O1[1], O1[2], O1[3], a1, b1[7]);
    auto fin = [](uint32_t (&d)[8], const unsigned long long (&E)[4], const unsigned long long (&O)[4]) {
        uint32_t x[8], y[8], sres[8];
        x[0] = (uint32_t)(O[0] >> 32); x[1] = (uint32_t)O[1]; x[2] = (uint32_t)(O[1] >> 32); x[3] = (uint32_t)O[2];
        x[4] = (uint32_t)(O[2] >> 32); x[5] = (uint32_t)O[3]; x[6] = (uint32_t)(O[3] >> 32); x[7] = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) { y[2 * j] = (uint32_t)E[j]; y[2 * j + 1] = (uint32_t)(E[j] >> 32); }
        add8(sres, x, y);
        cond_sub_mod(sres);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = sres[i];
    };
    fin(d0, E0, O0);
    fin(d1, E1, O1);
}

// ----------------------------------------------------------------------------------------------
// fma_ballast: steering ptxas' pipe balancing (measured, CUDA 12.9 ptxas for sm_100a).
// ptxas levels the STATIC instruction counts of the ALU pipe and of the multiplier (FMA) pipe over a whole kernel: when the ALU
// side is longer it lowers moves, plain additions, shifts and carry captures to IMAD.MOV / IMAD.IADD / IMAD.SHL / IMAD.X.  It counts
// an IMAD.WIDE like any other instruction, but on B200 the wide multiply-add occupies the multiplier pipe for twice the cycles of a
// plain IMAD and that pipe is what bounds every kernel of this library: the lowered instructions were 19 % of its busy cycles
// (profiles/r02g_exec_ntt.json: 52 per Montgomery product).  A block of FFMAs that is never executed (`never` is a run-time value
// that is false on every call: code size only, the hot path does not even fetch it) tips the static balance, and ptxas keeps those
// instructions on the ALU pipe: IMAD.MOV 192 -> 0, IMAD.IADD 34 -> 0, IMAD.SHL 13 -> 0, IMAD.X 53 -> 12 in ntt16x_kernel<6,0>
// (tools/sass_hist.py checks the shipped library).  HB_FMA_BALLAST_N = 0 compiles it away.
// ----------------------------------------------------------------------------------------------
#ifndef HB_FMA_BALLAST_N
#define HB_FMA_BALLAST_N 512
#endif
#if defined(__CUDACC__)
__device__ __forceinline__ void fma_ballast(bool never, unsigned int *sink) {
#if HB_FMA_BALLAST_N > 0
    if (never) {
        float f = __uint_as_float(sink[0]), g = __uint_as_float(sink[1]);
#pragma unroll
        for (int i = 0; i < HB_FMA_BALLAST_N; ++i) {
            f = fmaf(f, g, 1.0f + i);
            g = fmaf(g, f, 2.0f + i);
        }
        sink[0] = __float_as_uint(f + g);
    }
#else
    (void)never; (void)sink;
#endif
}
#endif

#if defined(__CUDACC__)
// fail[b] = 1, and -- exactly once per item, whichever thread gets there first -- b is appended to list[(*count)++] (list may be
// nullptr).  The checking kernels build the list of failing items themselves: no compaction pass over fail[] on the call path.
// (fail buffers are zero-filled and padded to a multiple of 4 bytes by the host side.)
__device__ __forceinline__ void mark_fail(unsigned char *fail, long long b, unsigned int *list, unsigned int *count) {
    const unsigned long long addr = (unsigned long long)(fail + b);
    unsigned int *w = reinterpret_cast<unsigned int *>(addr & ~3ull);
    const unsigned int bit = 1u << (8u * (unsigned int)(addr & 3ull));
    const unsigned int old = atomicOr(w, bit);
    if (!(old & bit) && list) list[atomicAdd(count, 1u)] = (unsigned int)b;
}
#endif

// R^2 mod r (to Montgomery form: mont_mul(x, R2)); 1 (from Montgomery form: mont_mul(x, 1))
HB_DEV void r2_limbs(uint32_t (&r)[8]) {
    r[0] = 0xf3f29c6du; r[1] = 0xc999e990u; r[2] = 0x87925c23u; r[3] = 0x2b6cedcbu;
    r[4] = 0x7254398fu; r[5] = 0x05d31496u; r[6] = 0x9f59ff11u; r[7] = 0x0748d9d9u;
}
// R mod r == Montgomery form of 1
HB_DEV void one_mont_limbs(uint32_t (&r)[8]) {
    r[0] = 0xfffffffeu; r[1] = 0x00000001u; r[2] = 0x00034802u; r[3] = 0x5884b7fau;
    r[4] = 0xecbc4ff5u; r[5] = 0x998c4fefu; r[6] = 0xacc5056fu; r[7] = 0x1824b159u;
}

// ---- K5 fused (elementwise_fused_kernel).  Canonical values in, canonical values out.
// z = a*b - c                                   (triple_generation.rs:332-340: share_mul, then Sub)
HB_DEV void k5_triple_mask(uint32_t (&z)[8], const uint32_t (&a)[8], const uint32_t (&b)[8], const uint32_t (&c)[8]) {
    uint32_t r2[8], am[8], p[8];
    r2_limbs(r2);
    mont_mul(am, a, r2);   // a*R
    mont_mul(p, am, b);    // a*b, canonical
    fr_sub(z, p, c);
}
// z = c - da*db - da*y - db*x                   (multiplication.rs:79-97: three Mul, three Sub)
// computed as c - [da*(db + y) + db*x]: the bracket is ONE lazily accumulated sum of two products (one reduction), one more product
// leaves the Montgomery domain -- about 2.7 products' worth of wide multiplies instead of the 5 of the literal formula, with which the
// pass is multiplier-bound instead of HBM-bound (tools/k5_fused_probe.py).  Field identities only, so the canonical result is the one
// the reference's operator sequence produces.
HB_DEV void k5_beaver_finalize(uint32_t (&z)[8], const uint32_t (&c)[8], const uint32_t (&x)[8], const uint32_t (&y)[8], const uint32_t (&da)[8],
                               const uint32_t (&db)[8]) {
    uint32_t r2[8], s[8], w[8], u[8];
    r2_limbs(r2);
    fr_add(s, db, y);
    acc_t A;
    acc_zero(A);
    acc_mac(A, da, s);
    acc_mac(A, db, x);
    acc_reduce(A, w);      // (...)/R, fully reduced
    mont_mul(u, w, r2);    // canonical
    fr_sub(z, c, u);
}

HB_DEV void load_fr(uint32_t (&a)[8], const uint4 lo, const uint4 hi) {
    a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w; a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
}

}  // namespace hb
