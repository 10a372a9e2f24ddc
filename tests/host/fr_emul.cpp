// Host emulation harness for mpc-protocols_b200/csrc/fr.cuh (TEST ONLY, see HB_HOST_EMULATION there).
// Prints, per case:  <terms> <a_0> <b_0> ... <a_{terms-1}> <b_{terms-1}> <acc_reduce result>   (hex, canonical ints)
#define HB_HOST_EMULATION 1
#include "../../mpc-protocols_b200/csrc/fr.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
static uint64_t sm_state;
static uint64_t splitmix() {
    uint64_t z = (sm_state += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static const uint32_t MODL[8] = {HB_R0, HB_R1, HB_R2, HB_R3, HB_R4, HB_R5, HB_R6, HB_R7};
static bool lt_mod(const uint32_t *a) {
    for (int i = 7; i >= 0; --i) { if (a[i] < MODL[i]) return true; if (a[i] > MODL[i]) return false; }
    return false;
}
static void rnd(uint32_t (&a)[8], int kind) {
    if (kind == 1) { for (int i = 0; i < 8; ++i) a[i] = MODL[i]; a[0] -= 1; return; }  // r-1
    if (kind == 2) { for (int i = 0; i < 8; ++i) a[i] = 0; return; }
    do { for (int i = 0; i < 8; i += 2) { uint64_t v = splitmix(); a[i] = (uint32_t)v; a[i + 1] = (uint32_t)(v >> 32); } a[7] &= 0x7fffffffu; } while (!lt_mod(a));
}
static void print(const uint32_t *a) { for (int i = 7; i >= 0; --i) printf("%08x", a[i]); }
int main(int argc, char **argv) {
    sm_state = argc > 1 ? strtoull(argv[1], 0, 0) : 1;
    const int terms_list[] = {1, 1, 2, 3, 22, 43, 64, 128, 256, 256};
    for (unsigned t = 0; t < sizeof(terms_list) / sizeof(int); ++t) {
        int terms = terms_list[t];
        int kind = (t == 1 || t == 9) ? 1 : 0;  // worst case: all factors r-1
        hb::acc_t A; hb::acc_zero(A);
        printf("%d", terms);
        for (int k = 0; k < terms; ++k) {
            uint32_t a[8], b[8];
            rnd(a, kind); rnd(b, (kind == 0 && k % 7 == 3) ? 2 : kind);
            hb::acc_mac(A, a, b);
            printf(" "); print(a); printf(" "); print(b);
        }
        uint32_t out[8];
        hb::acc_reduce(A, out);
        printf(" "); print(out); printf("\n");
    }
    // mont_mul / add / sub
    for (int t = 0; t < 64; ++t) {
        uint32_t a[8], b[8], m[8], s[8], d[8];
        rnd(a, t == 0 ? 1 : 0); rnd(b, t == 1 ? 2 : (t == 0 ? 1 : 0));
        hb::mont_mul_acc(m, a, b); hb::fr_add(s, a, b); hb::fr_sub(d, a, b);
        uint32_t mc[8]; hb::mont_mul(mc, a, b);
        printf("ops "); print(a); printf(" "); print(b); printf(" "); print(m); printf(" "); print(s); printf(" "); print(d); printf(" "); print(mc); printf("\n");
    }
    // K5 fused: triple mask and Beaver product share (edge operands in the first cases)
    for (int t = 0; t < 48; ++t) {
        uint32_t v[5][8], m[8], f[8];
        for (int k = 0; k < 5; ++k) rnd(v[k], t < 5 ? (k == t ? 1 : 2) : (t == 5 ? 1 : (t == 6 ? 2 : 0)));
        hb::k5_triple_mask(m, v[0], v[1], v[2]);
        hb::k5_beaver_finalize(f, v[0], v[1], v[2], v[3], v[4]);
        printf("k5");
        for (int k = 0; k < 5; ++k) { printf(" "); print(v[k]); }
        printf(" "); print(m); printf(" "); print(f); printf("\n");
    }
    return 0;
}
