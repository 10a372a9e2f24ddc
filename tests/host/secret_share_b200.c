/* secret_share_b200.c -- the reference's C-ABI consumer test (mpc/src/ffi/tests/secret_share.c) restated against
 * include/hbmpc_b200.h in plain C99: fixed U256 secrets {520,86,9,18}, {3,3,22,22}, {16,33,44,81} are shared
 * (n = 6, degree 2), recovered robustly (t = 1) and non-robustly, and compared with memcmp on the limbs
 * (secret_share.c:10,55-56,66,111-112,122,167-168).  The random polynomial coefficients come from the caller here
 * (the batch ABI has no hidden RNG).  Built with gcc and run on the GPU box by tests/test_gpu_host_mirror.py. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "hbmpc_b200.h"

static uint64_t sm = 0x2545F4914F6CDD1DULL;
static uint64_t next_u64(void) {
    uint64_t z = (sm += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static int roundtrip(hbmpc_ctx *ctx, const uint64_t secret[4]) {
    const size_t n = 6, d = 2, t = 1;
    uint64_t coeffs[3][4], shares[6][4], rec[3][4], out_secret[4];
    size_t ids[6] = {0, 1, 2, 3, 4, 5};
    int32_t path = -1, status = -1;
    memcpy(coeffs[0], secret, 32);
    for (int k = 1; k < 3; ++k) {
        for (int l = 0; l < 4; ++l) coeffs[k][l] = next_u64();
        coeffs[k][3] >>= 2; /* < 2^254 < r */
    }
    if (hbmpc_compute_shares_batch(ctx, n, d, 1, &coeffs[0][0], &shares[0][0]) != HBMPC_SUCCESS) return 1;
    /* robust_share_recover_secret */
    if (hbmpc_robust_interpolate_batch(ctx, n, d, t, n, ids, 1, &shares[0][0], &rec[0][0], out_secret, &path, NULL) != HBMPC_SUCCESS) return 2;
    if (memcmp(out_secret, secret, 32) != 0 || memcmp(rec, coeffs, sizeof coeffs) != 0 || path != 0) return 3;
    /* one corrupted share inside the examined prefix (ids 0..3) is corrected in OEC round 1 */
    shares[1][1] ^= 0x10;
    if (hbmpc_robust_interpolate_batch(ctx, n, d, t, n, ids, 1, &shares[0][0], &rec[0][0], out_secret, &path, NULL) != HBMPC_SUCCESS) return 4;
    if (memcmp(out_secret, secret, 32) != 0 || path != 1) return 5;
    shares[1][1] ^= 0x10;
    /* non_robust_share_recover_secret */
    if (hbmpc_nonrobust_recover_batch(ctx, n, d, n, ids, 1, &shares[0][0], 0, &rec[0][0], out_secret, &status) != HBMPC_SUCCESS) return 6;
    if (memcmp(out_secret, secret, 32) != 0 || status != (int32_t)d) return 7;
    return 0;
}

int main(void) {
    hbmpc_ctx *ctx = NULL;
    if (hbmpc_ctx_create(0, &ctx) != HBMPC_SUCCESS) { printf("no device\n"); return 100; }
    const uint64_t secrets[3][4] = {{520, 86, 9, 18}, {3, 3, 22, 22}, {16, 33, 44, 81}};
    for (int i = 0; i < 3; ++i) {
        int rc = roundtrip(ctx, secrets[i]);
        if (rc) { printf("FAIL secret %d step %d\n", i, rc); return rc; }
    }
    hbmpc_ctx_destroy(ctx);
    printf("secret_share_b200: all C-ABI round trips passed\n");
    return 0;
}
