/* group_test.c -- single-process multi-GPU entry points (hbmpc_group_*) from plain C: the batch is split into contiguous ranges over
 * the member contexts (one host thread per device, no collective) and must come back bit-identical to the same call on ONE context.
 * usage: group_test [n_devices]   (default: every visible device, at least two members -- on a 1-GPU box both sit on device 0) */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hbmpc_b200.h"

static uint64_t sm = 0x9E3779B97F4A7C15ULL;
static uint64_t next_u64(void) {
    uint64_t z = (sm += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
#define CHECK(c, msg) do { if (!(c)) { printf("FAIL %s (line %d)\n", msg, __LINE__); return 1; } } while (0)

int main(int argc, char **argv) {
    int members = argc > 1 ? atoi(argv[1]) : 0;
    hbmpc_ctx *one = NULL;
    if (hbmpc_ctx_create(0, &one) != HBMPC_SUCCESS) { printf("no device\n"); return 100; }
    int devices[8], visible = 0;
    for (int d = 0; d < 8; ++d) {   /* probe the visible devices */
        hbmpc_ctx *c = NULL;
        if (hbmpc_ctx_create(d, &c) != HBMPC_SUCCESS) break;
        hbmpc_ctx_destroy(c);
        ++visible;
    }
    if (members <= 0) members = visible < 2 ? 2 : visible;
    for (int i = 0; i < members; ++i) devices[i] = i % visible;
    hbmpc_group *grp = NULL;
    CHECK(hbmpc_group_create(devices, (size_t)members, &grp) == HBMPC_SUCCESS, "group_create");
    CHECK(hbmpc_group_size(grp) == (size_t)members, "group_size");

    const size_t n = 64, t = 21, d = 21, m = d + 1, B = 10007;   /* a prime: ragged ranges */
    uint64_t *coeffs = malloc(B * m * 32), *sh1 = malloc(B * n * 32), *shg = malloc(B * n * 32);
    uint64_t *ev = malloc(n * B * 32), *evg = malloc(n * B * 32);
    uint64_t *co1 = malloc(B * m * 32), *cog = malloc(B * m * 32), *fl1 = malloc(B * 8), *flg = malloc(B * 8), *se1 = malloc(B * 32), *seg = malloc(B * 32);
    int32_t *p1 = malloc(B * 4), *pg = malloc(B * 4);
    for (size_t i = 0; i < B * m; ++i) {
        for (int l = 0; l < 4; ++l) coeffs[4 * i + l] = next_u64();
        coeffs[4 * i + 3] >>= 2;
    }
    /* K1 */
    CHECK(hbmpc_compute_shares_batch(one, n, d, B, coeffs, sh1) == HBMPC_SUCCESS, "K1 single");
    CHECK(hbmpc_group_compute_shares_batch(grp, n, d, B, coeffs, shg) == HBMPC_SUCCESS, "K1 group");
    CHECK(memcmp(sh1, shg, B * n * 32) == 0, "K1 group == single context");
    /* K2 recipient-major (the batch axis is the INNER axis of the output: sharded through the leading dimension) */
    CHECK(hbmpc_apply_vandermonde_batch(one, n, m, B, coeffs, ev, 1) == HBMPC_SUCCESS, "K2 single");
    CHECK(hbmpc_group_apply_vandermonde_batch(grp, n, m, B, coeffs, evg, 1) == HBMPC_SUCCESS, "K2 group");
    CHECK(memcmp(ev, evg, n * B * 32) == 0, "K2 recipient-major group == single context");
    /* K3 with flags on corrupted sender vectors (sender-major input sharded along the batch axis) */
    size_t ids[64];
    for (size_t j = 0; j < n; ++j) ids[j] = (j * 37 + 11) % n;            /* an arrival order */
    for (size_t j = 0; j < n; ++j) memcpy(evg + j * B * 4, ev + ids[j] * B * 4, B * 32);
    for (size_t b = 0; b < B; b += 5) evg[(b % n) * B * 4 + b * 4] ^= 0x55;   /* every 5th chunk: one corrupted share */
    CHECK(hbmpc_batch_recover(one, n, d, t, n, ids, B, evg, co1, p1, fl1) == HBMPC_SUCCESS, "K3 single");
    CHECK(hbmpc_group_batch_recover(grp, n, d, t, n, ids, B, evg, cog, pg, flg) == HBMPC_SUCCESS, "K3 group");
    CHECK(memcmp(co1, cog, B * m * 32) == 0 && memcmp(p1, pg, B * 4) == 0 && memcmp(fl1, flg, B * 8) == 0, "K3 group == single context");
    CHECK(memcmp(cog, coeffs, B * m * 32) == 0, "K3 recovers the polynomials");
    CHECK(hbmpc_batch_recover_secrets(one, n, d, t, n, ids, B, evg, se1, p1) == HBMPC_SUCCESS, "K3 secrets single");
    CHECK(hbmpc_group_batch_recover_secrets(grp, n, d, t, n, ids, B, evg, seg, pg) == HBMPC_SUCCESS, "K3 secrets group");
    CHECK(memcmp(se1, seg, B * 32) == 0 && memcmp(p1, pg, B * 4) == 0, "K3 secrets group == single context");
    /* K4 codeword-major */
    for (size_t j = 0; j < n; ++j) ids[j] = j;
    sh1[7 * n * 4 + 3 * 4] ^= 1; sh1[7 * n * 4 + 9 * 4 + 1] ^= 2;
    CHECK(hbmpc_robust_interpolate_batch(one, n, d, t, n, ids, B, sh1, co1, se1, p1, fl1) == HBMPC_SUCCESS, "K4 single");
    CHECK(hbmpc_group_robust_interpolate_batch(grp, n, d, t, n, ids, B, sh1, cog, seg, pg, flg) == HBMPC_SUCCESS, "K4 group");
    CHECK(memcmp(co1, cog, B * m * 32) == 0 && memcmp(se1, seg, B * 32) == 0 && memcmp(p1, pg, B * 4) == 0 && memcmp(fl1, flg, B * 8) == 0, "K4 group == single context");
    CHECK(pg[7] == 2 || pg[7] == 1 || pg[7] == 0, "path of the corrupted codeword");
    /* ranges */
    size_t lo, hi, prev = 0;
    for (int i = 0; i < members; ++i) {
        hbmpc_group_shard_range(grp, B, (size_t)i, &lo, &hi);
        CHECK(lo == prev && hi >= lo, "contiguous ranges");
        prev = hi;
    }
    CHECK(prev == B, "ranges cover the batch");
    /* whole-call validation is the single-context one */
    CHECK(hbmpc_group_compute_shares_batch(grp, 4, 4, 10, coeffs, shg) == HBMPC_INVALID_INPUT, "n <= d");
    hbmpc_group_destroy(grp);
    hbmpc_ctx_destroy(one);
    printf("group of %d member contexts over %d device(s): all calls identical to the single-context results\n", members, visible);
    return 0;
}
