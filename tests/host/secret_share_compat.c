/* secret_share_compat.c -- the call sequences of the reference's C consumer test (mpc/src/ffi/tests/secret_share.c:8-170: Shamir on
 * ids 1..6, RobustShare and NonRobustShare on the domain; the same fixed secrets) against include/hbmpc_compat_share.h, i.e. against
 * the reference's OWN symbol names served by libhbmpc_b200.so.  In a checkout that has the reference, tests/test_gpu_host_mirror.py
 * additionally compiles the unmodified secret_share.c against the reference header and links it with this library. */
#include <assert.h>
#include <stdio.h>
#include <string.h>

#include "hbmpc_compat_share.h"

static ShareErrorCode shamir_case(void) {
    struct U256 secret_fr = {{520, 86, 9, 18}};
    struct ShamirShareSlice output_shares;
    uintptr_t id[6] = {1, 2, 3, 4, 5, 6};
    struct UsizeSlice ids = {id, 6};
    ShamirShare s = shamir_share_new(secret_fr, 9, 10, Bls12_381Fr);
    free_shamir_share(s);
    ShareErrorCode e = shamir_share_compute_shares(secret_fr, 4, &ids, Bls12_381Fr, &output_shares);
    if (e != ShareSuccess) return e;
    for (uintptr_t i = 0; i < output_shares.len; i++) {
        ByteSlice bytes = field_ptr_to_bytes(output_shares.pointer[i].share, true);
        U256 u = be_bytes_to_u256(bytes);
        free_bytes_slice(bytes);
        ByteSlice le = field_ptr_to_bytes(output_shares.pointer[i].share, false);
        U256 v = le_bytes_to_u256(le);
        free_bytes_slice(le);
        if (memcmp(u.data, v.data, 32) != 0 || output_shares.pointer[i].id != id[i] || output_shares.pointer[i].degree != 4) return PolynomialOperationError;
    }
    struct U256 recovered_secret;
    struct U256Slice recovered_coeff;
    e = shamir_share_recover_secret(output_shares, &recovered_secret, &recovered_coeff, Bls12_381Fr);
    if (e != ShareSuccess) return e;
    if (memcmp(recovered_secret.data, secret_fr.data, 32) != 0 || memcmp(recovered_coeff.pointer[0].data, secret_fr.data, 32) != 0) return DecodingError;
    /* a wrong degree claim: 6 points of a degree-4 polynomial claimed as degree 3 */
    for (uintptr_t i = 0; i < output_shares.len; i++) output_shares.pointer[i].degree = 3;
    free_u256_slice(recovered_coeff);
    e = shamir_share_recover_secret(output_shares, &recovered_secret, &recovered_coeff, Bls12_381Fr);
    if (e != DegreeMismatch) return PolynomialOperationError;
    free_shamir_share_slice(output_shares);
    return ShareSuccess;
}

static ShareErrorCode robust_case(void) {
    struct U256 secret_fr = {{3, 3, 22, 22}};
    struct RobustShareSlice output_shares;
    uintptr_t n = 6;
    RobustShare s = robust_share_new(secret_fr, 9, 10, Bls12_381Fr);
    free_robust_share(s);
    ShareErrorCode e = robust_share_compute_shares(secret_fr, 2, n, &output_shares, Bls12_381Fr);
    if (e != ShareSuccess) return e;
    if (output_shares.len != n) return InvalidInput;
    struct U256 recovered_secret;
    struct U256Slice recovered_coeff;
    e = robust_share_recover_secret(output_shares, n, 1, &recovered_secret, &recovered_coeff, Bls12_381Fr);
    if (e != ShareSuccess) return e;
    if (memcmp(recovered_secret.data, secret_fr.data, 32) != 0 || memcmp(recovered_coeff.pointer[0].data, secret_fr.data, 32) != 0) return DecodingError;
    if (recovered_coeff.len > 3) return PolynomialOperationError;
    free_u256_slice(recovered_coeff);
    /* n <= degree */
    struct RobustShareSlice none;
    if (robust_share_compute_shares(secret_fr, 6, 6, &none, Bls12_381Fr) != InvalidInput) return PolynomialOperationError;
    free_robust_share_slice(output_shares);
    return ShareSuccess;
}

static ShareErrorCode non_robust_case(void) {
    struct U256 secret_fr = {{16, 33, 44, 81}};
    struct NonRobustShareSlice output_shares;
    uintptr_t n = 6;
    NonRobustShare s = non_robust_share_new(secret_fr, 9, 10, Bls12_381Fr);
    free_non_robust_share(s);
    ShareErrorCode e = non_robust_share_compute_shares(secret_fr, 2, n, &output_shares, Bls12_381Fr);
    if (e != ShareSuccess) return e;
    struct U256 recovered_secret;
    struct U256Slice recovered_coeff;
    e = non_robust_share_recover_secret(output_shares, n, &recovered_secret, &recovered_coeff, Bls12_381Fr);
    if (e != ShareSuccess) return e;
    if (memcmp(recovered_secret.data, secret_fr.data, 32) != 0 || memcmp(recovered_coeff.pointer[0].data, secret_fr.data, 32) != 0) return DecodingError;
    free_u256_slice(recovered_coeff);
    free_non_robust_share_slice(output_shares);
    return ShareSuccess;
}

int main(void) {
    ShareErrorCode e;
    if ((e = shamir_case()) != ShareSuccess) { printf("shamir case failed: %d (library status %d)\n", (int)e, hbmpc_compat_last_status()); return hbmpc_compat_last_status() == 100 ? 100 : 1; }
    if ((e = robust_case()) != ShareSuccess) { printf("robust case failed: %d\n", (int)e); return 2; }
    if ((e = non_robust_case()) != ShareSuccess) { printf("non-robust case failed: %d\n", (int)e); return 3; }
    printf("reference share symbols over the B200 library: all round trips passed\n");
    return 0;
}
