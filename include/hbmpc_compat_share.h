/*
 * hbmpc_compat_share.h -- the share section of the reference's own C ABI, served by the B200 library.
 *
 * These are the symbols of /root/reference/mpc/src/ffi/c_bindings/share/mod.rs:55-597 and c_bindings/mod.rs:139-200 (header
 * honey_badger_bindings.h:113-132,214-245,555-625) with the reference's names, argument order, structure layouts, error codes and
 * ownership rules (outputs are callee-allocated and returned through the free_* helpers), so that a C consumer of the reference --
 * mpc/src/ffi/tests/secret_share.c -- links against libhbmpc_b200.so unchanged.  Each call is a thin wrapper over the batch entry
 * points of hbmpc_b200.h with B = 1 on a process-wide context (device 0, created on first use); like the reference they draw the
 * polynomial from the thread's RNG (share/mod.rs:418).  There is no CPU path: without a usable GPU every call that computes
 * returns InvalidInput-free failure code PolynomialOperationError (7) and `hbmpc_compat_last_status()` holds HBMPC_NO_DEVICE.
 * A program that includes the reference's generated header must not include this one as well (same type names).
 */
#ifndef HBMPC_COMPAT_SHARE_H
#define HBMPC_COMPAT_SHARE_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ShareErrorCode {
    ShareSuccess, InsufficientShares, DegreeMismatch, IdMismatch, InvalidInput, TypeMismatch, NoSuitableDomain,
    PolynomialOperationError, DecodingError,
} ShareErrorCode;
typedef enum FieldKind { Bls12_381Fr } FieldKind;

typedef struct U256 { uint64_t data[4]; } U256;                       /* canonical value, little-endian limbs */
typedef struct U256Slice { struct U256 *pointer; uintptr_t len; } U256Slice;
typedef struct ByteSlice { uint8_t *pointer; uintptr_t len; } ByteSlice;
typedef struct UsizeSlice { uintptr_t *pointer; uintptr_t len; } UsizeSlice;
typedef struct FieldOpaque FieldOpaque;                               /* boxed field element owned by the library */
typedef struct ShamirShare { struct FieldOpaque *share; uintptr_t id; uintptr_t degree; } ShamirShare;
typedef struct RobustShare { struct FieldOpaque *share; uintptr_t id; uintptr_t degree; } RobustShare;
typedef struct NonRobustShare { struct FieldOpaque *share; uintptr_t id; uintptr_t degree; } NonRobustShare;
typedef struct ShamirShareSlice { struct ShamirShare *pointer; uintptr_t len; } ShamirShareSlice;
typedef struct RobustShareSlice { struct RobustShare *pointer; uintptr_t len; } RobustShareSlice;
typedef struct NonRobustShareSlice { struct NonRobustShare *pointer; uintptr_t len; } NonRobustShareSlice;

/* c_bindings/mod.rs:139-200 */
void free_u256_slice(struct U256Slice slice);
void free_bytes_slice(struct ByteSlice slice);
struct U256 be_bytes_to_u256(struct ByteSlice bytes);
struct U256 le_bytes_to_u256(struct ByteSlice bytes);
struct ByteSlice u256_to_be_bytes(struct U256 num);
struct ByteSlice u256_to_le_bytes(struct U256 num);
/* share/mod.rs:55-196 */
struct ByteSlice field_ptr_to_bytes(struct FieldOpaque *field, bool be);
void free_shamir_share(struct ShamirShare share);
void free_robust_share(struct RobustShare share);
void free_non_robust_share(struct NonRobustShare share);
void free_shamir_share_slice(struct ShamirShareSlice slice);
void free_robust_share_slice(struct RobustShareSlice slice);
void free_non_robust_share_slice(struct NonRobustShareSlice slice);
/* share/mod.rs:286-384: Shamir sharing on the points x = id (common/share/shamir.rs:38-126) */
struct ShamirShare shamir_share_new(struct U256 secret, uintptr_t id, uintptr_t degree, enum FieldKind field_kind);
enum ShareErrorCode shamir_share_compute_shares(struct U256 secret, uintptr_t degree, const struct UsizeSlice *ids, enum FieldKind field_kind,
                                                struct ShamirShareSlice *output_shares);
enum ShareErrorCode shamir_share_recover_secret(struct ShamirShareSlice shares, struct U256 *output_secret, struct U256Slice *output_coeffs,
                                                enum FieldKind field_kind);
/* share/mod.rs:386-501: RobustShare (robust_interpolate.rs:52-157) */
struct RobustShare robust_share_new(struct U256 secret, uintptr_t id, uintptr_t degree, enum FieldKind field_kind);
enum ShareErrorCode robust_share_compute_shares(struct U256 secret, uintptr_t degree, uintptr_t n, struct RobustShareSlice *output_shares,
                                                enum FieldKind field_kind);
enum ShareErrorCode robust_share_recover_secret(struct RobustShareSlice shares, uintptr_t n, uintptr_t t, struct U256 *output_secret,
                                                struct U256Slice *output_coeffs, enum FieldKind field_kind);
/* share/mod.rs:503-597: NonRobustShare (common/share/shamir.rs:158-239) */
struct NonRobustShare non_robust_share_new(struct U256 secret, uintptr_t id, uintptr_t degree, enum FieldKind field_kind);
enum ShareErrorCode non_robust_share_compute_shares(struct U256 secret, uintptr_t degree, uintptr_t n, struct NonRobustShareSlice *output_shares,
                                                    enum FieldKind field_kind);
enum ShareErrorCode non_robust_share_recover_secret(struct NonRobustShareSlice shares, uintptr_t n, struct U256 *output_secret,
                                                    struct U256Slice *output_coeffs, enum FieldKind field_kind);

/* not in the reference: the hbmpc_b200.h status of the last wrapper call of this thread (e.g. HBMPC_NO_DEVICE = 100) */
int hbmpc_compat_last_status(void);

#ifdef __cplusplus
}
#endif
#endif /* HBMPC_COMPAT_SHARE_H */
