/*
 * hbmpc_b200.h -- C ABI of the B200-native finite-field hot path of HoneyBadgerMPC secret sharing.
 *
 * Drop-in boundary for the reference's share arithmetic (Stoffel-Labs/mpc-protocols, paths relative to
 * /root/reference/mpc/src).  Each entry point names the reference interface it replaces; INTEGRATION.md shows the
 * Rust `extern "C"` binding a maintainer would add on the reference side.
 *
 * Conventions (they mirror the reference's own C ABI, ffi/c_bindings/mod.rs:17-49 and share/mod.rs:18-37):
 *   - a field element is the CANONICAL value of ark_bls12_381::Fr as 4 x uint64_t little-endian limbs (== `U256`);
 *     arrays are dense, row-major; inputs must be < r (else HBMPC_INVALID_INPUT, like `from_bigint(..).unwrap()`);
 *   - return value = `ShareErrorCode` numbering (0 success ... 8 DecodingError) for whole-call failures;
 *     per-item outcomes are reported in `path[]`;
 *   - every data pointer may be a host pointer or a device pointer (detected with cudaPointerGetAttributes).
 *     Calls with a host buffer are pipelined in chunks over internal streams (host->device copy, kernels and
 *     device->host copy of neighbouring chunks overlap; use pinned memory for full PCIe speed) and return after the
 *     results are in the caller's buffers; calls with device buffers only run on the context's stream.  That stream is
 *     created non-blocking: device buffers written on another stream must be complete (or ordered by an event, or the
 *     context moved onto that stream with hbmpc_ctx_set_stream) before the call -- the library does not synchronize with
 *     streams it does not own;
 *   - outputs are CALLER-allocated (the reference leaks Vecs to C and frees them through free_* helpers);
 *   - no hidden RNG: share generation takes the polynomial coefficients, or a 32-byte StdRng seed from which the device draws them
 *     exactly as the reference's generator would (hbmpc_share_secrets_batch), so results are reproducible;
 *   - there is no CPU fallback: every call fails with HBMPC_NO_DEVICE if no CUDA device is usable.
 * A context is thread-compatible: one caller at a time per context; use one context per GPU / per thread.
 */
#ifndef HBMPC_B200_H
#define HBMPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ShareErrorCode (ffi/c_bindings/share/mod.rs:18-37) + library-level codes >= 100 */
enum {
    HBMPC_SUCCESS = 0,
    HBMPC_INSUFFICIENT_SHARES = 1,
    HBMPC_DEGREE_MISMATCH = 2,
    HBMPC_ID_MISMATCH = 3,
    HBMPC_INVALID_INPUT = 4,
    HBMPC_TYPE_MISMATCH = 5,
    HBMPC_NO_SUITABLE_DOMAIN = 6,
    HBMPC_POLYNOMIAL_OPERATION_ERROR = 7,
    HBMPC_DECODING_ERROR = 8,
    HBMPC_NO_DEVICE = 100,  /* no usable CUDA device / driver: the library never computes on the CPU */
    HBMPC_CUDA_ERROR = 101  /* a CUDA runtime call failed; see hbmpc_last_error() */
};

typedef struct hbmpc_ctx hbmpc_ctx;

/* One context per GPU: owns the stream, the cached constant tables (domain, Vandermonde, Lagrange, syndrome
 * matrices keyed by (n, d, t, id-set); reference analogue: the OnceLock caches at common/mod.rs:43,70) and scratch. */
int hbmpc_ctx_create(int device, hbmpc_ctx **out);
void hbmpc_ctx_destroy(hbmpc_ctx *ctx);
/* Run on an existing stream (e.g. the caller's torch / application stream) instead of the context's own. */
int hbmpc_ctx_set_stream(hbmpc_ctx *ctx, void *cuda_stream);
/* Synchronous mode (default): every call waits for its kernels and returns the device-side status (non-canonical
 * input -> HBMPC_INVALID_INPUT, undecodable item -> HBMPC_DECODING_ERROR).  Asynchronous mode: calls whose buffers are
 * all device pointers only enqueue work on the stream and return 0; the accumulated status is returned (and cleared)
 * by hbmpc_ctx_synchronize. */
int hbmpc_ctx_set_async(hbmpc_ctx *ctx, int async);
int hbmpc_ctx_synchronize(hbmpc_ctx *ctx);
/* Number of this library's kernels launched on the context so far. */
uint64_t hbmpc_ctx_launch_count(const hbmpc_ctx *ctx);
const char *hbmpc_last_error(const hbmpc_ctx *ctx);

/* K1.  Replaces the body of RobustShare::compute_shares (honeybadger/robust_interpolate/robust_interpolate.rs:52-82)
 * and NonRobustShare::compute_shares (common/share/shamir.rs:158-196) for a batch of B secrets:
 *   coeffs[B][d+1] (coeffs[b][0] = secret b, the rest = the random polynomial)  ->  shares[B][n],
 *   shares[b][j] = P_b(w_N^j), N = next_pow2(n)  (share id j, degree d).
 * Errors: n <= d -> INVALID_INPUT (:59-64); n > 256 -> NO_SUITABLE_DOMAIN (:65-66; the reference caps n at 255,
 * honeybadger/mod.rs:441-444). */
int hbmpc_compute_shares_batch(hbmpc_ctx *ctx, size_t n, size_t d, size_t B, const uint64_t *coeffs, uint64_t *shares);

/* K2.  Replaces make_vandermonde + apply_vandermonde (common/share/mod.rs:31-76) looped over B chunks
 * (batch_recon.rs:157-165, ran_dou_sha/mod.rs:392-403, share_gen.rs:415-419):
 *   out[b][j] = sum_k w_N^(j*k) * in[b][k],  j < n, k < cols.   recipient_major != 0 writes out[j][b]
 * (the per-recipient transposition of batch_recon.rs:158-165). */
int hbmpc_apply_vandermonde_batch(hbmpc_ctx *ctx, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *out,
                                  int recipient_major);

/* K2'.  Same with a caller-supplied rows x cols matrix (canonical limbs, host pointer). */
int hbmpc_apply_matrix_batch(hbmpc_ctx *ctx, size_t rows, size_t cols, const uint64_t *matrix, size_t B,
                             const uint64_t *in, uint64_t *out, int recipient_major);

/* K3 (+K4).  Replaces batch_recover_secret (robust_interpolate.rs:284-443):
 *   evals[S][B] sender-major with sender_ids[S] (arrival order), B chunks  ->  coeffs[B][d+1] (always d+1 wide,
 *   zero padded), path[B]: 0 = optimistic path, r > 0 = accepted in OEC round r, < 0 = -(ShareErrorCode) for that chunk,
 *   flags[B][ceil(S/64)] (optional, may be NULL): bit i set <=> supplied share i (arrival order) disagrees with the
 *   decoded polynomial.  Chunks failing the optimistic check are decoded by the robust path (K4), like :429-439.
 * Return: whole-call validation errors (:290-341); else, like the reference's `?` at :437, the code of the first
 * chunk (lowest index) whose robust decode failed, 0 if none.  All chunks are decoded either way. */
int hbmpc_batch_recover(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B,
                        const uint64_t *evals, uint64_t *coeffs, int32_t *path, uint64_t *flags);

/* Same decode, but only the opened values P_b(0) are produced: secrets[B] (what round 1 of batch reconstruction
 * consumes, batch_recon.rs:384-391). */
int hbmpc_batch_recover_secrets(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids,
                                size_t B, const uint64_t *evals, uint64_t *secrets, int32_t *path);

/* K4 direct.  Replaces RobustShare::recover_secret (robust_interpolate.rs:94-157: optimistic -> OEC -> Gao) for a
 * batch of B codewords with a common id set: shares[B][S] codeword-major, ids[S]  ->  coeffs[B][d+1], secrets[B]
 * (may be NULL), path[B], flags (may be NULL).  Same return convention as hbmpc_batch_recover. */
int hbmpc_robust_interpolate_batch(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *ids, size_t B,
                                   const uint64_t *shares, uint64_t *coeffs, uint64_t *secrets, int32_t *path,
                                   uint64_t *flags);

/* a10.  Replaces NonRobustShare::recover_secret (common/share/shamir.rs:199-239: Lagrange interpolation through ALL
 * supplied points + degree check) for a batch with a common id set; it is the consistency check of the RanDouSha
 * checkers (ran_dou_sha/mod.rs:568-602) and DoubleShare.  shares[B][S] (sender_major == 0) or shares[S][B]
 * (sender_major != 0)  ->  coeffs[B][deg+1] (zero padded), secrets[B] = coefficient 0 (may be NULL),
 * status[B] = degree of the interpolant (>= 0; the caller compares it with t / 2t like ran_dou_sha/mod.rs:589-595) or
 * -HBMPC_DEGREE_MISMATCH when the interpolant exceeds `deg` (outputs of that item are zeroed).
 * Whole-call errors: empty / duplicate ids / id >= n -> INVALID_INPUT, S < deg+1 -> INSUFFICIENT_SHARES. */
int hbmpc_nonrobust_recover_batch(hbmpc_ctx *ctx, size_t n, size_t deg, size_t S, const size_t *ids, size_t B,
                                  const uint64_t *shares, int sender_major, uint64_t *coeffs, uint64_t *secrets, int32_t *status);

/* K5.  Element-wise share algebra (common/mod.rs:167-300; triple_generation.rs:332-340,196-208;
 * multiplication.rs:79-97,417-426): out[i] = a[i] (op) b[i], op: 0 add, 1 sub, 2 mul (share_mul / Mul<F>). */
int hbmpc_elementwise(hbmpc_ctx *ctx, int op, size_t count, const uint64_t *a, const uint64_t *b, uint64_t *out);

/* K5 fused.  The share algebra of one protocol step in ONE pass over HBM; every array holds `count` canonical
 * values (host or device pointers, like every other call); results are bit-identical to the operator-by-operator route above.
 *   HBMPC_K5_TRIPLE_MASK      in = {a, b, r_2t}                  out = {a*b - r_2t}                triple_generation.rs:332-340
 *   HBMPC_K5_BEAVER_MASK      in = {a, x, b, y}                  out = {a - x, b - y}              multiplication.rs:417-426
 *   HBMPC_K5_BEAVER_FINALIZE  in = {c, x, y, a-x (open), b-y (open)}
 *                             out = {c - (a-x)(b-y) - (a-x)*y - (b-y)*x}                           multiplication.rs:79-97
 * An output may alias an input.  A non-canonical input value -> HBMPC_INVALID_INPUT. */
enum { HBMPC_K5_TRIPLE_MASK = 0, HBMPC_K5_BEAVER_MASK = 1, HBMPC_K5_BEAVER_FINALIZE = 2 };
int hbmpc_share_algebra_fused(hbmpc_ctx *ctx, int op, size_t count, const uint64_t *const *in, uint64_t *const *out);

/* N1 (wire format).  ark-serialize writes a ShamirShare<F,1,_> record as 32-byte LE canonical value + u64 id + u64 degree
 * (48 bytes; Vec<RobustShare> payloads of share_gen.rs:255-268, ran_dou_sha/messages.rs:40-46).  These helpers split such
 * records into the value array the kernels consume (ids / degrees optional, may be NULL) and build records from values
 * without a host-side repacking pass; `records` may be only 8-byte aligned (payload + 8 after the Vec length prefix).
 * pack: record i gets id = i / per_id and the given degree. */
int hbmpc_unpack_share_records(hbmpc_ctx *ctx, size_t count, const void *records, uint64_t *values, uint64_t *ids, uint64_t *degrees);
int hbmpc_pack_share_records(hbmpc_ctx *ctx, size_t count, const uint64_t *values, size_t per_id, size_t degree, void *records);

/* N1 (message payloads).  The per-sender vectors of batch reconstruction arrive, and the per-recipient vectors leave, as SEPARATE
 * message payloads (`Vec<F>::serialize_compressed`: u64 length + 32-byte LE canonical values; batch_recon.rs:174-175, 339, 419;
 * common/utils.rs:3-21).  These variants take one HOST pointer per sender / recipient -- `payload + 8`, 8-byte aligned -- and move the
 * bytes straight between the message buffers and the device: no host-side gather into a contiguous [S][B] array, no scatter out of
 * [n][B], (de)serialisation and transposition are the copy pattern of the call.  Same results, errors and chunked pipeline as
 * hbmpc_batch_recover / hbmpc_batch_recover_secrets / hbmpc_apply_vandermonde_batch(recipient_major = 1). */
int hbmpc_batch_recover_msgs(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B,
                             const uint64_t *const *sender_evals, uint64_t *coeffs, int32_t *path, uint64_t *flags);
int hbmpc_batch_recover_secrets_msgs(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B,
                                     const uint64_t *const *sender_evals, uint64_t *secrets, int32_t *path);
int hbmpc_apply_vandermonde_msgs(hbmpc_ctx *ctx, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *const *recipient_out);

/* N4.  The randomness of a sharing as the reference draws it, on the device.  The reference samples the polynomial inside
 * compute_shares (robust_interpolate.rs:68-69: DensePolynomial::rand(d, rng) with coefficient 0 overwritten by the secret; callers draw
 * the secret with F::rand first: share_gen.rs:250, double_share_generation.rs:167) from rand 0.8 `StdRng` (ChaCha12, 64-bit block
 * counter from 0, stream 0) with ark-ff 0.5 `Fp::rand` (four next_u64 limbs, top bit cleared, redrawn while >= r, accepted limbs = the
 * Montgomery representation).  seed32 = the 32-byte StdRng::from_seed seed.
 *   hbmpc_sample_fr_batch:     out[count] = the first `count` elements `F::rand(&mut rng)` returns (canonical limbs).
 *   hbmpc_sample_polynomials:  coeffs[B][d+1] of B consecutive sharings drawn from ONE generator.  secrets == NULL: every sharing draws
 *       its secret and then d+1 coefficients of which the first is dropped (d+2 draws: the RanSha / DouSha dealers);  secrets != NULL
 *       (secrets[B], host or device): d+1 draws per sharing, the first dropped, coeffs[b][0] = secrets[b] (the C ABI path,
 *       ffi/c_bindings/share/mod.rs:418-425).  Feed the result to hbmpc_compute_shares_batch: a dealer uploads a seed, not 32*(d+1)
 *       bytes per secret.  out / coeffs may be host or device pointers; at most ~3.8e9 draws per call. */
int hbmpc_sample_fr_batch(hbmpc_ctx *ctx, const uint8_t *seed32, size_t count, uint64_t *out);
int hbmpc_sample_polynomials(hbmpc_ctx *ctx, const uint8_t *seed32, size_t B, size_t d, const uint64_t *secrets, uint64_t *coeffs);
/* K1 with the reference's own argument meaning -- RobustShare::compute_shares(secret, n, degree, None, rng)
 * (robust_interpolate.rs:52-82; C ABI: robust_share_compute_shares, ffi/c_bindings/share/mod.rs:410-445) for B secrets: the
 * polynomials are drawn on the device as B consecutive calls on ONE StdRng::from_seed(seed32) would draw them
 * (== hbmpc_sample_polynomials with secrets, then hbmpc_compute_shares_batch), so a dealer moves 32 bytes per secret to the device
 * instead of 32*(d+1).  secrets[B] -> shares[B][n]; coeffs_out[B][d+1] (may be NULL) receives the polynomials.  Host or device
 * pointers; same errors as hbmpc_compute_shares_batch. */
int hbmpc_share_secrets_batch(hbmpc_ctx *ctx, const uint8_t *seed32, size_t n, size_t d, size_t B, const uint64_t *secrets,
                              uint64_t *shares, uint64_t *coeffs_out);

/* N4 (tail).  The same path over the reference's second field, GoldilocksField = Fp64, p = 2^64 - 2^32 + 1, generator 7
 * (common/math/goldilocks.rs:4-13; the RandBit / PRandInt pipeline instantiates the generic sharing code with it).  An element is its
 * canonical value in ONE uint64_t (< p, else HBMPC_INVALID_INPUT); shapes, evaluation domain (GeneralEvaluationDomain::new(n):
 * w_N = 7^((p-1)/N)), id conventions, validation and error codes as in the Fr entry points of the same name.  hbmpc_gl_batch_recover is
 * batch_recover_secret's optimistic path (interpolate the lowest d+1 ids, check the next t: robust_interpolate.rs:343-428): a chunk
 * that fails the check gets path[b] = -HBMPC_DECODING_ERROR and zeroed outputs, and the call returns HBMPC_DECODING_ERROR -- the
 * error-correcting decoder is not instantiated for this field (hand those chunks to the reference's CPU decoder).  coeffs or secrets
 * (not both) may be NULL.  Matrices are limited to rows*cols <= 25600 (n = 64 .. 128 shapes fit).  8-byte elements make these calls
 * HBM / PCIe bound. */
int hbmpc_gl_compute_shares_batch(hbmpc_ctx *ctx, size_t n, size_t d, size_t B, const uint64_t *coeffs, uint64_t *shares);
int hbmpc_gl_apply_vandermonde_batch(hbmpc_ctx *ctx, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *out, int recipient_major);
int hbmpc_gl_batch_recover(hbmpc_ctx *ctx, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B, const uint64_t *evals,
                           uint64_t *coeffs, uint64_t *secrets, int32_t *path);
int hbmpc_gl_nonrobust_recover_batch(hbmpc_ctx *ctx, size_t n, size_t deg, size_t S, const size_t *ids, size_t B, const uint64_t *shares,
                                     int sender_major, uint64_t *coeffs, uint64_t *secrets, int32_t *status);
int hbmpc_gl_elementwise(hbmpc_ctx *ctx, int op, size_t count, const uint64_t *a, const uint64_t *b, uint64_t *out);
/* device ordinal of a context */
int hbmpc_ctx_device(const hbmpc_ctx *ctx);

/* Single-process multi-GPU.  The reference party is one process that issues all sessions' work before awaiting
 * (honeybadger/mod.rs:245-257,1362-1375); every call of this path is a map over independent secrets / chunks / codewords, so a group
 * (one context per device, tables replicated) splits the batch into contiguous ranges [g*B/G, (g+1)*B/G), one internal host thread
 * per device, with NO collective on the data path.  Group calls take HOST buffers (pinned for full PCIe speed) laid out exactly as in
 * the single-context calls and return when every range is complete; return value = the first non-zero status in device order.
 * hbmpc_group_ctx(i) / hbmpc_group_shard_range give the member contexts and ranges to callers that keep their batches resident on
 * the devices themselves (member contexts are ordinary contexts: one caller per context at a time). */
typedef struct hbmpc_group hbmpc_group;
int hbmpc_group_create(const int *devices, size_t n_devices, hbmpc_group **out);
void hbmpc_group_destroy(hbmpc_group *grp);
size_t hbmpc_group_size(const hbmpc_group *grp);
hbmpc_ctx *hbmpc_group_ctx(hbmpc_group *grp, size_t i);
void hbmpc_group_shard_range(const hbmpc_group *grp, size_t B, size_t i, size_t *lo, size_t *hi);
int hbmpc_group_compute_shares_batch(hbmpc_group *grp, size_t n, size_t d, size_t B, const uint64_t *coeffs, uint64_t *shares);
int hbmpc_group_apply_vandermonde_batch(hbmpc_group *grp, size_t n, size_t cols, size_t B, const uint64_t *in, uint64_t *out,
                                        int recipient_major);
int hbmpc_group_batch_recover(hbmpc_group *grp, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B,
                              const uint64_t *evals, uint64_t *coeffs, int32_t *path, uint64_t *flags);
int hbmpc_group_batch_recover_secrets(hbmpc_group *grp, size_t n, size_t d, size_t t, size_t S, const size_t *sender_ids, size_t B,
                                      const uint64_t *evals, uint64_t *secrets, int32_t *path);
int hbmpc_group_robust_interpolate_batch(hbmpc_group *grp, size_t n, size_t d, size_t t, size_t S, const size_t *ids, size_t B,
                                         const uint64_t *shares, uint64_t *coeffs, uint64_t *secrets, int32_t *path, uint64_t *flags);

/* Integer-pipe roofline probe: runs a register-only dependent-chain microkernel on every SM and returns the sustained
 * rate in 1e9 thread-level instructions per second.  variant: 0 = mad.lo.u32 (IMAD, the north star's "IMAD peak"),
 * 1 = IMAD.WIDE.U32(.X) carry chains (the 32x32->64 multiply-add the product kernels issue), 2 = DFMA (FP64 pipe). */
int hbmpc_measure_imad_peak(hbmpc_ctx *ctx, int variant, double *giga_inst_per_s, double *elapsed_ms);

/* Latency probe behind the kernels' occupancy choices: `chains` (1, 2, 4, 8) independent serial IMAD.WIDE.U32.X carry chains per
 * thread at `warps_per_smsp` (1 .. 16) resident warps per SM sub-partition; returns 1e9 thread-level multiply-adds per second. */
int hbmpc_measure_wide_chains(hbmpc_ctx *ctx, int chains, int warps_per_smsp, double *giga_inst_per_s);
/* Throughput of the Montgomery product the transforms and the decoder are built on (fr.cuh: mont_mul / mont_mul2), register-only:
 * ilp (1, 2, 4) independent product chains per thread at warps_per_smsp resident warps; returns 1e9 products per second. */
int hbmpc_measure_mont_mul(hbmpc_ctx *ctx, int ilp, int warps_per_smsp, double *giga_products_per_s);

#ifdef __cplusplus
}
#endif
#endif /* HBMPC_B200_H */
