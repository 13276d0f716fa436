/* bgls_b200 -- C ABI of the B200 aggregate-signature verification engine.
 *
 * Drop-in boundary for the hot path of Project-Arda/bgls.  Every entry point names the
 * reference interface it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - curve ids: BGLS_ALTBN128 (reference singleton `Altbn128`, curves/altbn128.go:32) and
 *     BGLS_BLS12_381 (`Bls12`, curves/bls12_381.go:31).
 *   - points cross the boundary as the reference's *uncompressed affine big-endian* records,
 *     i.e. exactly what Point.MarshalUncompressed() emits (curves/altbn128.go:91-93,223-225;
 *     curves/bls12_381.go:61-63,122-124):
 *         G1 = x || y                      (2 * F bytes)      F = 32 (altbn128), 48 (bls12-381)
 *         G2 = x_im || x_re || y_im || y_re (4 * F bytes)
 *         GT = 12 * F bytes, coefficients of w^5,w^3,w^1,w^4,w^2,w^0, each (im, re)
 *              (the cloudflare bn256 gfP12 marshal order, curves/altbn128.go:378-380)
 *     The point at infinity is the all-zero record (curves/README.md:19); for bls12-381 a record
 *     whose first byte has bit 0x40 set is also accepted as infinity.  Outputs always use zeros.
 *   - the arithmetic entry points assume valid points: reduced (< p) coordinates on the curve (and, where the
 *     reference enforces it, in the order-r subgroup).  The reference rejects anything else when the Point is
 *     constructed (curves/altbn128.go:42-57,160-179; curves/bls12_381.go:197-226,242-264), i.e. before this boundary;
 *     a binding does the same with bgls_validate_points (uncompressed records) / bgls_decompress_points (compressed)
 *     when it builds a Point from untrusted coordinates or bytes.  Verdicts on invalid points are undefined.
 *   - every function returns BGLS_OK (0) or a negative error; the Go side maps non-zero to
 *     `ok == false` / `nil` (curves/curve.go:126-128,137-139).  bgls_last_error() gives the text.
 *   - host-buffer entry points copy in, run on the context's stream and synchronise before
 *     returning.  `_dev` entry points take device pointers (16-byte aligned) and a cudaStream_t
 *     (as void*), enqueue the work and return without synchronising.
 *   - a context is bound to one CUDA device and is safe to share between threads / goroutines (the
 *     reference calls Pair / PairingProduct from many goroutines, curves/curve.go:132-134).  It owns
 *     64 execution slots (stream + device scratch each): concurrent host-buffer calls run on different
 *     slots and overlap on the GPU -- the single-warp final exponentiation of one product beside the
 *     Miller loops of the next.  `_dev` calls are keyed by the caller's stream: work enqueued on
 *     different streams uses different scratch; more than 64 streams in flight are serialised.
 *     The first call of a given size on a slot allocates its scratch and SYNCHRONISES the device (also for `_dev`
 *     calls; not legal inside a stream capture: warm the context up with a call of the largest size first).
 *     bgls_last_error() returns the text of the most recent failing call on the context.
 */
#ifndef BGLS_B200_H
#define BGLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGLS_ALTBN128 0
#define BGLS_BLS12_381 1

#define BGLS_G1 1
#define BGLS_G2 2

#define BGLS_OK 0
#define BGLS_ERR_ARG (-1)   /* bad curve / group / null pointer */
#define BGLS_ERR_CUDA (-2)  /* CUDA runtime error, see bgls_last_error */
#define BGLS_ERR_NODEV (-3) /* no usable CUDA device: the engine has no CPU fallback */

typedef struct bgls_ctx bgls_ctx;

int bgls_ctx_create(int device, bgls_ctx** out);
void bgls_ctx_destroy(bgls_ctx* ctx);
const char* bgls_last_error(const bgls_ctx* ctx);
/* library / kernel build identification, e.g. "bgls_b200 r1 sm_100a" */
const char* bgls_version(void);

/* CurveSystem.PairingProduct -- curves/curve.go:48,125-170; altbn128.go:143-145; bls12_381.go:238-240.
 * out_gt = prod_i e(g1[i], g2[i]).  n == 0 yields the GT identity.  is_identity (may be NULL)
 * receives PointT.Equals(GetGTIdentity()) (bgls/bgls.go:115-116). */
int bgls_pairing_product(bgls_ctx* ctx, int curve, const uint8_t* g1, const uint8_t* g2, size_t n,
                         uint8_t* out_gt, int* is_identity);

/* CurveSystem.Pair -- curves/curve.go:46; altbn128.go:130-141; bls12_381.go:228-236. */
int bgls_pair(bgls_ctx* ctx, int curve, const uint8_t* g1, const uint8_t* g2, uint8_t* out_gt);

/* PointT.Add (GT multiplication) -- curves/altbn128.go:264-271; bls12_381.go:160-168. */
int bgls_gt_mul(bgls_ctx* ctx, int curve, const uint8_t* a, const uint8_t* b, uint8_t* out_gt);

/* AggregatePoints -- curves/curve.go:73-110 (also AggregateSignatures / AggregateKeys,
 * bgls/bgls.go:123-131).  n >= 1; n == 0 returns BGLS_ERR_ARG (the reference never returns).  Records at infinity,
 * repeated points and cancelling points are all legal inputs (complete addition formulas); coordinates must be < q. */
int bgls_aggregate_points(bgls_ctx* ctx, int curve, int group, const uint8_t* pts, size_t n, uint8_t* out);

/* ScalePoints / Point.Mul -- curves/curve.go:190-214; altbn128.go:107-121,235-249.
 * scalars: n records of 32 bytes, big-endian, non-negative. */
int bgls_scale_points(bgls_ctx* ctx, int curve, int group, const uint8_t* pts, const uint8_t* scalars,
                      size_t n, uint8_t* out);

/* CurveSystem.HashToG1 -- curves/altbn128.go:509-513 (Keccak-256 try-and-increment, curves/hash.go:53-77) and
 * curves/bls12_381.go:349-351 (blake2b + Fouque-Tibouchi + cofactor, curves/hash.go:79-167): the pre-step of
 * verifyAggSig (bgls/bgls.go:106-111).  Message i is msgs[offsets[i] .. offsets[i+1]); out receives n uncompressed G1
 * records.  Two forms per curve, chosen by load (same bytes either way): a latency form (several lanes per message) and a
 * throughput form (altbn128: the lanes of a warp are re-dealt over the messages still open every round; bls12-381: one
 * cofactor multiplication for both halves of a message). */
int bgls_hash_to_g1(bgls_ctx* ctx, int curve, const uint8_t* msgs, const uint64_t* offsets, size_t n, uint8_t* out);

/* Signer-set sharding: product of the *Miller values* of n pairs without the
 * final exponentiation (GT layout, 12*F bytes).  Partials from several GPUs are exchanged
 * (e.g. NCCL all-gather) and finished with bgls_final_exp_product. */
int bgls_miller_product(bgls_ctx* ctx, int curve, const uint8_t* g1, const uint8_t* g2, size_t n, uint8_t* out_f);
int bgls_final_exp_product(bgls_ctx* ctx, int curve, const uint8_t* partials, size_t k, uint8_t* out_gt,
                           int* is_identity);

/* Throughput mode: nbatch independent pairing-product checks in one call.  Pairs of check b are
 * [offsets[b], offsets[b+1]) in g1 / g2 (offsets has nbatch+1 entries).  out_ok[b] = 1 iff the
 * product is the GT identity (one verifyAggSig each, bgls/bgls.go:94-119). */
int bgls_pairing_check_batch(bgls_ctx* ctx, int curve, const uint8_t* g1, const uint8_t* g2,
                             const uint64_t* offsets, size_t nbatch, uint8_t* out_ok);

/* ---- device-resident variants: pointers are device memory, work is enqueued on `stream` ---- */
/* d_out_gt: 12*F bytes; d_is_identity: one int32 (may be NULL) */
int bgls_pairing_product_dev(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, size_t n,
                             void* d_out_gt, void* d_is_identity, void* stream);
/* do_final = 0: raw Miller product (as bgls_miller_product) */
int bgls_miller_product_dev(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, size_t n,
                            void* d_out_f, void* stream);
int bgls_final_exp_product_dev(bgls_ctx* ctx, int curve, const void* d_partials, size_t k, void* d_out_gt,
                               void* d_is_identity, void* stream);
int bgls_aggregate_points_dev(bgls_ctx* ctx, int curve, int group, const void* d_pts, size_t n, void* d_out,
                              void* stream);
int bgls_scale_points_dev(bgls_ctx* ctx, int curve, int group, const void* d_pts, const void* d_scalars,
                          size_t n, void* d_out, void* stream);
int bgls_hash_to_g1_dev(bgls_ctx* ctx, int curve, const void* d_msgs, const void* d_offsets, size_t n, void* d_out,
                        void* stream);
int bgls_pairing_check_batch_dev(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2,
                                 const void* d_offsets, size_t nbatch, size_t total_pairs, void* d_out_ok,
                                 void* stream);

/* ---- measurement helpers (bench.py); not part of the reference surface ---- */
/* number of kernel launches issued through this context since creation */
uint64_t bgls_launch_count(const bgls_ctx* ctx);
/* when on, the pairing entry points bracket their kernels with CUDA events on the launching stream */
int bgls_set_profiling(bgls_ctx* ctx, int on);
/* device time of the last pairing call: Miller-loop kernel and finishing (product + final exp) kernel */
int bgls_last_kernel_ms(bgls_ctx* ctx, float* ms_main, float* ms_finish);
/* measured full-rate IMAD.WIDE.U32 throughput of this GPU (32x32+64 multiply-accumulates / s) */
int bgls_intpipe_peak(bgls_ctx* ctx, double* wide_mac_per_s);

/* Peer-memory exchange of the per-GPU Miller products (multi-GPU sharding of one product; nothing in the
 * reference corresponds to it).  One process per GPU: every rank creates a mailbox, the 64-byte CUDA IPC handles are
 * exchanged out of band (e.g. one all-gather at start-up), every rank connects to every peer.  Then, per product and
 * per `lane` (an independent sequence of products; `epoch` = 1, 2, 3, ... within a lane):
 *   bgls_miller_product_exchange_dev  Miller loops + product of this rank's pairs, then ONE kernel stores the 12F-byte
 *                                     partial into the mailbox of every rank over NVLink peer memory and raises the flags;
 *   bgls_final_exp_exchanged_dev      waits (on the device, bounded) until all `world` partials of (lane, epoch) have
 *                                     arrived, multiplies them and runs the final exponentiation -- on every rank.
 * No host synchronisation and no collective library call is involved.  bgls_exchange_error reports a wait that timed
 * out (a peer that never sent).  world <= 8. */
int bgls_exchange_create(bgls_ctx* ctx, int world, int rank, int lanes, uint8_t* handle_out /* 64 bytes */);
int bgls_exchange_connect(bgls_ctx* ctx, int peer_rank, const uint8_t* handle /* 64 bytes */);
int bgls_exchange_error(bgls_ctx* ctx, int* err);
int bgls_miller_product_exchange_dev(bgls_ctx* ctx, int curve, const void* d_g1, const void* d_g2, size_t n, int lane,
                                     uint64_t epoch, void* stream);
int bgls_final_exp_exchanged_dev(bgls_ctx* ctx, int curve, int lane, uint64_t epoch, void* d_out_gt, void* d_is_identity,
                                 void* stream);

/* verifyAggSig -- bgls/bgls.go:94-119, the body of VerifyAggregateSignature (bgls.go:82-84, allow_duplicates = 0) and
 * KoskVerifyAggregateSignature / DistinctMsgVerifyAggregateSignature (allow_duplicates = 1; the caller prepends the
 * 0x01 byte / the public key to each message as bgls/blsKosk.go:100-106 and blsDistinctMessage.go:45-57 do).
 * Message i is msgs[offsets[i] .. offsets[i+1]); keys: n uncompressed G2 records; sig: one uncompressed G1 record.
 * One call does what the reference does in n goroutines plus PairingProduct: the duplicate-message check on the host
 * (bgls.go:139-150: a duplicate makes the verdict false before any pairing), HashToG1 of every message, sigma -> -sigma,
 * the (n+1)-pair product against [keys..., g2] and the comparison with the GT identity.  *ok = 1 iff it verifies. */
int bgls_verify_aggregate_signature(bgls_ctx* ctx, int curve, const uint8_t* msgs, const uint64_t* offsets, size_t n,
                                    const uint8_t* keys, const uint8_t* sig, int allow_duplicates, int* ok);

/* verifyMultiSignature -- bgls/bgls.go:89-92 (exported as KoskVerifyMultiSignature, bgls/blsKosk.go:117-120, whose caller
 * prepends the 0x01 byte to the message): vs = AggregatePoints(keys) over n >= 1 uncompressed G2 records, then the
 * single-signature check of bgls.go:65-70 on HashToG1(msg).  *ok = 1 iff it verifies. */
int bgls_verify_multi_signature(bgls_ctx* ctx, int curve, const uint8_t* msg, size_t msg_len, const uint8_t* keys, size_t n,
                                const uint8_t* sig, int* ok);

/* Point.Marshal (compressed form) -- curves/altbn128.go:81-89 (G1), :203-221 (G2); curves/bls12_381.go:57-59,118-120.
 * pts: n uncompressed records; out: n compressed records of F (G1) / 2F (G2) bytes.
 *   altbn128: x, bit 7 of byte 0 set iff 2y > q; G2 = x_im || x_re with one sign bit per y component.
 *   bls12-381: the zcash serialisation of the upstream dis2/bls12 library (0x80 compressed, 0x40 infinity,
 *   0x20 larger root) -- restated from the format's definition, the upstream source is not in the reference tree. */
int bgls_compress_points(bgls_ctx* ctx, int curve, int group, const uint8_t* pts, size_t n, uint8_t* out);
int bgls_compress_points_dev(bgls_ctx* ctx, int curve, int group, const void* d_pts, size_t n, void* d_out, void* stream);

/* Validation of uncompressed records -- what CurveSystem.MakeG1Point / MakeG2Point / UnmarshalG1 / UnmarshalG2 enforce
 * before a Point exists (curves/altbn128.go:42-57,160-179; curves/bls12_381.go:197-226,242-264).
 * out_ok[i] = 1 when record i is the point at infinity or has coordinates < q that satisfy the curve equation and
 *   BGLS_VALIDATE_ONCURVE    nothing more (bls12-381 MakeG*Point with check = false does not even do this)
 *   BGLS_VALIDATE_REFERENCE  what the reference's decoding enforces: additionally r*P = infinity for altbn128 G2 (the
 *                            upstream bn256 twist check) and for bls12-381 G1 / G2 (Check()); altbn128 G1 has cofactor 1. */
#define BGLS_VALIDATE_ONCURVE 0
#define BGLS_VALIDATE_REFERENCE 1
int bgls_validate_points(bgls_ctx* ctx, int curve, int group, const uint8_t* pts, size_t n, int mode, uint8_t* out_ok);
int bgls_validate_points_dev(bgls_ctx* ctx, int curve, int group, const void* d_pts, size_t n, int mode, void* d_out_ok, void* stream);

/* PointT.Mul -- curves/curve.go:63-70; curves/altbn128.go:290-294; curves/bls12_381.go:186-195.
 * out_gt = a^e for a GT element a; exponent32 = |e| as 32 big-endian bytes, negative != 0 for e < 0 (a^-1 is the conjugate). */
int bgls_gt_pow(bgls_ctx* ctx, int curve, const uint8_t* a, const uint8_t* exponent32, int negative, uint8_t* out_gt);

/* CurveSystem.UnmarshalG1 / UnmarshalG2 on compressed input -- curves/altbn128.go:296-376 (square roots
 * curves/hash.go:178-223); curves/bls12_381.go:242-264.  out_pts: n uncompressed records (zeros where rejected);
 * out_ok[i] = 1 when record i decoded to a curve point (or infinity), 0 when the reference would return
 * (nil, false): malformed flags, coordinate >= q, or x not on the curve.  check_subgroup != 0 additionally
 * requires r*P = infinity (the reference's bls12 path calls Check(); its altbn128 path does not). */
int bgls_decompress_points(bgls_ctx* ctx, int curve, int group, const uint8_t* in, size_t n, int check_subgroup,
                           uint8_t* out_pts, uint8_t* out_ok);
int bgls_decompress_points_dev(bgls_ctx* ctx, int curve, int group, const void* d_in, size_t n, int check_subgroup,
                               void* d_out_pts, void* d_out_ok, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BGLS_B200_H */
