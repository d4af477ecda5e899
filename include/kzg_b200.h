/* kzg_b200.h — C ABI of the B200-native KZG engine (libkzg_b200.so).
 *
 * Drop-in boundary for the hot path of VGLoic/kzg-poly-commit-exploration:
 * trusted-setup generation -> commit -> evaluate -> evaluation proof over BLS12-381.
 * One FFI crossing per BULK operation (the reference crosses into blst once per group /
 * field operation).  Citations are into the reference tree (/root/reference).
 *
 * Conventions
 *   - every function returns a kzg_status (0 = OK) and never throws or aborts;
 *     kzg_last_error(ctx) returns a message for the last failure on that ctx.
 *   - canonical encodings only: Fr = 32 bytes little-endian (inputs >= r are reduced mod r,
 *     as Scalar::from_le_bytes does, src/scalar.rs:54-61); G1 = 48-byte compressed ZCash
 *     format (Serialize for G1Point, src/curves.rs:99-110).
 *   - pointers are HOST pointers unless the parameter name ends in _dev (device pointers in
 *     the ctx's CUDA device, e.g. a torch tensor's data_ptr()).
 *   - a kzg_ctx is externally synchronised (one caller at a time) and is either bound to ONE CUDA
 *     device (kzg_ctx_create) or drives SEVERAL devices from one host process
 *     (kzg_ctx_create_multi): the same calls then shard internally — one sum by point range
 *     with the partial sums exchanged over NVLink peer memory, batches by row, setup generation
 *     by index range — which is what a single-process host like the reference
 *     (src/main.rs:170) needs to reach every GPU of the node.  The one-process-per-GPU mode
 *     (kzg_p2p_* + kzg_msm_g1_sharded_*, or NCCL all-gather of kzg_g1_partial blobs folded with
 *     kzg_g1_combine_dev) serves torchrun / MPI style hosts.
 *   - setups must be destroyed before the context they were created on.
 *   - there is NO CPU fallback for bulk operations: without a CUDA device kzg_ctx_create
 *     fails with KZG_E_CUDA.  The kzg_fr_* / kzg_g1_* single-element helpers are host code
 *     (they replace per-element blst calls of the public Rust API, not the hot path).
 */
#ifndef KZG_B200_H
#define KZG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kzg_ctx kzg_ctx;     /* device, streams, scratch arenas                        */
typedef struct kzg_setup kzg_setup; /* device-resident G1 powers of tau (+ precomputed tables) */

typedef enum {
  KZG_OK = 0,
  KZG_E_DEGREE_TOO_HIGH = 1,   /* Polynomial::commit, src/polynomial.rs:201-205            */
  KZG_E_CONSTANT_POLY = 2,     /* divide_by_root, src/polynomial.rs:159-166                */
  KZG_E_REMAINDER_NONZERO = 3, /* divide_by_root, src/polynomial.rs:188-192                */
  KZG_E_BAD_POINT = 4,         /* bytes_to_blst_p1, src/curves.rs:131-137                  */
  KZG_E_BAD_LENGTH = 5,        /* src/curves.rs:120-125, src/scalar.rs:249-270             */
  KZG_E_SETUP_TOO_SHORT = 6,   /* verify_proof indexes setup[1], src/polynomial.rs:284     */
  KZG_E_BAD_ARG = 7,
  KZG_E_CUDA = 100,
  KZG_E_NOMEM = 102
} kzg_status;

/* size of an exported partial G1 sum (XYZZ, Montgomery limbs) used between GPUs */
#define KZG_G1_PARTIAL_BYTES 192

/* ---- context ------------------------------------------------------------------------- */
int32_t kzg_ctx_create(kzg_ctx** out, int32_t device_id);
/* One context over n_dev devices of this process (1 <= n_dev <= 32; all pairs need peer access: NVLink / NVSwitch).
 * Every call below works on it unchanged; what changes is where the work runs:
 *   kzg_setup_generate / _import_g1 : up to 2^16 points the setup is REPLICATED on every device (the blob-batch shape),
 *        above that it is SHARDED BY POINT RANGE, device g holding items [g n/N, (g+1) n/N) (option "setup_layout":
 *        0 auto, 1 range, 2 replicated);
 *   kzg_commit / kzg_msm_g1 / _submit / kzg_open : range-sharded setup: every device sums its range, ONE kernel per
 *        device stores the 192-byte partial into every peer's mailbox, waits, folds and compresses (no NCCL, no second
 *        process); replicated setup: device 0;
 *   kzg_*_batch : rows split evenly over the devices (replicated setup), no exchange;
 *   kzg_evaluate / kzg_quotient : device 0 (a single polynomial's scan does not shard).
 * Device pointers (`_dev` parameters) may live on any device of the group.  The kzg_p2p_* / kzg_msm_g1_sharded_* /
 * kzg_msm_g1_partial_dev calls belong to the one-process-per-GPU mode and return KZG_E_BAD_ARG here.
 * Listing one device twice is allowed (two members share it): it exercises the whole path on a single-GPU box. */
int32_t kzg_ctx_create_multi(kzg_ctx** out, const int32_t* device_ids, int32_t n_dev);
int32_t kzg_ctx_device_count(const kzg_ctx* ctx);
void kzg_ctx_destroy(kzg_ctx* ctx);
const char* kzg_last_error(const kzg_ctx* ctx);
const char* kzg_status_string(int32_t status);
/* block until all work queued on the ctx stream has finished */
int32_t kzg_ctx_sync(kzg_ctx* ctx);
/* the ctx's cudaStream_t: lets a host enqueue its own work (e.g. the NCCL all-gather of the partial sums,
 * through torch.cuda.ExternalStream) in order with the engine's kernels instead of synchronising the host */
void* kzg_ctx_stream(const kzg_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t kzg_ctx_launch_count(const kzg_ctx* ctx);
/* device time of the last bulk call's dominant kernel(s) in ms, measured with CUDA events on
 * the ctx stream: which = 0 accumulation phase (affine halving rounds + bucket accumulate, or the fixed-base
 * batch), 1 whole call, 2 the single dominant kernel (round-1 pair additions, or the bucket accumulate
 * when the MSM is too small for the halving rounds) */
float kzg_ctx_last_kernel_ms(const kzg_ctx* ctx, int32_t which);
/* CUDA-event stopwatch on the ctx stream (bench.py times its K-step region with it) */
int32_t kzg_ctx_timer_start(kzg_ctx* ctx);
int32_t kzg_ctx_timer_stop(kzg_ctx* ctx, float* ms);
/* Page-locked host memory for the buffers handed to the host-pointer calls: copies from it are asynchronous and run at
 * the full PCIe rate (a pageable 32 MiB scalar vector costs several ms of staging per call).  Plain malloc'd memory works
 * everywhere, this is the fast path; a Rust / C++ host needs no CUDA headers for it. */
void* kzg_host_alloc(size_t bytes);
void kzg_host_free(void* p);
/* tuning: window bits for the bucket MSM (0 = auto) and for the fixed-base batch tables */
int32_t kzg_ctx_set_option(kzg_ctx* ctx, const char* name, int64_t value);

/* ---- trusted setup: SetupArtifactsGenerator::new(secret).take(n) ---------------------
 * src/trusted_setup.rs:20-28,40-78.  Item i = [s^i]G1 with s = secret_be mod r; item 0 is
 * the generator whatever s is.  `first` lets a rank generate only items first..first+n
 * (index-range sharding); the setup then holds those n items as its bases 0..n. */
int32_t kzg_setup_generate(kzg_ctx* ctx, const uint8_t secret_be[32], uint64_t first, uint64_t n,
                           kzg_setup** out);
/* Deserialize for G1Point x n (src/curves.rs:112-183): 48-byte compressed points, on-curve
 * check, no subgroup check; KZG_E_BAD_POINT on the first malformed entry. */
int32_t kzg_setup_import_g1(kzg_ctx* ctx, const uint8_t* g1_c48, uint64_t n, kzg_setup** out);
/* Serialize for G1Point x count (src/curves.rs:99-110) */
int32_t kzg_setup_export_g1(kzg_ctx* ctx, const kzg_setup* s, uint64_t first, uint64_t count,
                            uint8_t* out_c48);
uint64_t kzg_setup_len(const kzg_setup* s);
/* Geometry the bucket MSM uses for n scalars over this setup (for reports; needs kzg_setup_prepare_msm first):
 * plan = { window_bits c, windows W = ceil(255/c), buckets 2^(c-1), affine halving rounds, table pitch in bytes }. */
int32_t kzg_msm_plan(const kzg_ctx* ctx, const kzg_setup* s, uint64_t n, int32_t plan[5]);
/* G2 half of SetupArtifact (src/trusted_setup.rs:31-35, 64-72): [s^i]G2 for items 0..count of the setup
 * (count <= len; the reference carries one per item, verify_proof reads only item 1), kept on the device
 * as 96-byte compressed points = Serialize for G2Point (src/curves.rs:260-271). */
int32_t kzg_setup_generate_g2(kzg_ctx* ctx, kzg_setup* s, const uint8_t secret_be[32], uint64_t count);
/* Deserialize for G2Point x count (src/curves.rs:273-343): on-curve check, no subgroup check */
int32_t kzg_setup_import_g2(kzg_ctx* ctx, kzg_setup* s, const uint8_t* g2_c96, uint64_t count);
int32_t kzg_setup_export_g2(kzg_ctx* ctx, const kzg_setup* s, uint64_t first, uint64_t count,
                            uint8_t* out_c96);
uint64_t kzg_setup_g2_len(const kzg_setup* s);
void kzg_setup_destroy(kzg_setup* s);
/* Precompute per-base tables once per setup (amortised over every later commit):
 *   kzg_setup_prepare_msm  : 2^(c*j) * P_i for all windows j (folds every Pippenger window
 *                            into one bucket set; c = window_bits, 0 = auto from n)
 *   kzg_setup_prepare_fixed: d * 2^(c*j) * P_i, d = 1..2^(c-1), for the first n_points bases
 *                            (bucket-free fixed-base commit used for blob batches)           */
int32_t kzg_setup_prepare_msm(kzg_ctx* ctx, kzg_setup* s, int32_t window_bits);
int32_t kzg_setup_prepare_fixed(kzg_ctx* ctx, kzg_setup* s, uint64_t n_points, int32_t window_bits);
/* window_bits = 0 picks the widest window (<= 15) whose table fits the memory budget: option "fixed_table_max_bytes",
 * else 66 % of the device memory free at the time (blob throughput is ~ 1 / ceil(255 / c); n = 4096: c = 15 is 102 GiB,
 * c = 14 51 GiB, c = 13 26 GiB).  plan = { window_bits c, lookups per coefficient W, points covered, table bytes }. */
int32_t kzg_setup_fixed_plan(const kzg_setup* s, int64_t plan[4]);

/* ---- commit / evaluate / open ----------------------------------------------------------
 * Coefficients: n x 32-byte LE, ascending degree; trailing zero coefficients are trimmed as
 * Polynomial::try_from does (src/polynomial.rs:55-75).                                      */
/* Polynomial::commit (src/polynomial.rs:200-215) */
int32_t kzg_commit(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t n,
                   uint8_t out_c48[48]);
/* Polynomial::evaluate (src/polynomial.rs:112-123) */
int32_t kzg_evaluate(kzg_ctx* ctx, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32],
                     uint8_t y_le[32]);
/* sub + divide_by_root (src/polynomial.rs:128-195): q = (p - y)/(x - z); q_len receives the
 * trimmed length (n-1 untrimmed values are written).  y_le == NULL means "use p(z)". */
int32_t kzg_quotient(kzg_ctx* ctx, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32],
                     const uint8_t* y_le, uint8_t* q_le32, uint64_t* q_len);
/* Evaluation::generate_proof (src/polynomial.rs:260-269).  If y_expected_le is given and
 * differs from p(z): KZG_E_REMAINDER_NONZERO (KZG_E_CONSTANT_POLY for a constant polynomial),
 * mirroring divide_by_root's errors. */
int32_t kzg_open(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t n,
                 const uint8_t z_le[32], const uint8_t* y_expected_le, uint8_t y_out_le[32],
                 uint8_t proof_c48[48]);
/* plain G1 MSM over the first n bases, no trimming semantics (BASELINE config C3) */
int32_t kzg_msm_g1(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_le32, uint64_t n,
                   uint8_t out_c48[48]);

/* the same call split in two, for callers that commit one polynomial after another (the reference's benches
 * and CLI do: src/polynomial.rs:200-215 called in a loop): submit returns once the work is queued, wait blocks
 * for the 48-byte result.  At most L tickets may be outstanding (L = option "lanes", default 4 = the most; a
 * submit beyond that fails with KZG_E_BAD_ARG); the host->device copy of a newer call overlaps the kernels of the older ones.  `scalars_le32` must stay valid (and should be page-locked for the copy
 * to be asynchronous) until its ticket has been waited for. */
int32_t kzg_msm_g1_submit(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_le32, uint64_t n,
                          int32_t* ticket);
int32_t kzg_msm_g1_wait(kzg_ctx* ctx, int32_t ticket, uint8_t out_c48[48]);
/* Calls run on L internal lanes (own streams and scratch memory) of the context in turn, so the copy and the bulk
 * kernels of one call also overlap the latency-bound phases (inversion chains, bucket reduction) of the others.
 * Option "msm_coalesce" = B (1..4, default 1; like "lanes" fixed by the first submit) additionally FUSES consecutive
 * submitted calls that share the setup and the length: a submit then only stages its scalars, and B staged calls are
 * launched together as ONE bucket problem with B times the buckets (same table, one sort / accumulation for all, the
 * bucket reduction per call).  Up to L x B tickets may then be outstanding.  Staged calls are also launched as soon as
 * somebody waits for one of them, calls kzg_ctx_sync / kzg_ctx_timer_stop, or submits a call of another shape, so no
 * ticket ever waits for calls that do not come; every ticket returns the bytes of the one-shot call.  Worth 6 % at
 * n = 2^20 and 30 % at n = 2^17 (the share of one GPU out of eight), for callers that keep >= 2 B calls in flight.
 * Variant with the scalars already in device memory (ready before the call: the engine does not know the producer's
 * stream): */
int32_t kzg_msm_g1_submit_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n,
                              int32_t* ticket);

/* ---- blob batches (BASELINE config C4): `blobs` polynomials of n coefficients each ------ */
int32_t kzg_commit_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32,
                         uint64_t blobs, uint64_t n, uint8_t* out_c48);
int32_t kzg_open_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32,
                       uint64_t blobs, uint64_t n, const uint8_t* z_le32, uint8_t* y_le32,
                       uint8_t* proofs_c48);
/* Polynomial::commit + Evaluation::generate_proof of every blob (src/polynomial.rs:200-215, 260-269) with ONE
 * host->device pass over the coefficients: the EIP-4844-style "commit + proof" unit of BASELINE config C4. */
int32_t kzg_commit_open_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32,
                              uint64_t blobs, uint64_t n, const uint8_t* z_le32, uint8_t* commits_c48,
                              uint8_t* y_le32, uint8_t* proofs_c48);

/* ---- device-resident variants (inputs already in HBM; results still land on the host) --- */
int32_t kzg_msm_g1_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n,
                       uint8_t out_c48[48]);
/* partial sum of a shard, left on the device as a KZG_G1_PARTIAL_BYTES blob (for all-gather) */
int32_t kzg_msm_g1_partial_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev,
                               uint64_t n, void* partial_dev);
/* ---- multi-GPU (one process per GPU, BASELINE config C3): the per-GPU partial sums are exchanged by direct stores
 * into peer memory over NVLink and folded + compressed by the same kernel (no NCCL call on the data path).
 *   kzg_p2p_init     allocates this rank's mailbox and returns its CUDA IPC handle (KZG_P2P_HANDLE_BYTES bytes);
 *   kzg_p2p_connect  takes the handles of ALL ranks, rank-major (world x KZG_P2P_HANDLE_BYTES; how they travel
 *                    between the processes is the caller's business: torch.distributed, MPI, a file ...);
 *   kzg_msm_g1_sharded_dev  is collective: every rank calls it with its shard, every rank receives the full result.
 * A rank that never shows up turns into KZG_E_CUDA after 20 s (option "p2p_timeout_ms") instead of a hang, and it does so
 * on EVERY rank: the rank that gives up leaves a sticky mark in all mailboxes, so a peer that arrives late fails too
 * instead of returning a result alone; the contexts then refuse further collective calls.  Build the tables
 * (kzg_setup_prepare_msm) and rendezvous on the host before the first collective call: a lazy 2 GiB table build on one
 * rank is exactly the kind of skew the timeout is for.  Collective calls are numbered by their own counter: every rank
 * must issue the same sequence of kzg_msm_g1_sharded_* calls (other calls in between do not matter). */
#define KZG_P2P_HANDLE_BYTES 64
int32_t kzg_p2p_init(kzg_ctx* ctx, int32_t world, int32_t rank, uint8_t handle_out[KZG_P2P_HANDLE_BYTES]);
int32_t kzg_p2p_connect(kzg_ctx* ctx, const uint8_t* handles);
int32_t kzg_msm_g1_sharded_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n,
                               uint8_t out_c48[48]);
/* the same with up to `lanes` x `msm_coalesce` collective calls in flight (collect with kzg_msm_g1_wait); every rank
 * must use the same two options: fused calls take consecutive sequence numbers and share one exchange launch */
int32_t kzg_msm_g1_sharded_submit_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n,
                                      int32_t* ticket);
/* ... and with this rank's shard of the scalars in a HOST buffer (copied on the lane's copy stream, like kzg_msm_g1_submit) */
int32_t kzg_msm_g1_sharded_submit(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_le32, uint64_t n,
                                  int32_t* ticket);
/* fold `count` partial sums (device) into one compressed point */
int32_t kzg_g1_combine_dev(kzg_ctx* ctx, const void* partials_dev, uint64_t count,
                           uint8_t out_c48[48]);
int32_t kzg_commit_batch_dev(kzg_ctx* ctx, const kzg_setup* s, const void* coeffs_dev,
                             uint64_t blobs, uint64_t n, void* out_c48_dev);
/* Batches compare the TRIMMED length of the longest row with the setup, like Polynomial::commit
 * (src/polynomial.rs:201-205): rows may be longer than the setup when the excess is zero padding (the _dev variants
 * then spend one pass and one synchronisation on the check).  Any number of rows. */
int32_t kzg_open_batch_dev(kzg_ctx* ctx, const kzg_setup* s, const void* coeffs_dev,
                           uint64_t blobs, uint64_t n, const void* z_dev, void* y_dev,
                           void* proofs_c48_dev);
/* commit + open of the same rows, the two MSM batches on two lanes of the context (inputs ready on the engine stream) */
int32_t kzg_commit_open_batch_dev(kzg_ctx* ctx, const kzg_setup* s, const void* coeffs_dev,
                                  uint64_t blobs, uint64_t n, const void* z_dev, void* commits_c48_dev,
                                  void* y_dev, void* proofs_c48_dev);
int32_t kzg_evaluate_dev(kzg_ctx* ctx, const void* coeffs_dev, uint64_t n, const uint8_t z_le[32],
                         uint8_t y_le[32]);

/* ---- single-element host helpers (public per-element Rust API; not the hot path) -------- */
/* Scalar::mul/add/sub/neg/pow (src/scalar.rs:111-223) on canonical 32-byte LE values */
void kzg_fr_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);
void kzg_fr_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);
void kzg_fr_sub(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);
void kzg_fr_neg(const uint8_t a[32], uint8_t out[32]);
void kzg_fr_pow(const uint8_t a[32], uint64_t n, uint8_t out[32]);
/* Scalar::from_le_bytes reduction (src/scalar.rs:54-61) */
void kzg_fr_reduce(const uint8_t a[32], uint8_t out[32]);
/* G1Point::add/sub/mult/from_scalar (src/curves.rs:50-96) on 48-byte compressed points */
int32_t kzg_g1_add(const uint8_t a[48], const uint8_t b[48], uint8_t out[48]);
int32_t kzg_g1_sub(const uint8_t a[48], const uint8_t b[48], uint8_t out[48]);
int32_t kzg_g1_mul(const uint8_t p[48], const uint8_t k_le[32], uint8_t out[48]);
void kzg_g1_from_scalar(const uint8_t k_le[32], uint8_t out[48]);
/* on-curve / encoding check of one compressed point: KZG_OK or KZG_E_BAD_POINT */
int32_t kzg_g1_check(const uint8_t p[48]);
/* G2Point::{from_scalar, sub} (src/curves.rs:225-249) + add / neg on 96-byte compressed points */
void kzg_g2_from_scalar(const uint8_t k_le[32], uint8_t out[96]);
int32_t kzg_g2_add(const uint8_t a[96], const uint8_t b[96], uint8_t out[96]);
int32_t kzg_g2_sub(const uint8_t a[96], const uint8_t b[96], uint8_t out[96]);
int32_t kzg_g2_neg(const uint8_t a[96], uint8_t out[96]);
int32_t kzg_g2_check(const uint8_t p[96]);
/* bilinear_map(a1, b1) == bilinear_map(a2, b2) (src/curves.rs:355-371; only equality is ever consumed,
 * src/polynomial.rs:293): one product of Miller loops + one final exponentiation, host code */
int32_t kzg_pairing_equal(const uint8_t a1[48], const uint8_t b1[96], const uint8_t a2[48],
                          const uint8_t b2[96], int32_t* equal);
/* Evaluation::verify_proof (src/polynomial.rs:276-294): e(proof, [s]_2 - [z]_2) == e(C - [y]_1, [1]_2);
 * g2_tau = setup[1].g2.  *ok receives 1 / 0. */
int32_t kzg_verify_proof(const uint8_t proof[48], const uint8_t commitment[48], const uint8_t z_le[32],
                         const uint8_t y_le[32], const uint8_t g2_tau[96], int32_t* ok);

/* ---- diagnostics ------------------------------------------------------------------------ */
/* integer-pipe roofline probe: independent 32x32->64 products (IMAD.WIDE.U32, zero addend) on every SM, each folded
 * back with one LOP3; returns wide multiplies per second = the issue rate of the multiplier pipe (BASELINE.md §4
 * "measured integer peak"; nominal 64 lanes / clk / SM).  The additions of a multiply-ACCUMULATE come on top. */
int32_t kzg_probe_imad_peak(kzg_ctx* ctx, double* wide_mac_per_s);
/* sustained Montgomery Fq multiplications per second of the shipped field kernel (all SMs busy) */
int32_t kzg_probe_fq_mul(kzg_ctx* ctx, double* fq_mul_per_s);
/* field self-test on the device: n Montgomery products a*b for a_i = U(seed,2i), b_i = U(seed,2i+1)
 * reduced into the field; writes canonical results (48 or 32 bytes LE each) for the oracle */
int32_t kzg_selftest_fq_mul(kzg_ctx* ctx, const uint8_t* a_le48, const uint8_t* b_le48, uint64_t n,
                            uint8_t* out_le48);

#ifdef __cplusplus
}
#endif
#endif /* KZG_B200_H */
