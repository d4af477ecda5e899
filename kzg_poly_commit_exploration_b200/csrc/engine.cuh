// Internal engine declarations shared by the .cu translation units of libkzg_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/kzg_b200.h"
#include "g1.cuh"
#include "recode.cuh"

namespace kzg {

enum ScratchSlot {
  SCR_SCALARS = 0,   // staged scalars / coefficients (host -> device)
  SCR_KEYS_A, SCR_KEYS_B, SCR_VALS_A, SCR_VALS_B, SCR_SORT_TMP,
  SCR_BUCKET_START, SCR_BUCKETS, SCR_RED_A, SCR_RED_B, SCR_HEAVY,
  SCR_XYZZ_TMP, SCR_PREFIX, SCR_QUOT, SCR_SEG, SCR_OUT, SCR_MISC, SCR_FR_A, SCR_FR_B,
  SCR_BATCH_PTS, SCR_BATCH_AFF, SCR_PARTIALS, SCR_HEAVY_PARTS,
  SCR_TREE_A, SCR_TREE_B, SCR_TREE_PREFIX, SCR_TREE_TOT, SCR_TREE_SEG, SCR_TREE_IOTA, SCR_ORDER,
  SCR_L2_A, SCR_L2_B, SCR_L2_PREFIX, SCR_L2_TOT, SCR_L2_SEG, SCR_L2_IOTA, SCR_L2_START, SCR_L2_VALS, SCR_L2_MISC,
  SCR_COUNT
};

constexpr int KZG_MAX_PEERS = 32;
constexpr int KZG_MAX_LANES = 4;      // calls in flight per context (kzg_msm_g1_submit*): one lane (streams + scratch) each
constexpr int KZG_MAX_COALESCE = 4;   // submitted calls fused into one bucket problem (option msm_coalesce, msm_bucket_multi)
constexpr int KZG_MAX_INFLIGHT = KZG_MAX_LANES * KZG_MAX_COALESCE;   // tickets that can be out at once
constexpr int KZG_P2P_DEPTH = 2 * KZG_MAX_INFLIGHT;   // mailbox slots per peer: with at most F calls in flight per rank, call
                                     // k + F may overwrite the slot of call k - F, which every rank has finished reading (msm_tail.cu)
}  // namespace kzg

struct kzg_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host->device staging of batch chunks, overlapped with compute on `stream`
  cudaEvent_t ev_copy[8] = {};
  std::string err;
  uint64_t launches = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float last_ms[3] = {0.f, 0.f, 0.f};
  bool timers_stale = false;           // a wait finished a call whose phase events have not been read yet (one_ctx_last_kernel_ms does)
  cudaEvent_t ev_k[2] = {nullptr, nullptr};   // around the single dominant kernel of the last bulk call
  cudaEvent_t ev_t[2] = {nullptr, nullptr};
  void* scratch[kzg::SCR_COUNT] = {};
  size_t scratch_cap[kzg::SCR_COUNT] = {};
  void* pinned = nullptr;
  size_t pinned_cap = 0;
  kzg::G1Affine* gtable = nullptr;     // [16][32768] multiples of the generator (setup generation)
  void* g2table = nullptr;             // [20][4096] signed-window multiples of the G2 generator (setup_g2.cu)
  int opt_msm_window = 0;              // 0 = auto
  int opt_fixed_window = 0;            // 0 = auto: the widest window whose table fits the memory budget (setup.cu auto_fixed_window)
  int64_t opt_fixed_table_max_bytes = 0;   // budget of that choice; 0 = 66 % of the device memory free at the time
  int opt_msm_sort = -1;               // bucket sort of the MSM entries: 0 / -1 = counting sort (count / tiled scan / scatter), 1 = cub radix sort, 2 = two-level partition sort (both kept for comparison)
  int opt_acc_order = 1;               // hand the buckets to k_bucket_acc sorted by length (k_bucket_order)
  int opt_tree_lean = 4;               // pair-addition kernel of the dense halving rounds: 0 = points held in registers, 4 / 5 / 6 = reloading variant at that many blocks per SM
  int opt_tree_lean_r1 = 4;            // the same for round 1 (table gathers)
  int opt_reduce_chunk = 16;
  int opt_heavy_cap = 1024;
  int opt_acc_split = -1;              // log2(threads per bucket), -1 = auto
  int opt_r1_row_minor = 1;            // blob batches, round 1: neighbouring threads = neighbouring rows at ONE coefficient index, so the lookups in
                                       // flight fall into ~1 GB of the fixed-base table instead of all 102 GiB (DRAM row locality: 24.6 -> 26.4 k blobs/s)
  int opt_tree_min_rows = 512;         // batches of at least this many rows take the affine tree path (batch_tree.cu)
  int opt_tree_chunk_rows = 2048;      // rows reduced together (bounds the scratch: ~8 KiB * n/4096 * W/2 per row... see DESIGN)
  int64_t opt_tree_min_pairs = 1 << 19;  // stop halving when a round would have fewer independent additions
  int opt_inv_top_fan = 0;             // values per thread in the fused top kernel of the batch inversion: 4 or 8; 0 = auto:
                                       // 4 for a call on its own (shortest chain), 8 for calls submitted through the lanes
                                       // (half as many binary-GCD inversions; n = 2^20: 5.51 -> 5.43 ms with four in flight,
                                       // 6.45 -> 6.61 ms alone)
  bool throughput_mode = false;        // set around the MSM of a submitted call (abi.cu submit_common)
  int64_t opt_inv_top = 1 << 18;       // batch inversion: values left for the fused top kernel
  int opt_msm_tree_rounds = -1;        // affine halving rounds before the XYZZ accumulate: -1 = auto (log2(entries) - 19, 0..6)
  int opt_acc_min_entries = 16;        // keep splitting a bucket across threads while a thread still gets more than this
                                       // (measured at n = 2^17 .. 2^18, the per-GPU share of a sharded 2^20 MSM)
  int opt_tree_k = 0;                  // pairs per thread in the halving rounds: 0 = auto (fill the last wave), else 8..16
  int opt_coop_tail = 1;               // 4-warp cooperative additions in the bucket-reduction tail
  int opt_tree_r1_ool = 0;             // round-1 pair additions from the translation unit with the product out of line (msm_tree_r1.cu)
  int opt_den_side = 0;                // round-1 denominators of calls in flight as this many resident blocks per SM on a high-priority side stream (0 = off)
  cudaStream_t side_stream = nullptr;
  cudaEvent_t side_ev[2] = {nullptr, nullptr};
  int opt_msm_l2 = -1;                 // bucket sums reduced as a second bucket problem (msm.cu finish_level2): 0 off, 1 on, -1 = by window (on from 18 bits)
  int opt_l2_rounds = -1;              // affine halving rounds of that second problem, -1 = auto
  int opt_reduce_lk = -1;              // level 1 of the bucket reduction: log2 of the buckets a lane walks serially (msm_tail.cu k_reduce_l1k); -1 = auto
  int opt_acc_minblocks = 0;           // resident blocks per SM the accumulate kernel is compiled for (0 = auto: 4 after halving rounds, else 3)
  int sm_count = 148;
  // submission queue of kzg_msm_g1_submit / kzg_msm_g1_wait (host->device copy of call k+1 beside call k)
  struct Pending {
    void* scalars = nullptr; size_t cap = 0;   // device staging of this slot
    void* result = nullptr;                    // device: one G1Xyzz + 48 compressed bytes + status word
    uint8_t* out_host = nullptr;               // pinned, 64 bytes (result + status of a sharded call)
    bool sharded = false;
    cudaEvent_t copied = nullptr, done = nullptr;
    int32_t status = 0;                        // failure of the launch that carried this call (reported by the wait)
    bool busy = false;
    bool launched = false;                     // false: still in the open group (msm_coalesce), nothing issued yet
    bool host = false;
    uint32_t ticket = 0;
    int lane_index = 0;
    const uint8_t* dev_scalars = nullptr;      // what the launch reads: the staging buffer or the caller's device pointer
  } pending[kzg::KZG_MAX_INFLIGHT];            // ticket t lives in slot t mod (lanes x coalesce)
  uint32_t next_ticket = 0;
  // Calls fused into one bucket problem (option msm_coalesce = B > 1, msm.cu msm_bucket_multi): a submit only stages its
  // scalars and joins the OPEN group; the group is launched on the next lane when it holds B calls, when a call with
  // another setup / length / kind arrives, or when somebody waits for (or synchronises with) one of its tickets.
  int opt_msm_coalesce = 1;
  int coalesce_in_use = 0;                     // latched with lanes_in_use by the first submit
  struct OpenGroup {
    int count = 0;
    uint32_t first_ticket = 0;
    const kzg_setup* s = nullptr;
    uint64_t n = 0;
    bool sharded = false;
  } open_group;
  uint32_t groups_launched = 0;                // group g runs on lane g mod L
  int opt_setup_tree_min = 1 << 18;    // setup generation: affine-tree path from this many points on (0 = never)
  int opt_lanes = 4;                   // calls in flight (1 .. KZG_MAX_LANES); fixed once the first ticket is out
  int lanes_in_use = 0;                // opt_lanes latched by the first submit
  kzg_ctx* lane[kzg::KZG_MAX_LANES] = {};   // ticket t runs on lane t mod L; lane 0 is the context itself, the others are hidden
                                       // sibling contexts (own streams and scratch), so that the latency-bound phases of one
                                       // call (inversion chains, bucket reduction) overlap the bulk kernels of the others
  // multi-GPU exchange over peer memory (msm_tail.cu k_exchange_combine): box[r] = rank r's mailbox as seen from here
  int p2p_world = 0, p2p_rank = 0;
  bool p2p_connected = false;
  bool p2p_ipc = true;                 // peers' mailboxes are CUDA IPC mappings (one process per GPU); false: plain peer pointers
  bool p2p_failed = false;             // a collective call timed out or was abandoned: the sequence numbers are out of step for good
  uint32_t p2p_seq = 0;                // number of the last collective (sharded) call issued on this context
  void* p2p_box[kzg::KZG_MAX_PEERS] = {};
  int sm_clock_khz = 1965000;          // cudaDevAttrClockRate: clock64() ticks per ms, for the exchange timeout
  long long p2p_timeout_cycles = 20000ll * 1965000;   // spinning time before a missing peer becomes an error (20 s; option p2p_timeout_ms)
  uint64_t reserved_n = 0;             // largest sum this lane's arenas have been sized for ahead of time (reserve_call)
  int last_lane = 0;                   // lane the most recent bulk call ran on (its events feed kzg_ctx_last_kernel_ms)
  // single-process multi-GPU (kzg_ctx_create_multi, abi.cu): this context is the head of a GROUP of per-device member
  // contexts; members[0] == this.  Every member has a host worker thread that issues its launches, so that one caller
  // drives all devices in parallel.  Empty for a plain single-device context.
  std::vector<kzg_ctx*> members;
  kzg_ctx* group_head = nullptr;       // set on every member (also on the head itself) of a group
  void* group_state = nullptr;         // abi.cu: worker threads, group tickets (owned by the head)
  bool group_shares_devices = false;   // head: two members sit on one device (testing on a single-GPU box): submits run in two phases
  int opt_setup_layout = 0;            // group setups: 0 = auto (replicate up to 2^16 points, else shard by point range), 1 = range, 2 = replicated
};

struct kzg_setup {
  kzg_ctx* ctx = nullptr;
  uint64_t n = 0;
  uint64_t first = 0;                  // index of bases[0] in the powers-of-tau sequence (range sharding)
  kzg::G1Affine* bases = nullptr;      // [n] affine Montgomery, (0,0) = infinity
  uint8_t* g2_c96 = nullptr;           // [g2_n] compressed G2 powers (device), items 0..g2_n of this setup
  uint64_t g2_n = 0;
  // folded-window table for the bucket MSM: tab[j*n + i] = 2^(c*j) * P_i ; tab window 0 == bases
  int msm_c = 0, msm_w = 0;
  // Entries sit on a 128-byte pitch (kzg::MSM_TAB_PITCH): the round-1 gathers of the bucket accumulation fetch
  // whole 128-byte DRAM lines whatever they ask for, and a 96-byte pitch makes one entry in two straddle two lines
  // (ncu: 166 B of DRAM traffic per 48-byte x read); on the 128-byte pitch every random point read is exactly one line.
  char* msm_tab = nullptr;             // [(w) * n] x 128 B: G1Affine + 32 B of padding
  // fixed-base lookup table: fix[((i*w + j) << (c-1)) + d-1] = d * 2^(c*j) * P_i
  int fix_c = 0, fix_w = 0;
  uint64_t fix_n = 0;
  kzg::G1Affine* fix_tab = nullptr;
  // setup of a multi-device context: one per-device setup per member.  layout 1: shard g holds the contiguous point range
  // [shard_off[g], shard_off[g+1]) of the n items (bases == nullptr here); layout 2: every shard holds all n items.
  int layout = 0;
  std::vector<kzg_setup*> shards;
  std::vector<uint64_t> shard_off;
};

namespace kzg {

#define KZG_CUDA_OK(ctx, call)                                                              \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                     \
      return (e__ == cudaErrorMemoryAllocation) ? KZG_E_NOMEM : KZG_E_CUDA;                 \
    }                                                                                       \
  } while (0)

#define KZG_TRY(expr)                    \
  do {                                   \
    int32_t st__ = (expr);               \
    if (st__ != KZG_OK) return st__;     \
  } while (0)

// count + check a kernel launch
#define KZG_LAUNCHED(ctx)                                   \
  do {                                                      \
    (ctx)->launches++;                                      \
    KZG_CUDA_OK(ctx, cudaGetLastError());                   \
  } while (0)

int32_t scratch_get(kzg_ctx* ctx, int slot, size_t bytes, void** out);
int32_t pinned_get(kzg_ctx* ctx, size_t bytes, void** out);
inline int32_t fail(kzg_ctx* ctx, int32_t st, const char* msg) {
  if (ctx) ctx->err = msg;
  return st;
}

// pitch of the folded-window MSM table in bytes (one G1Affine per 128-byte DRAM line)
constexpr uint32_t MSM_TAB_PITCH = 128;
constexpr uint32_t DENSE_PITCH = 96;   // sizeof(G1Affine): the dense point arrays between the halving rounds

// ---- window geometry ------------------------------------------------------------------
// window geometry and signed-digit recoding of folded scalars: recode.cuh (windows_for, SignedDigits)
int auto_msm_window(uint64_t n);
int auto_tree_rounds(uint64_t entries);   // msm.cu: affine halving rounds for W*n bucket entries

// ---- poly.cu (K4: Horner evaluation + quotient scan) -----------------------------------
// coeffs_dev: blobs x n x 32 B canonical-or-not LE; z_dev: blobs x 32 B; outputs canonical.
// q_dev may be null (evaluate only). y_dev: blobs x 32 B.
int32_t poly_eval_quotient(kzg_ctx* ctx, const uint8_t* coeffs_dev, uint64_t blobs, uint64_t n,
                           const uint8_t* z_dev, uint8_t* y_dev, uint8_t* q_dev);

// ---- setup.cu (K3) -------------------------------------------------------------------------
int32_t setup_generate_dev(kzg_ctx* ctx, const uint8_t secret_be[32], uint64_t first, uint64_t n,
                           G1Affine* bases_dev);
int32_t batch_to_affine(kzg_ctx* ctx, const G1Xyzz* in_dev, uint64_t n, G1Affine* out_dev);
int32_t compress_points(kzg_ctx* ctx, const G1Affine* in_dev, uint64_t n, uint8_t* out_dev);
int32_t uncompress_points(kzg_ctx* ctx, const uint8_t* in_dev, uint64_t n, G1Affine* out_dev, int* bad_index);
int32_t build_msm_table(kzg_ctx* ctx, kzg_setup* s, int c);
int32_t build_fixed_table(kzg_ctx* ctx, kzg_setup* s, uint64_t n_points, int c);
uint64_t fixed_table_bytes(uint64_t n_points, int c);
int auto_fixed_window(kzg_ctx* ctx, uint64_t n_points);   // widest window whose table fits the memory budget

// ---- setup_g2.cu (K3') ---------------------------------------------------------------------
// [s^(first+i)]G2, i < n, as 96-byte compressed points (device)
int32_t setup_generate_g2_dev(kzg_ctx* ctx, const uint8_t secret_be[32], uint64_t first, uint64_t n,
                              uint8_t* out_c96_dev);
// on-curve / encoding check of n compressed G2 points (device); bad_index = first malformed point or -1
int32_t validate_g2_points(kzg_ctx* ctx, const uint8_t* in_c96_dev, uint64_t n, int64_t* bad_index);

// ---- msm.cu (K2) -----------------------------------------------------------------------------
// result_dev: one G1Xyzz (device).  Uses s->msm_tab (built on demand).
int32_t msm_bucket(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_dev, uint64_t n,
                   G1Xyzz* result_dev);
// B sums of n scalars each over the same bases as ONE bucket problem (B <= KZG_MAX_COALESCE); results_dev: B G1Xyzz
int32_t msm_bucket_multi(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* const* scalars_dev, int B, uint64_t n,
                         G1Xyzz* results_dev);
// sum_b (b+1) * A_b, A_b = sum_k parts[b*T + k] with T = 2^logT partials per bucket (msm_tail.cu)
int32_t reduce_buckets(kzg_ctx* ctx, const G1Xyzz* parts, uint32_t M, uint32_t logT, G1Xyzz* result_dev);
// the same for B consecutive bucket sets of M <= 32768 buckets each, one launch set (blockIdx.y = set); results_dev: B G1Xyzz
int32_t reduce_buckets_multi(kzg_ctx* ctx, const G1Xyzz* parts, uint32_t M, uint32_t logT, int B, G1Xyzz* results_dev);
// out_dev[0..count) = infinity
int32_t set_infinity(kzg_ctx* ctx, G1Xyzz* out_dev, uint32_t count);
// sum `count` XYZZ points (device) into out_dev[0]
int32_t sum_points(kzg_ctx* ctx, const G1Xyzz* in_dev, uint64_t count, G1Xyzz* out_dev);
// XYZZ (device) x count -> 48-byte compressed (device)
int32_t finalize_points(kzg_ctx* ctx, const G1Xyzz* in_dev, uint64_t count, uint8_t* out_c48_dev);

// one partial sum per GPU -> every GPU's mailbox (peer stores), fold, compress: out_c48_dev[48], *status_dev = 0 / 1 (timeout)
struct PeerMailbox {
  G1Xyzz slot[KZG_P2P_DEPTH][KZG_MAX_PEERS];     // [call number mod depth][source rank]
  uint32_t flag[KZG_P2P_DEPTH][KZG_MAX_PEERS];   // sequence number of the partial in the slot
  uint32_t dead[KZG_MAX_PEERS];                  // != 0: that rank gave up (timed out) in the call of this number — sticky
};
struct PeerTable { PeerMailbox* box[KZG_MAX_PEERS]; };
// `ctx` supplies the stream, `owner` the mailboxes; seq = number of this collective call (same on every rank).
// calls > 1: mine_dev[k] belongs to collective call seq + k, its output and status word sit k * out_stride bytes on.
int32_t exchange_combine(kzg_ctx* ctx, const kzg_ctx* owner, uint32_t seq, const G1Xyzz* mine_dev, uint8_t* out_c48_dev,
                         uint32_t* status_dev, int calls = 1, uint32_t out_stride = 0);

// ---- batch.cu (K5) ---------------------------------------------------------------------------
// blobs MSMs of n scalars each against the fixed-base table; out: blobs XYZZ
int32_t msm_fixed_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_dev, uint64_t blobs,
                        uint64_t n, uint64_t stride_scalars, G1Xyzz* out_dev);

// ---- batch_tree.cu (K5') ---------------------------------------------------------------------
// same contract as msm_fixed_batch; affine pairwise tree with grid-wide shared inversions (large batches)
int32_t msm_fixed_batch_tree(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_dev, uint64_t rows,
                             uint64_t n, uint64_t stride_scalars, G1Xyzz* out_dev);
// N fixed-base multiplications of one base through the same affine trees; affine results (setup generation)
int32_t fixed_base_points_tree(kzg_ctx* ctx, const G1Affine* tab, int c, int W, const uint8_t* scalars_dev, uint64_t N,
                               G1Affine* out_dev);

// grid-wide batch inversion: v[i] <- 1 / v[i], i < N, no zeros; ws holds invert_scratch_elems(N, top) Fq
size_t invert_scratch_elems(uint64_t N, uint64_t top);
int32_t batch_invert_inplace(kzg_ctx* ctx, Fq* v, uint64_t N, Fq* ws, uint64_t top);

// ---- msm_tree.cu (K2') ----------------------------------------------------------------------
// `rounds` halvings of every bucket's sorted entries in affine coordinates (shared inversions).  In: bucket
// bounds start[0..M] over vals (table index | sign << 31).  Out: dense affine points + their bounds.
int32_t msm_tree_rounds(kzg_ctx* ctx, const uint32_t* start_dev, const uint32_t* vals_dev, const char* tab_pitched,
                        uint32_t M, uint64_t entries, int rounds, const G1Affine** pts_out,
                        const uint32_t** seg_out, const uint32_t** iota_out, uint64_t* entries_out,
                        uint32_t tab_pitch = MSM_TAB_PITCH, int level = 1);
// result[k] = 2^shift * res2[2k] + res2[2k + 1], k < calls (msm_tail.cu; the two halves of a second-level bucket reduction)
int32_t combine_level2(kzg_ctx* ctx, const G1Xyzz* res2, int shift, int calls, G1Xyzz* results_dev);

// ---- engine_one.cu: the single-device engine API (one_* = the kzg_* entry point of the same name on ONE device) ----
int32_t one_ctx_create(kzg_ctx** out, int32_t device_id);
void one_ctx_destroy(kzg_ctx* ctx);
const char* one_last_error(const kzg_ctx* ctx);
int32_t one_ctx_sync(kzg_ctx* ctx);
int32_t one_ctx_timer_start(kzg_ctx* ctx);
int32_t one_ctx_timer_stop(kzg_ctx* ctx, float* ms);
uint64_t one_ctx_launch_count(const kzg_ctx* ctx);
void* one_ctx_stream(const kzg_ctx* ctx);
float one_ctx_last_kernel_ms(const kzg_ctx* ctx, int32_t which);
int32_t one_ctx_set_option(kzg_ctx* ctx, const char* name, int64_t value);
int32_t one_setup_generate(kzg_ctx* ctx, const uint8_t secret_be[32], uint64_t first, uint64_t n, kzg_setup** out);
int32_t one_setup_import_g1(kzg_ctx* ctx, const uint8_t* g1_c48, uint64_t n, kzg_setup** out);
int32_t one_setup_export_g1(kzg_ctx* ctx, const kzg_setup* s, uint64_t first, uint64_t count, uint8_t* out_c48);
uint64_t one_setup_len(const kzg_setup* s);
int32_t one_setup_generate_g2(kzg_ctx* ctx, kzg_setup* s, const uint8_t secret_be[32], uint64_t count);
int32_t one_setup_import_g2(kzg_ctx* ctx, kzg_setup* s, const uint8_t* g2_c96, uint64_t count);
int32_t one_setup_export_g2(kzg_ctx* ctx, const kzg_setup* s, uint64_t first, uint64_t count, uint8_t* out_c96);
uint64_t one_setup_g2_len(const kzg_setup* s);
int32_t one_msm_plan(const kzg_ctx* ctx, const kzg_setup* s, uint64_t n, int32_t plan[5]);
void one_setup_destroy(kzg_setup* s);
int32_t one_setup_prepare_msm(kzg_ctx* ctx, kzg_setup* s, int32_t window_bits);
int32_t one_setup_prepare_fixed(kzg_ctx* ctx, kzg_setup* s, uint64_t n_points, int32_t window_bits);
int32_t one_msm_g1_partial_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, void* partial_dev);
int32_t one_msm_g1_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, uint8_t out_c48[48]);
int32_t one_msm_g1(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_le32, uint64_t n, uint8_t out_c48[48]);
int32_t one_msm_g1_submit(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_le32, uint64_t n, int32_t* ticket);
int32_t one_msm_g1_submit_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, int32_t* ticket);
int32_t one_msm_g1_sharded_submit_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, int32_t* ticket);
int32_t one_msm_g1_sharded_submit(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_le32, uint64_t n, int32_t* ticket);
int32_t one_msm_g1_wait(kzg_ctx* ctx, int32_t ticket, uint8_t out_c48[48]);
int32_t one_p2p_init(kzg_ctx* ctx, int32_t world, int32_t rank, uint8_t handle_out[64]);
int32_t one_p2p_connect(kzg_ctx* ctx, const uint8_t* handles);
int32_t one_msm_g1_sharded_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, uint8_t out_c48[48]);
int32_t one_g1_combine_dev(kzg_ctx* ctx, const void* partials_dev, uint64_t count, uint8_t out_c48[48]);
int32_t one_commit(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t n, uint8_t out_c48[48]);
int32_t one_evaluate_dev(kzg_ctx* ctx, const void* coeffs_dev, uint64_t n, const uint8_t z_le[32], uint8_t y_le[32]);
int32_t one_evaluate(kzg_ctx* ctx, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32], uint8_t y_le[32]);
int32_t one_quotient(kzg_ctx* ctx, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32], const uint8_t* y_le, uint8_t* q_le32, uint64_t* q_len);
int32_t one_open(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32], const uint8_t* y_expected_le, uint8_t y_out_le[32], uint8_t proof_c48[48]);
int32_t one_commit_batch_dev(kzg_ctx* ctx, const kzg_setup* s, const void* coeffs_dev, uint64_t blobs, uint64_t n, void* out_c48_dev);
int32_t one_open_batch_dev(kzg_ctx* ctx, const kzg_setup* s, const void* coeffs_dev, uint64_t blobs, uint64_t n, const void* z_dev, void* y_dev, void* proofs_c48_dev);
int32_t one_commit_open_batch_dev(kzg_ctx* ctx, const kzg_setup* s, const void* coeffs_dev, uint64_t blobs, uint64_t n, const void* z_dev, void* commits_c48_dev, void* y_dev, void* proofs_c48_dev);
int32_t one_commit_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t blobs, uint64_t n, uint8_t* out_c48);
int32_t one_open_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t blobs, uint64_t n, const uint8_t* z_le32, uint8_t* y_le32, uint8_t* proofs_c48);
int32_t one_commit_open_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t blobs, uint64_t n, const uint8_t* z_le32, uint8_t* commits_c48, uint8_t* y_le32, uint8_t* proofs_c48);
int32_t one_probe_imad_peak(kzg_ctx* ctx, double* wide_mac_per_s);
int32_t one_probe_fq_mul(kzg_ctx* ctx, double* fq_mul_per_s);
int32_t submit_common(kzg_ctx* ctx, const kzg_setup* s, const void* scalars, uint64_t n, bool host, bool sharded, int32_t* ticket);
void copy_options(kzg_ctx* dst, const kzg_ctx* src);
// launch the calls a context still holds back for fusing (option msm_coalesce); their failures are reported by their waits
int32_t flush_submitted(kzg_ctx* ctx);
// tickets that can be out at once (lanes x fused calls); fixes both numbers if no call has been submitted yet
uint32_t queue_capacity(kzg_ctx* ctx);
int32_t reserve_call(kzg_ctx* ctx, const kzg_setup* s, const void* scalars, uint64_t n, bool host);
// stages the coefficients, runs the scan, applies divide_by_root's error rules; *q_dev = trimmed quotient on the device
int32_t quotient_core(kzg_ctx* ctx, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32], const uint8_t* y_le,
                      uint8_t y_out[32], uint8_t** q_dev, uint64_t* q_len);
// mailbox of a context for the peer exchange, and the wiring when all ranks live in one process (plain peer pointers)
int32_t p2p_alloc_box(kzg_ctx* ctx, int32_t world, int32_t rank);
int32_t p2p_attach_local(kzg_ctx* ctx, void* const* boxes);
// blobs x n coefficients on the device: is any coefficient at index >= limit non-zero mod r?  (poly.cu; synchronises)
int32_t any_nonzero_beyond(kzg_ctx* ctx, const uint8_t* coeffs_dev, uint64_t blobs, uint64_t n, uint64_t limit, bool* any);

// ---- probe.cu --------------------------------------------------------------------------------
int32_t probe_imad(kzg_ctx* ctx, double* wide_mac_per_s);
int32_t probe_fq_mul(kzg_ctx* ctx, double* fq_mul_per_s);

}  // namespace kzg
