// extern "C" boundary of libkzg_b200.so (include/kzg_b200.h).
//
// A context is either ONE CUDA device (kzg_ctx_create; every entry point forwards to the single-device engine,
// engine_one.cu) or a GROUP of devices driven by one host caller (kzg_ctx_create_multi) — the shape of the reference,
// which is a single process (src/main.rs:170, benches/polynomial_commitment.rs:38-47): a Rust Polynomial::commit
// linked against this library reaches every GPU of the node without torchrun, NCCL or a second process.
//
// Group layout:
//   * a setup is sharded by POINT RANGE (member g holds bases [off_g, off_g+1)) or REPLICATED (small setups: blob batches);
//   * one MSM  = every member's partial sum over its range, exchanged and folded by k_exchange_combine (msm_tail.cu)
//                over plain peer pointers (cudaDeviceEnablePeerAccess; NVLink 5 / NVSwitch) — the same kernel the
//                one-process-per-GPU mode runs over CUDA IPC mappings;
//   * a batch  = rows split by member, no exchange;  setup generation = index range per member, no exchange;
//   * evaluate / quotient of one polynomial: member 0 (replicas only).
// Every member has a host worker thread that issues its launches (a chain of ~45 launches costs the host ~0.1 ms: eight
// devices fed from one thread would be launch-bound at n = 2^20), so one caller drives all devices in parallel.
// Device pointers passed to a group may live on any member device: the kernels read them over NVLink.
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include "engine.cuh"

using namespace kzg;

namespace {

// ---- host worker: one per member >= 1 (member 0 runs on the caller's thread) ------------------------------------------
class Worker {
 public:
  explicit Worker(int device) : device_(device), th_([this] { run(); }) {}
  ~Worker() {
    { std::lock_guard<std::mutex> l(m_); stop_ = true; }
    cv_.notify_all();
    th_.join();
  }
  void post(std::function<void()> job) {
    { std::lock_guard<std::mutex> l(m_); job_ = std::move(job); has_job_ = true; done_ = false; }
    cv_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> l(m_);
    cv_done_.wait(l, [this] { return done_; });
  }

 private:
  void run() {
    cudaSetDevice(device_);
    std::unique_lock<std::mutex> l(m_);
    for (;;) {
      cv_.wait(l, [this] { return has_job_ || stop_; });
      if (stop_) return;
      std::function<void()> job = std::move(job_);
      has_job_ = false;
      l.unlock();
      job();
      l.lock();
      done_ = true;
      cv_done_.notify_all();
    }
  }
  int device_;
  std::mutex m_;
  std::condition_variable cv_, cv_done_;
  std::function<void()> job_;
  bool has_job_ = false, stop_ = false, done_ = true;
  std::thread th_;
};

struct GroupTicket {
  int32_t member_ticket[KZG_MAX_PEERS];
  bool busy = false;
};

struct GroupState {
  std::vector<std::unique_ptr<Worker>> workers;     // workers[g - 1] serves member g
  GroupTicket tickets[KZG_MAX_INFLIGHT];
};

inline bool is_group(const kzg_ctx* ctx) { return ctx && !ctx->members.empty(); }
inline GroupState* gs(kzg_ctx* head) { return reinterpret_cast<GroupState*>(head->group_state); }
inline int group_size(const kzg_ctx* ctx) { return (int)ctx->members.size(); }

// fn(g, member) on every member at once (member 0 on this thread); the first failure wins and its message moves to the head
int32_t for_members(kzg_ctx* head, const std::function<int32_t(int, kzg_ctx*)>& fn) {
  const int G = group_size(head);
  std::vector<int32_t> st((size_t)G, KZG_OK);
  GroupState* S = gs(head);
  for (int g = 1; g < G; g++) S->workers[(size_t)g - 1]->post([&, g] { st[(size_t)g] = fn(g, head->members[(size_t)g]); });
  st[0] = fn(0, head);
  for (int g = 1; g < G; g++) S->workers[(size_t)g - 1]->wait();
  for (int g = 0; g < G; g++) {
    if (st[(size_t)g] == KZG_OK) continue;
    if (g) head->err = "device " + std::to_string(head->members[(size_t)g]->device) + ": " + head->members[(size_t)g]->err;
    return st[(size_t)g];
  }
  return KZG_OK;
}

// contiguous ranges of n items over G members, remainders to the low members (parallel.shard_range on the Python side)
std::vector<uint64_t> shard_offsets(uint64_t n, int G) {
  std::vector<uint64_t> off((size_t)G + 1, 0);
  const uint64_t base = n / (uint64_t)G, rem = n % (uint64_t)G;
  for (int g = 0; g < G; g++) off[(size_t)g + 1] = off[(size_t)g] + base + ((uint64_t)g < rem ? 1 : 0);
  return off;
}

// part of [0, n) that falls into shard g of a range-sharded setup
inline uint64_t shard_share(const kzg_setup* S, int g, uint64_t n) {
  const uint64_t lo = S->shard_off[(size_t)g], hi = S->shard_off[(size_t)g + 1];
  return n <= lo ? 0 : std::min(n, hi) - lo;
}

int choose_layout(const kzg_ctx* head, uint64_t n) {
  if (head->opt_setup_layout == 1 || head->opt_setup_layout == 2) return head->opt_setup_layout;
  return n <= (1u << 16) ? 2 : 1;      // small setups (blob batches: 4096 points) are replicated, large ones sharded by range
}

kzg_setup* new_group_setup(kzg_ctx* head, uint64_t n, int layout) {
  kzg_setup* S = new kzg_setup();
  S->ctx = head; S->n = n; S->layout = layout;
  S->shards.assign((size_t)group_size(head), nullptr);
  S->shard_off = layout == 1 ? shard_offsets(n, group_size(head)) : std::vector<uint64_t>((size_t)group_size(head) + 1, 0);
  if (layout == 2) for (size_t g = 1; g < S->shard_off.size(); g++) S->shard_off[g] = n;   // every shard: all n items
  return S;
}

void destroy_group_setup(kzg_setup* S) {
  for (kzg_setup* sh : S->shards) if (sh) one_setup_destroy(sh);
  delete S;
}

bool setup_matches(const kzg_ctx* ctx, const kzg_setup* s) {
  return s && (is_group(ctx) ? (s->layout != 0 && s->ctx == ctx) : s->layout == 0);
}

#define KZG_CHECK_SETUP(ctx, s)                                                                                   \
  do {                                                                                                            \
    if (!setup_matches(ctx, s)) return fail(ctx, KZG_E_BAD_ARG, "this setup belongs to another kind of context"); \
  } while (0)

int32_t group_create(kzg_ctx** out, const int32_t* ids, int32_t n) {
  std::vector<kzg_ctx*> m;
  auto undo = [&m] { for (kzg_ctx* c : m) one_ctx_destroy(c); };
  for (int g = 0; g < n; g++) {
    kzg_ctx* c = nullptr;
    int32_t st = one_ctx_create(&c, ids[g]);
    if (st != KZG_OK) { undo(); return st; }
    m.push_back(c);
  }
  int prev = 0;
  cudaGetDevice(&prev);
  // every member reads and writes every other member's memory (mailboxes, scalars, results): plain peer access
  for (int a = 0; a < n; a++) {
    for (int b = 0; b < n; b++) {
      if (ids[a] == ids[b]) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, ids[a], ids[b]) != cudaSuccess || !can) {
        cudaGetLastError(); cudaSetDevice(prev); undo();
        return KZG_E_CUDA;          // no NVLink / PCIe peer path between two of the devices
      }
      cudaSetDevice(ids[a]);
      cudaError_t e = cudaDeviceEnablePeerAccess(ids[b], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); cudaSetDevice(prev); undo(); return KZG_E_CUDA; }
      cudaGetLastError();
    }
  }
  void* boxes[KZG_MAX_PEERS] = {};
  for (int g = 0; g < n; g++) {
    cudaSetDevice(ids[g]);
    int32_t st = p2p_alloc_box(m[(size_t)g], n, g);
    if (st != KZG_OK) { cudaSetDevice(prev); undo(); return st; }
    boxes[g] = m[(size_t)g]->p2p_box[g];
  }
  for (int g = 0; g < n; g++) p2p_attach_local(m[(size_t)g], boxes);
  cudaSetDevice(prev);
  kzg_ctx* head = m[0];
  GroupState* S = new GroupState();
  for (int g = 1; g < n; g++) S->workers.emplace_back(new Worker(ids[g]));
  head->group_state = S;
  head->members = m;
  for (int a = 0; a < n; a++) for (int b = a + 1; b < n; b++) if (ids[a] == ids[b]) head->group_shares_devices = true;
  for (kzg_ctx* c : m) c->group_head = head;
  *out = head;
  return KZG_OK;
}

void group_destroy(kzg_ctx* head) {
  for (kzg_ctx* c : head->members) one_ctx_sync(c);
  delete gs(head);                       // joins the workers
  head->group_state = nullptr;
  std::vector<kzg_ctx*> m = head->members;
  head->members.clear();
  for (size_t g = 1; g < m.size(); g++) one_ctx_destroy(m[g]);
  one_ctx_destroy(head);
}

// launch what the members still hold back for fusing
void group_flush(kzg_ctx* head) {
  bool any = false;
  for (kzg_ctx* m : head->members) any = any || m->open_group.count > 0;
  if (any) for_members(head, [](int, kzg_ctx* m) { return flush_submitted(m); });
}

// ---- one MSM on a group --------------------------------------------------------------------------------------------
// `scalars` is a host buffer (host == true) or a device pointer on any member device; *ticket = the head's ticket
int32_t group_msm_submit(kzg_ctx* head, const kzg_setup* S, const void* scalars, uint64_t n, bool host, int32_t* ticket) {
  if (n > S->n) return fail(head, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  const int G = group_size(head);
  const uint32_t F = queue_capacity(head);              // tickets that can be out at once: lanes x fused calls
  GroupTicket& T = gs(head)->tickets[head->next_ticket % F];
  if (T.busy) return fail(head, KZG_E_BAD_ARG, "every lane has a call in flight: wait for the oldest ticket first");
  for (int g = 0; g < G; g++) T.member_ticket[g] = -1;
  // a replicated setup, or a sum that lives entirely in member 0's range: that device alone, no exchange
  if (S->layout == 2 || n <= S->shard_off[1]) {
    KZG_TRY(submit_common(head, S->shards[0], scalars, n, host, false, &T.member_ticket[0]));
  } else {
    const uint8_t* base = reinterpret_cast<const uint8_t*>(scalars);
    // members that share a device: allocate everywhere first, launch afterwards (see reserve_call, engine_one.cu)
    if (head->group_shares_devices)
      KZG_TRY(for_members(head, [&](int g, kzg_ctx* m) {
        return reserve_call(m, S->shards[(size_t)g], base + S->shard_off[(size_t)g] * 32, shard_share(S, g, n), host);
      }));
    int32_t st = for_members(head, [&](int g, kzg_ctx* m) {
      return submit_common(m, S->shards[(size_t)g], base + S->shard_off[(size_t)g] * 32, shard_share(S, g, n), host, true,
                           &T.member_ticket[g]);
    });
    if (st != KZG_OK) {        // some members may have queued their part: drain them, the group is out of step for good
      uint8_t sink[48];
      for (int g = 0; g < G; g++) {
        if (T.member_ticket[g] >= 0) one_msm_g1_wait(head->members[(size_t)g], T.member_ticket[g], sink);
        head->members[(size_t)g]->p2p_failed = true;
      }
      return st;
    }
  }
  T.busy = true;
  *ticket = T.member_ticket[0];
  return KZG_OK;
}

int32_t group_msm_wait(kzg_ctx* head, int32_t ticket, uint8_t out_c48[48]) {
  if (ticket < 0) return KZG_E_BAD_ARG;
  GroupTicket& T = gs(head)->tickets[(uint32_t)ticket % queue_capacity(head)];
  if (!T.busy || T.member_ticket[0] != ticket) return fail(head, KZG_E_BAD_ARG, "no call in flight under this ticket");
  T.busy = false;
  // calls still staged for fusing (option msm_coalesce): EVERY member launches its part before anybody waits — a member's
  // exchange kernel spins until all the others have published
  group_flush(head);
  int32_t first_err = KZG_OK;
  uint8_t sink[48];
  for (int g = group_size(head) - 1; g >= 0; g--) {      // every member computed the fold; member 0's copy is the answer
    if (T.member_ticket[g] < 0) continue;
    int32_t st = one_msm_g1_wait(head->members[(size_t)g], T.member_ticket[g], g == 0 ? out_c48 : sink);
    if (st != KZG_OK && first_err == KZG_OK) {
      first_err = st;
      if (g) head->err = "device " + std::to_string(head->members[(size_t)g]->device) + ": " + head->members[(size_t)g]->err;
    }
  }
  return first_err;
}

int32_t group_msm(kzg_ctx* head, const kzg_setup* S, const void* scalars, uint64_t n, bool host, uint8_t out_c48[48]) {
  int32_t t = -1;
  KZG_TRY(group_msm_submit(head, S, scalars, n, host, &t));
  return group_msm_wait(head, t, out_c48);
}

// host-side trimming of Polynomial::try_from (src/polynomial.rs:63-72), as engine_one.cu does for one device
uint64_t trimmed_len_host(const uint8_t* coeffs, uint64_t n) {
  while (n > 0) {
    Fr a; fr_load_le(a, coeffs + (n - 1) * 32); fr_canon_reduce(a);
    if (!is_zero(a)) break;
    n--;
  }
  return n;
}

// rows of a batch over the members of a group: member g takes rows [off_g, off_g+1)
template <typename F>
int32_t group_rows(kzg_ctx* head, uint64_t blobs, F&& fn) {
  const std::vector<uint64_t> off = shard_offsets(blobs, group_size(head));
  return for_members(head, [&](int g, kzg_ctx* m) {
    const uint64_t r0 = off[(size_t)g], rows = off[(size_t)g + 1] - r0;
    return rows ? fn(g, m, r0, rows) : (int32_t)KZG_OK;
  });
}

inline const uint8_t* cu8(const void* p) { return reinterpret_cast<const uint8_t*>(p); }
inline uint8_t* u8(void* p) { return reinterpret_cast<uint8_t*>(p); }

}  // namespace

extern "C" {

// ---- context ---------------------------------------------------------------------------------------------------------
int32_t kzg_ctx_create(kzg_ctx** out, int32_t device_id) { return one_ctx_create(out, device_id); }

int32_t kzg_ctx_create_multi(kzg_ctx** out, const int32_t* device_ids, int32_t n_dev) {
  if (!out || !device_ids || n_dev < 1 || n_dev > KZG_MAX_PEERS) return KZG_E_BAD_ARG;
  *out = nullptr;
  if (n_dev == 1) return one_ctx_create(out, device_ids[0]);
  return group_create(out, device_ids, n_dev);
}

int32_t kzg_ctx_device_count(const kzg_ctx* ctx) { return !ctx ? 0 : (is_group(ctx) ? group_size(ctx) : 1); }

void kzg_ctx_destroy(kzg_ctx* ctx) {
  if (is_group(ctx)) group_destroy(ctx); else one_ctx_destroy(ctx);
}

const char* kzg_last_error(const kzg_ctx* ctx) { return one_last_error(ctx); }

void* kzg_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}

void kzg_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int32_t kzg_ctx_sync(kzg_ctx* ctx) {
  if (!ctx) return KZG_E_BAD_ARG;
  if (!is_group(ctx)) return one_ctx_sync(ctx);
  group_flush(ctx);
  for (kzg_ctx* m : ctx->members) {
    int32_t st = one_ctx_sync(m);
    if (st != KZG_OK) { if (m != ctx) ctx->err = m->err; return st; }
  }
  return KZG_OK;
}

void* kzg_ctx_stream(const kzg_ctx* ctx) { return one_ctx_stream(ctx); }

uint64_t kzg_ctx_launch_count(const kzg_ctx* ctx) {
  if (!ctx) return 0;
  if (!is_group(ctx)) return one_ctx_launch_count(ctx);
  uint64_t c = 0;
  for (const kzg_ctx* m : ctx->members) c += one_ctx_launch_count(m);
  return c;
}

float kzg_ctx_last_kernel_ms(const kzg_ctx* ctx, int32_t which) { return ctx ? one_ctx_last_kernel_ms(ctx, which) : 0.f; }

// group: the interval is timed on every member's own device and the longest one is reported
int32_t kzg_ctx_timer_start(kzg_ctx* ctx) {
  if (!ctx) return KZG_E_BAD_ARG;
  if (!is_group(ctx)) return one_ctx_timer_start(ctx);
  for (kzg_ctx* m : ctx->members) KZG_TRY(one_ctx_timer_start(m));
  return KZG_OK;
}

int32_t kzg_ctx_timer_stop(kzg_ctx* ctx, float* ms) {
  if (!ctx || !ms) return KZG_E_BAD_ARG;
  if (!is_group(ctx)) return one_ctx_timer_stop(ctx, ms);
  group_flush(ctx);
  float worst = 0.f;
  for (kzg_ctx* m : ctx->members) {
    float t = 0.f;
    int32_t st = one_ctx_timer_stop(m, &t);
    if (st != KZG_OK) { if (m != ctx) ctx->err = m->err; return st; }
    worst = std::max(worst, t);
  }
  *ms = worst;
  return KZG_OK;
}

int32_t kzg_ctx_set_option(kzg_ctx* ctx, const char* name, int64_t value) {
  if (!ctx) return KZG_E_BAD_ARG;
  if (!is_group(ctx)) return one_ctx_set_option(ctx, name, value);
  for (kzg_ctx* m : ctx->members) {
    int32_t st = one_ctx_set_option(m, name, value);
    if (st != KZG_OK) { if (m != ctx) ctx->err = m->err; return st; }
  }
  return KZG_OK;
}

// ---- trusted setup ---------------------------------------------------------------------------------------------------
int32_t kzg_setup_generate(kzg_ctx* ctx, const uint8_t secret_be[32], uint64_t first, uint64_t n, kzg_setup** out) {
  if (!ctx || !secret_be || !out) return KZG_E_BAD_ARG;
  if (!is_group(ctx)) return one_setup_generate(ctx, secret_be, first, n, out);
  kzg_setup* S = new_group_setup(ctx, n, choose_layout(ctx, n));
  S->first = first;
  int32_t st = for_members(ctx, [&](int g, kzg_ctx* m) {      // index-range sharding: no exchange
    const uint64_t lo = S->layout == 1 ? S->shard_off[(size_t)g] : 0;
    const uint64_t cnt = S->layout == 1 ? S->shard_off[(size_t)g + 1] - lo : n;
    return one_setup_generate(m, secret_be, first + lo, cnt, &S->shards[(size_t)g]);
  });
  if (st != KZG_OK) { destroy_group_setup(S); return st; }
  *out = S;
  return KZG_OK;
}

int32_t kzg_setup_import_g1(kzg_ctx* ctx, const uint8_t* g1_c48, uint64_t n, kzg_setup** out) {
  if (!ctx || (!g1_c48 && n) || !out) return KZG_E_BAD_ARG;
  if (!is_group(ctx)) return one_setup_import_g1(ctx, g1_c48, n, out);
  kzg_setup* S = new_group_setup(ctx, n, choose_layout(ctx, n));
  int32_t st = for_members(ctx, [&](int g, kzg_ctx* m) {
    const uint64_t lo = S->layout == 1 ? S->shard_off[(size_t)g] : 0;
    const uint64_t cnt = S->layout == 1 ? S->shard_off[(size_t)g + 1] - lo : n;
    int32_t r = one_setup_import_g1(m, g1_c48 + lo * 48, cnt, &S->shards[(size_t)g]);
    if (r == KZG_OK) S->shards[(size_t)g]->first = lo;        // position of the shard in the sequence (G2 powers follow it)
    return r;
  });
  if (st != KZG_OK) { destroy_group_setup(S); return st; }
  *out = S;
  return KZG_OK;
}

int32_t kzg_setup_export_g1(kzg_ctx* ctx, const kzg_setup* s, uint64_t first, uint64_t count, uint8_t* out_c48) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_setup_export_g1(ctx, s, first, count, out_c48);
  if (first > s->n || count > s->n - first) return fail(ctx, KZG_E_BAD_ARG, "export range outside the setup");
  if (s->layout == 2) return one_setup_export_g1(ctx, s->shards[0], first, count, out_c48);
  for (int g = 0; g < group_size(ctx); g++) {
    const uint64_t lo = std::max(first, s->shard_off[(size_t)g]), hi = std::min(first + count, s->shard_off[(size_t)g + 1]);
    if (lo >= hi) continue;
    kzg_ctx* m = ctx->members[(size_t)g];
    int32_t st = one_setup_export_g1(m, s->shards[(size_t)g], lo - s->shard_off[(size_t)g], hi - lo, out_c48 + (lo - first) * 48);
    if (st != KZG_OK) { if (g) ctx->err = m->err; return st; }
  }
  return KZG_OK;
}

uint64_t kzg_setup_len(const kzg_setup* s) { return s ? s->n : 0; }

int32_t kzg_setup_generate_g2(kzg_ctx* ctx, kzg_setup* s, const uint8_t secret_be[32], uint64_t count) {
  if (!ctx || !s || !secret_be) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_setup_generate_g2(ctx, s, secret_be, count);
  if (count > s->n) return fail(ctx, KZG_E_BAD_ARG, "more G2 items than setup items");
  if (s->layout == 2) return one_setup_generate_g2(ctx, s->shards[0], secret_be, count);
  return for_members(ctx, [&](int g, kzg_ctx* m) {
    return one_setup_generate_g2(m, s->shards[(size_t)g], secret_be, shard_share(s, g, count));
  });
}

int32_t kzg_setup_import_g2(kzg_ctx* ctx, kzg_setup* s, const uint8_t* g2_c96, uint64_t count) {
  if (!ctx || !s || (!g2_c96 && count)) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_setup_import_g2(ctx, s, g2_c96, count);
  if (count > s->n) return fail(ctx, KZG_E_BAD_ARG, "more G2 items than setup items");
  if (s->layout == 2) return one_setup_import_g2(ctx, s->shards[0], g2_c96, count);
  return for_members(ctx, [&](int g, kzg_ctx* m) {
    return one_setup_import_g2(m, s->shards[(size_t)g], g2_c96 + s->shard_off[(size_t)g] * 96, shard_share(s, g, count));
  });
}

uint64_t kzg_setup_g2_len(const kzg_setup* s) {
  if (!s) return 0;
  if (s->layout == 0) return one_setup_g2_len(s);
  if (s->layout == 2) return one_setup_g2_len(s->shards[0]);
  uint64_t c = 0;
  for (const kzg_setup* sh : s->shards) c += one_setup_g2_len(sh);
  return c;
}

int32_t kzg_setup_export_g2(kzg_ctx* ctx, const kzg_setup* s, uint64_t first, uint64_t count, uint8_t* out_c96) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_setup_export_g2(ctx, s, first, count, out_c96);
  if (s->layout == 2) return one_setup_export_g2(ctx, s->shards[0], first, count, out_c96);
  const uint64_t have = kzg_setup_g2_len(s);
  if (first > have || count > have - first) return fail(ctx, KZG_E_BAD_ARG, "export range outside the G2 items");
  for (int g = 0; g < group_size(ctx); g++) {
    const uint64_t lo = std::max(first, s->shard_off[(size_t)g]), hi = std::min(first + count, s->shard_off[(size_t)g + 1]);
    if (lo >= hi) continue;
    kzg_ctx* m = ctx->members[(size_t)g];
    int32_t st = one_setup_export_g2(m, s->shards[(size_t)g], lo - s->shard_off[(size_t)g], hi - lo, out_c96 + (lo - first) * 96);
    if (st != KZG_OK) { if (g) ctx->err = m->err; return st; }
  }
  return KZG_OK;
}

void kzg_setup_destroy(kzg_setup* s) {
  if (!s) return;
  if (s->layout == 0) one_setup_destroy(s); else destroy_group_setup(s);
}

int32_t kzg_setup_prepare_msm(kzg_ctx* ctx, kzg_setup* s, int32_t window_bits) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_setup_prepare_msm(ctx, s, window_bits);
  if (s->layout == 2) return one_setup_prepare_msm(ctx, s->shards[0], window_bits);     // single sums run on member 0
  return for_members(ctx, [&](int g, kzg_ctx* m) { return one_setup_prepare_msm(m, s->shards[(size_t)g], window_bits); });
}

int32_t kzg_setup_prepare_fixed(kzg_ctx* ctx, kzg_setup* s, uint64_t n_points, int32_t window_bits) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_setup_prepare_fixed(ctx, s, n_points, window_bits);
  if (s->layout != 2)
    return fail(ctx, KZG_E_BAD_ARG, "fixed-base tables serve blob batches and need a replicated setup (option setup_layout = 2)");
  return for_members(ctx, [&](int g, kzg_ctx* m) { return one_setup_prepare_fixed(m, s->shards[(size_t)g], n_points, window_bits); });
}

int32_t kzg_msm_plan(const kzg_ctx* ctx, const kzg_setup* s, uint64_t n, int32_t plan[5]) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  if (!setup_matches(ctx, s)) return KZG_E_BAD_ARG;
  if (!is_group(ctx)) return one_msm_plan(ctx, s, n, plan);
  return one_msm_plan(ctx, s->shards[0], s->layout == 1 ? shard_share(s, 0, n) : n, plan);      // member 0's share
}

int32_t kzg_setup_fixed_plan(const kzg_setup* s, int64_t plan[4]) {
  if (!s || !plan) return KZG_E_BAD_ARG;
  const kzg_setup* one = s->layout == 0 ? s : s->shards[0];
  if (!one || !one->fix_tab) return KZG_E_BAD_ARG;
  plan[0] = one->fix_c; plan[1] = one->fix_w; plan[2] = (int64_t)one->fix_n;
  plan[3] = (int64_t)fixed_table_bytes(one->fix_n, one->fix_c);
  return KZG_OK;
}

// ---- one polynomial ----------------------------------------------------------------------------------------------------
int32_t kzg_msm_g1(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_le32, uint64_t n, uint8_t out_c48[48]) {
  if (!ctx || !s || (!scalars_le32 && n) || !out_c48) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_msm_g1(ctx, s, scalars_le32, n, out_c48);
  return group_msm(ctx, s, scalars_le32, n, true, out_c48);
}

int32_t kzg_msm_g1_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, uint8_t out_c48[48]) {
  if (!ctx || !s || (!scalars_dev && n) || !out_c48) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_msm_g1_dev(ctx, s, scalars_dev, n, out_c48);
  return group_msm(ctx, s, scalars_dev, n, false, out_c48);
}

int32_t kzg_msm_g1_submit(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_le32, uint64_t n, int32_t* ticket) {
  if (!ctx || !s || (!scalars_le32 && n) || !ticket) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_msm_g1_submit(ctx, s, scalars_le32, n, ticket);
  return group_msm_submit(ctx, s, scalars_le32, n, true, ticket);
}

int32_t kzg_msm_g1_submit_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, int32_t* ticket) {
  if (!ctx || !s || (!scalars_dev && n) || !ticket) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_msm_g1_submit_dev(ctx, s, scalars_dev, n, ticket);
  return group_msm_submit(ctx, s, scalars_dev, n, false, ticket);
}

int32_t kzg_msm_g1_wait(kzg_ctx* ctx, int32_t ticket, uint8_t out_c48[48]) {
  if (!ctx || !out_c48) return KZG_E_BAD_ARG;
  if (!is_group(ctx)) return one_msm_g1_wait(ctx, ticket, out_c48);
  return group_msm_wait(ctx, ticket, out_c48);
}

int32_t kzg_commit(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t n, uint8_t out_c48[48]) {
  if (!ctx || !s || (!coeffs_le32 && n) || !out_c48) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_commit(ctx, s, coeffs_le32, n, out_c48);
  const uint64_t nt = trimmed_len_host(coeffs_le32, n);
  // (degree() + 1) > setup.len(), degree() == 0 for the empty / zero polynomial (src/polynomial.rs:95-100, 201)
  if (std::max<uint64_t>(nt, 1) > s->n) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  return group_msm(ctx, s, coeffs_le32, nt, true, out_c48);
}

int32_t kzg_evaluate(kzg_ctx* ctx, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32], uint8_t y_le[32]) {
  return one_evaluate(ctx, coeffs_le32, n, z_le, y_le);          // group: member 0 (one polynomial does not shard)
}

int32_t kzg_evaluate_dev(kzg_ctx* ctx, const void* coeffs_dev, uint64_t n, const uint8_t z_le[32], uint8_t y_le[32]) {
  return one_evaluate_dev(ctx, coeffs_dev, n, z_le, y_le);
}

int32_t kzg_quotient(kzg_ctx* ctx, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32], const uint8_t* y_le,
                     uint8_t* q_le32, uint64_t* q_len) {
  return one_quotient(ctx, coeffs_le32, n, z_le, y_le, q_le32, q_len);
}

int32_t kzg_open(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32],
                 const uint8_t* y_expected_le, uint8_t y_out_le[32], uint8_t proof_c48[48]) {
  if (!ctx || !s || (!coeffs_le32 && n) || !z_le || !proof_c48) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_open(ctx, s, coeffs_le32, n, z_le, y_expected_le, y_out_le, proof_c48);
  // the quotient scan on member 0, then the sum over every member's range: the members read their slice of the
  // quotient straight out of member 0's memory (it stays put until the wait below)
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(ctx->device);
  uint8_t* qd = nullptr;
  uint64_t ql = 0;
  int32_t st = quotient_core(ctx, coeffs_le32, n, z_le, y_expected_le, y_out_le, &qd, &ql);
  cudaSetDevice(prev);
  KZG_TRY(st);
  if (std::max<uint64_t>(ql, 1) > s->n) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  return group_msm(ctx, s, qd, ql, false, proof_c48);
}

// ---- one process per GPU: exchange over CUDA IPC mailboxes ----------------------------------------------------------------
#define KZG_NOT_ON_GROUP(ctx)                                                                                                  \
  do {                                                                                                                         \
    if (is_group(ctx))                                                                                                         \
      return fail(ctx, KZG_E_BAD_ARG, "a multi-device context shards internally: use the plain call (kzg_msm_g1, kzg_commit ...)"); \
  } while (0)

int32_t kzg_msm_g1_partial_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, void* partial_dev) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  KZG_NOT_ON_GROUP(ctx);
  KZG_CHECK_SETUP(ctx, s);
  return one_msm_g1_partial_dev(ctx, s, scalars_dev, n, partial_dev);
}
int32_t kzg_p2p_init(kzg_ctx* ctx, int32_t world, int32_t rank, uint8_t handle_out[KZG_P2P_HANDLE_BYTES]) {
  if (!ctx) return KZG_E_BAD_ARG;
  KZG_NOT_ON_GROUP(ctx);
  return one_p2p_init(ctx, world, rank, handle_out);
}
int32_t kzg_p2p_connect(kzg_ctx* ctx, const uint8_t* handles) {
  if (!ctx) return KZG_E_BAD_ARG;
  KZG_NOT_ON_GROUP(ctx);
  return one_p2p_connect(ctx, handles);
}
int32_t kzg_msm_g1_sharded_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, uint8_t out_c48[48]) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  KZG_NOT_ON_GROUP(ctx);
  KZG_CHECK_SETUP(ctx, s);
  return one_msm_g1_sharded_dev(ctx, s, scalars_dev, n, out_c48);
}
int32_t kzg_msm_g1_sharded_submit_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, int32_t* ticket) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  KZG_NOT_ON_GROUP(ctx);
  KZG_CHECK_SETUP(ctx, s);
  return one_msm_g1_sharded_submit_dev(ctx, s, scalars_dev, n, ticket);
}
int32_t kzg_msm_g1_sharded_submit(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_le32, uint64_t n, int32_t* ticket) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  KZG_NOT_ON_GROUP(ctx);
  KZG_CHECK_SETUP(ctx, s);
  return one_msm_g1_sharded_submit(ctx, s, scalars_le32, n, ticket);
}
int32_t kzg_g1_combine_dev(kzg_ctx* ctx, const void* partials_dev, uint64_t count, uint8_t out_c48[48]) {
  return one_g1_combine_dev(ctx, partials_dev, count, out_c48);
}

// ---- blob batches ------------------------------------------------------------------------------------------------------
// Group, replicated setup: rows split by member, no exchange.  Group, range-sharded setup (polynomials too long to
// replicate): row after row through the sharded sum.
int32_t kzg_commit_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t blobs, uint64_t n, uint8_t* out_c48) {
  if (!ctx || !s || (!coeffs_le32 && blobs * n) || !out_c48) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_commit_batch(ctx, s, coeffs_le32, blobs, n, out_c48);
  if (s->layout == 2)
    return group_rows(ctx, blobs, [&](int g, kzg_ctx* m, uint64_t r0, uint64_t rows) {
      return one_commit_batch(m, s->shards[(size_t)g], coeffs_le32 + r0 * n * 32, rows, n, out_c48 + r0 * 48);
    });
  for (uint64_t b = 0; b < blobs; b++) KZG_TRY(kzg_commit(ctx, s, coeffs_le32 + b * n * 32, n, out_c48 + b * 48));
  return KZG_OK;
}

int32_t kzg_open_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t blobs, uint64_t n,
                       const uint8_t* z_le32, uint8_t* y_le32, uint8_t* proofs_c48) {
  if (!ctx || !s || (!coeffs_le32 && blobs * n) || !z_le32 || !y_le32 || !proofs_c48) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_open_batch(ctx, s, coeffs_le32, blobs, n, z_le32, y_le32, proofs_c48);
  if (s->layout == 2)
    return group_rows(ctx, blobs, [&](int g, kzg_ctx* m, uint64_t r0, uint64_t rows) {
      return one_open_batch(m, s->shards[(size_t)g], coeffs_le32 + r0 * n * 32, rows, n, z_le32 + r0 * 32, y_le32 + r0 * 32,
                            proofs_c48 + r0 * 48);
    });
  for (uint64_t b = 0; b < blobs; b++)
    KZG_TRY(kzg_open(ctx, s, coeffs_le32 + b * n * 32, n, z_le32 + b * 32, nullptr, y_le32 + b * 32, proofs_c48 + b * 48));
  return KZG_OK;
}

int32_t kzg_commit_open_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t blobs, uint64_t n,
                              const uint8_t* z_le32, uint8_t* commits_c48, uint8_t* y_le32, uint8_t* proofs_c48) {
  if (!ctx || !s || (!coeffs_le32 && blobs * n) || !z_le32 || !commits_c48 || !y_le32 || !proofs_c48) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_commit_open_batch(ctx, s, coeffs_le32, blobs, n, z_le32, commits_c48, y_le32, proofs_c48);
  if (s->layout == 2)
    return group_rows(ctx, blobs, [&](int g, kzg_ctx* m, uint64_t r0, uint64_t rows) {
      return one_commit_open_batch(m, s->shards[(size_t)g], coeffs_le32 + r0 * n * 32, rows, n, z_le32 + r0 * 32,
                                   commits_c48 + r0 * 48, y_le32 + r0 * 32, proofs_c48 + r0 * 48);
    });
  KZG_TRY(kzg_commit_batch(ctx, s, coeffs_le32, blobs, n, commits_c48));
  return kzg_open_batch(ctx, s, coeffs_le32, blobs, n, z_le32, y_le32, proofs_c48);
}

// device-resident rows (anywhere the members can reach); results are queued, kzg_ctx_sync waits for every member
int32_t kzg_commit_batch_dev(kzg_ctx* ctx, const kzg_setup* s, const void* coeffs_dev, uint64_t blobs, uint64_t n, void* out_c48_dev) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_commit_batch_dev(ctx, s, coeffs_dev, blobs, n, out_c48_dev);
  if (s->layout != 2) return fail(ctx, KZG_E_BAD_ARG, "device-resident batches on a multi-device context need a replicated setup");
  return group_rows(ctx, blobs, [&](int g, kzg_ctx* m, uint64_t r0, uint64_t rows) {
    return one_commit_batch_dev(m, s->shards[(size_t)g], cu8(coeffs_dev) + r0 * n * 32, rows, n, u8(out_c48_dev) + r0 * 48);
  });
}

int32_t kzg_open_batch_dev(kzg_ctx* ctx, const kzg_setup* s, const void* coeffs_dev, uint64_t blobs, uint64_t n, const void* z_dev,
                           void* y_dev, void* proofs_c48_dev) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_open_batch_dev(ctx, s, coeffs_dev, blobs, n, z_dev, y_dev, proofs_c48_dev);
  if (s->layout != 2) return fail(ctx, KZG_E_BAD_ARG, "device-resident batches on a multi-device context need a replicated setup");
  return group_rows(ctx, blobs, [&](int g, kzg_ctx* m, uint64_t r0, uint64_t rows) {
    return one_open_batch_dev(m, s->shards[(size_t)g], cu8(coeffs_dev) + r0 * n * 32, rows, n, cu8(z_dev) + r0 * 32,
                              u8(y_dev) + r0 * 32, u8(proofs_c48_dev) + r0 * 48);
  });
}

int32_t kzg_commit_open_batch_dev(kzg_ctx* ctx, const kzg_setup* s, const void* coeffs_dev, uint64_t blobs, uint64_t n,
                                  const void* z_dev, void* commits_c48_dev, void* y_dev, void* proofs_c48_dev) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  KZG_CHECK_SETUP(ctx, s);
  if (!is_group(ctx)) return one_commit_open_batch_dev(ctx, s, coeffs_dev, blobs, n, z_dev, commits_c48_dev, y_dev, proofs_c48_dev);
  if (s->layout != 2) return fail(ctx, KZG_E_BAD_ARG, "device-resident batches on a multi-device context need a replicated setup");
  return group_rows(ctx, blobs, [&](int g, kzg_ctx* m, uint64_t r0, uint64_t rows) {
    return one_commit_open_batch_dev(m, s->shards[(size_t)g], cu8(coeffs_dev) + r0 * n * 32, rows, n, cu8(z_dev) + r0 * 32,
                                     u8(commits_c48_dev) + r0 * 48, u8(y_dev) + r0 * 32, u8(proofs_c48_dev) + r0 * 48);
  });
}

// ---- diagnostics (member 0 of a group) -----------------------------------------------------------------------------------
int32_t kzg_probe_imad_peak(kzg_ctx* ctx, double* wide_mac_per_s) { return one_probe_imad_peak(ctx, wide_mac_per_s); }
int32_t kzg_probe_fq_mul(kzg_ctx* ctx, double* fq_mul_per_s) { return one_probe_fq_mul(ctx, fq_mul_per_s); }

}  // extern "C"
