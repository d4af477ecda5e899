// Single-device engine behind the extern "C" boundary (abi.cu dispatches here, or to the multi-device group layer in
// abi.cu itself): context, setup objects and the bulk operations of ONE CUDA device.  Semantics (trimming, error
// cases) mirror the reference's public API:
//   Polynomial::try_from / degree / commit / evaluate / sub / divide_by_root  src/polynomial.rs:38-215
//   Evaluation::generate_proof                                                 src/polynomial.rs:260-269
//   SetupArtifactsGenerator                                                    src/trusted_setup.rs:20-78
// There is no CPU fallback for any bulk operation.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "engine.cuh"

using namespace kzg;

namespace kzg {

int32_t scratch_get(kzg_ctx* ctx, int slot, size_t bytes, void** out) {
  if (bytes == 0) bytes = 16;
  if (ctx->scratch_cap[slot] < bytes) {
    if (ctx->scratch[slot]) {
      KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->scratch[slot]);
      ctx->scratch[slot] = nullptr; ctx->scratch_cap[slot] = 0;
    }
    size_t cap = (bytes + (1u << 20) - 1) & ~((size_t)(1u << 20) - 1);
    cudaError_t e = cudaMalloc(&ctx->scratch[slot], cap);
    if (e != cudaSuccess) {
      cudaGetLastError();
      ctx->err = "cudaMalloc(scratch): " + std::string(cudaGetErrorString(e));
      return KZG_E_NOMEM;
    }
    ctx->scratch_cap[slot] = cap;
  }
  *out = ctx->scratch[slot];
  return KZG_OK;
}

int32_t pinned_get(kzg_ctx* ctx, size_t bytes, void** out) {
  if (ctx->pinned_cap < bytes) {
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr; ctx->pinned_cap = 0;
    KZG_CUDA_OK(ctx, cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_cap = bytes;
  }
  *out = ctx->pinned;
  return KZG_OK;
}

}  // namespace kzg

namespace {

struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) { cudaSetDevice(dev); changed = true; } }
  ~DeviceGuard() { if (changed && prev >= 0) cudaSetDevice(prev); }
};

// value == 0 mod r  <=>  value in {0, r, 2r} (2^256 < 3r)
bool le32_is_zero_mod_r(const uint8_t* b) {
  Fr a; fr_load_le(a, b); fr_canon_reduce(a);
  return is_zero(a);
}
// Polynomial::try_from trimming (src/polynomial.rs:63-72): number of coefficients up to the last
// non-zero one; 0 when every coefficient is zero (the reference keeps one zero coefficient then,
// which evaluates / commits identically to none).
uint64_t trimmed_len(const uint8_t* coeffs, uint64_t n) {
  while (n > 0 && le32_is_zero_mod_r(coeffs + (n - 1) * 32)) n--;
  return n;
}

int32_t ensure_msm_table(kzg_ctx* ctx, const kzg_setup* s) {
  if (s->msm_tab) return KZG_OK;
  return build_msm_table(ctx, const_cast<kzg_setup*>(s), 0);
}

// one MSM over the first n bases with device scalars -> XYZZ on device
int32_t msm_any(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_dev, uint64_t n, G1Xyzz* out_dev) {
  if (n == 0) return set_infinity(ctx, out_dev, 1);      // the empty sum: no table is touched (or built)
  if (s->fix_tab && n <= s->fix_n) return msm_fixed_batch(ctx, s, scalars_dev, 1, n, n, out_dev);
  KZG_TRY(ensure_msm_table(ctx, s));
  return msm_bucket(ctx, s, scalars_dev, n, out_dev);
}

int32_t finish_one(kzg_ctx* ctx, const G1Xyzz* pt_dev, uint8_t out_c48[48]) {
  void* o;
  KZG_TRY(scratch_get(ctx, SCR_OUT, 256, &o));
  KZG_TRY(finalize_points(ctx, pt_dev, 1, reinterpret_cast<uint8_t*>(o)));
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(out_c48, o, 48, cudaMemcpyDeviceToHost, ctx->stream));
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return KZG_OK;
}

// phase timers of the last bulk call: the events live on the lane the call ran on (lane 0 = the context itself)
void update_timers(kzg_ctx* ctx) {
  const kzg_ctx* src = (ctx->last_lane > 0 && ctx->lane[ctx->last_lane]) ? ctx->lane[ctx->last_lane] : ctx;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, src->ev[0], src->ev[1]) == cudaSuccess) ctx->last_ms[0] = ms;
  if (cudaEventElapsedTime(&ms, src->ev[2], src->ev[3]) == cudaSuccess) ctx->last_ms[1] = ms;
  if (cudaEventElapsedTime(&ms, src->ev_k[0], src->ev_k[1]) == cudaSuccess) ctx->last_ms[2] = ms;
  cudaGetLastError();
}

}  // namespace

namespace kzg {

// A context owns up to 2 x lanes streams, and the kernels of the peer exchange SPIN until other streams (of other devices,
// or of another member on the same device) have made progress.  With more streams than hardware queues
// (CUDA_DEVICE_MAX_CONNECTIONS, default 8) the driver multiplexes streams onto queues, and a spinning kernel then blocks
// unrelated streams that happen to share its queue — across devices that can close a cycle.  Ask for 32 queues unless
// the host decided otherwise; it only takes effect if CUDA is not initialised yet in this process.
static void want_hardware_queues() {
  static bool done = false;
  if (done) return;
  done = true;
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
}

int32_t one_ctx_create(kzg_ctx** out, int32_t device_id) {
  if (!out) return KZG_E_BAD_ARG;
  *out = nullptr;
  want_hardware_queues();
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device_id < 0 || device_id >= count) {
    cudaGetLastError();
    return KZG_E_CUDA;  // no CPU fallback: the engine needs a CUDA device
  }
  kzg_ctx* ctx = new kzg_ctx();
  ctx->device = device_id;
  if (cudaSetDevice(device_id) != cudaSuccess) { delete ctx; return KZG_E_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return KZG_E_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return KZG_E_CUDA; }
  for (int i = 0; i < 8; i++) cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming);
  for (int i = 0; i < 4; i++) cudaEventCreate(&ctx->ev[i]);
  for (int i = 0; i < 4; i++) cudaEventRecord(ctx->ev[i], ctx->stream);
  for (int i = 0; i < 2; i++) { cudaEventCreate(&ctx->ev_t[i]); cudaEventRecord(ctx->ev_t[i], ctx->stream); }
  for (int i = 0; i < 2; i++) { cudaEventCreate(&ctx->ev_k[i]); cudaEventRecord(ctx->ev_k[i], ctx->stream); }
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device_id) == cudaSuccess && v > 0) ctx->sm_count = v;
  // clock64() ticks at the SM clock: the peer-exchange timeout is stated in ms and converted with the device's own rate
  if (cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device_id) == cudaSuccess && v > 0) ctx->sm_clock_khz = v;
  ctx->p2p_timeout_cycles = 20000ll * ctx->sm_clock_khz;      // 20 s: above any lazy table build or first-call skew
  cudaGetLastError();
  cudaStreamSynchronize(ctx->stream);
  *out = ctx;
  return KZG_OK;
}

void one_ctx_destroy(kzg_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < SCR_COUNT; i++) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (auto& q : ctx->pending) {
    if (q.scalars) cudaFree(q.scalars);
    if (q.result) cudaFree(q.result);
    if (q.out_host) cudaFreeHost(q.out_host);
    if (q.copied) cudaEventDestroy(q.copied);
    if (q.done) cudaEventDestroy(q.done);
  }
  for (int r = 0; r < ctx->p2p_world; r++) {
    if (!ctx->p2p_box[r]) continue;
    if (r == ctx->p2p_rank) cudaFree(ctx->p2p_box[r]);
    else if (ctx->p2p_ipc) cudaIpcCloseMemHandle(ctx->p2p_box[r]);      // plain peer pointers belong to their own context
  }
  for (int l = 1; l < KZG_MAX_LANES; l++) if (ctx->lane[l]) one_ctx_destroy(ctx->lane[l]);
  if (ctx->gtable) cudaFree(ctx->gtable);
  if (ctx->g2table) cudaFree(ctx->g2table);
  for (int i = 0; i < 4; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 2; i++) if (ctx->ev_t[i]) cudaEventDestroy(ctx->ev_t[i]);
  for (int i = 0; i < 2; i++) if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]);
  for (int i = 0; i < 8; i++) if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
  for (int i = 0; i < 2; i++) if (ctx->side_ev[i]) cudaEventDestroy(ctx->side_ev[i]);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* one_last_error(const kzg_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int32_t one_ctx_sync(kzg_ctx* ctx) {
  DeviceGuard g(ctx->device);
  flush_submitted(ctx);
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  for (int l = 1; l < KZG_MAX_LANES; l++) if (ctx->lane[l]) KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->lane[l]->stream));
  update_timers(ctx);
  return KZG_OK;
}

int32_t one_ctx_timer_start(kzg_ctx* ctx) {
  DeviceGuard g(ctx->device);
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev_t[0], ctx->stream));
  return KZG_OK;
}

int32_t one_ctx_timer_stop(kzg_ctx* ctx, float* ms) {
  DeviceGuard g(ctx->device);
  flush_submitted(ctx);
  for (int l = 1; l < KZG_MAX_LANES; l++) {     // the interval ends when EVERY lane is done
    if (!ctx->lane[l]) continue;
    KZG_CUDA_OK(ctx, cudaEventRecord(ctx->lane[l]->ev_t[1], ctx->lane[l]->stream));
    KZG_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lane[l]->ev_t[1], 0));
  }
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev_t[1], ctx->stream));
  KZG_CUDA_OK(ctx, cudaEventSynchronize(ctx->ev_t[1]));
  KZG_CUDA_OK(ctx, cudaEventElapsedTime(ms, ctx->ev_t[0], ctx->ev_t[1]));
  return KZG_OK;
}

uint64_t one_ctx_launch_count(const kzg_ctx* ctx) {
  uint64_t c = ctx->launches;
  for (int l = 1; l < KZG_MAX_LANES; l++) if (ctx->lane[l]) c += ctx->lane[l]->launches;
  return c;
}

void* one_ctx_stream(const kzg_ctx* ctx) { return ctx ? reinterpret_cast<void*>(ctx->stream) : nullptr; }

float one_ctx_last_kernel_ms(const kzg_ctx* ctx, int32_t which) {
  if (ctx->timers_stale) {
    kzg_ctx* c = const_cast<kzg_ctx*>(ctx);
    DeviceGuard g(c->device);
    update_timers(c);
    c->timers_stale = false;
  }
  return (which >= 0 && which <= 2) ? ctx->last_ms[which] : 0.f;
}

int32_t one_ctx_set_option(kzg_ctx* ctx, const char* name, int64_t value) {
  std::string n(name ? name : "");
  if (n == "msm_window_bits") ctx->opt_msm_window = (int)value;
  else if (n == "fixed_window_bits") ctx->opt_fixed_window = (int)value;
  else if (n == "msm_sort") ctx->opt_msm_sort = (int)value;
  else if (n == "acc_order") ctx->opt_acc_order = (int)value;
  else if (n == "tree_lean") ctx->opt_tree_lean = (int)value;
  else if (n == "tree_lean_r1") ctx->opt_tree_lean_r1 = (int)value;
  else if (n == "reduce_chunk") ctx->opt_reduce_chunk = (int)value;
  else if (n == "acc_split") ctx->opt_acc_split = (int)value;
  else if (n == "acc_min_entries") ctx->opt_acc_min_entries = (int)std::max<int64_t>(1, value);
  else if (n == "coop_tail") ctx->opt_coop_tail = (int)value;
  else if (n == "reduce_lk") ctx->opt_reduce_lk = (int)value;
  else if (n == "msm_l2") ctx->opt_msm_l2 = (int)value;
  else if (n == "tree_r1_ool") ctx->opt_tree_r1_ool = (int)value;
  else if (n == "den_side") ctx->opt_den_side = (int)std::min<int64_t>(8, std::max<int64_t>(0, value));
  else if (n == "l2_rounds") ctx->opt_l2_rounds = (int)value;
  else if (n == "tree_k") ctx->opt_tree_k = (int)value;
  else if (n == "p2p_timeout_ms") ctx->p2p_timeout_cycles = std::max<int64_t>(1, value) * (long long)ctx->sm_clock_khz;
  else if (n == "fixed_table_max_bytes") ctx->opt_fixed_table_max_bytes = std::max<int64_t>(0, value);
  else if (n == "setup_layout") ctx->opt_setup_layout = (int)value;   // multi-device contexts: 0 auto, 1 point range, 2 replicated
  else if (n == "acc_minblocks") ctx->opt_acc_minblocks = (int)value;
  else if (n == "inv_top_fan") ctx->opt_inv_top_fan = (int)value;
  else if (n == "inv_top") ctx->opt_inv_top = std::max<int64_t>(8, value);
  else if (n == "msm_tree_rounds") ctx->opt_msm_tree_rounds = (int)value;
  else if (n == "tree_min_rows") ctx->opt_tree_min_rows = (int)value;
  else if (n == "r1_row_minor") ctx->opt_r1_row_minor = (int)value;
  else if (n == "tree_chunk_rows") ctx->opt_tree_chunk_rows = (int)value;
  else if (n == "tree_min_pairs") ctx->opt_tree_min_pairs = value;
  else if (n == "setup_tree_min") ctx->opt_setup_tree_min = (int)value;
  else if (n == "lanes") {
    if (ctx->lanes_in_use) return fail(ctx, KZG_E_BAD_ARG, "the number of lanes is fixed once a call has been submitted");
    ctx->opt_lanes = (int)std::min<int64_t>(KZG_MAX_LANES, std::max<int64_t>(1, value));
  }
  else if (n == "msm_coalesce") {
    if (ctx->lanes_in_use) return fail(ctx, KZG_E_BAD_ARG, "the number of fused calls is fixed once a call has been submitted");
    ctx->opt_msm_coalesce = (int)std::min<int64_t>(KZG_MAX_COALESCE, std::max<int64_t>(1, value));
  }
  else if (n == "heavy_cap") ctx->opt_heavy_cap = (int)std::max<int64_t>(1, value);
  else return fail(ctx, KZG_E_BAD_ARG, "unknown option");
  return KZG_OK;
}

// ---- setup ---------------------------------------------------------------------------------
static int32_t setup_alloc(kzg_ctx* ctx, uint64_t n, kzg_setup** out) {
  kzg_setup* s = new kzg_setup();
  s->ctx = ctx; s->n = n;
  if (n) {
    cudaError_t e = cudaMalloc(&s->bases, n * sizeof(G1Affine));
    if (e != cudaSuccess) { cudaGetLastError(); delete s; return fail(ctx, KZG_E_NOMEM, "cudaMalloc(bases)"); }
  }
  *out = s;
  return KZG_OK;
}

int32_t one_setup_generate(kzg_ctx* ctx, const uint8_t secret_be[32], uint64_t first, uint64_t n, kzg_setup** out) {
  if (!ctx || !secret_be || !out) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  kzg_setup* s = nullptr;
  KZG_TRY(setup_alloc(ctx, n, &s));
  s->first = first;
  int32_t st = setup_generate_dev(ctx, secret_be, first, n, s->bases);
  if (st == KZG_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, KZG_E_CUDA, "sync after setup generation");
  if (st != KZG_OK) { one_setup_destroy(s); return st; }
  *out = s;
  return KZG_OK;
}

int32_t one_setup_import_g1(kzg_ctx* ctx, const uint8_t* g1_c48, uint64_t n, kzg_setup** out) {
  if (!ctx || (!g1_c48 && n) || !out) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  kzg_setup* s = nullptr;
  KZG_TRY(setup_alloc(ctx, n, &s));
  int32_t st = KZG_OK;
  if (n) {
    void* in;
    st = scratch_get(ctx, SCR_SCALARS, n * 48, &in);
    if (st == KZG_OK && cudaMemcpyAsync(in, g1_c48, n * 48, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
      st = fail(ctx, KZG_E_CUDA, "H2D compressed points");
    int bad = -1;
    if (st == KZG_OK) st = uncompress_points(ctx, reinterpret_cast<uint8_t*>(in), n, s->bases, &bad);
    if (st == KZG_OK && bad >= 0) {
      ctx->err = "Got error while uncompressing: point " + std::to_string(bad);
      st = KZG_E_BAD_POINT;
    }
  }
  if (st != KZG_OK) { one_setup_destroy(s); return st; }
  *out = s;
  return KZG_OK;
}

int32_t one_setup_export_g1(kzg_ctx* ctx, const kzg_setup* s, uint64_t first, uint64_t count, uint8_t* out_c48) {
  if (!ctx || !s || first > s->n || count > s->n - first) return fail(ctx, KZG_E_BAD_ARG, "export range outside the setup");
  if (count == 0) return KZG_OK;
  DeviceGuard g(ctx->device);
  void* o;
  KZG_TRY(scratch_get(ctx, SCR_OUT, count * 48, &o));
  KZG_TRY(compress_points(ctx, s->bases + first, count, reinterpret_cast<uint8_t*>(o)));
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(out_c48, o, count * 48, cudaMemcpyDeviceToHost, ctx->stream));
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return KZG_OK;
}

uint64_t one_setup_len(const kzg_setup* s) { return s ? s->n : 0; }

// ---- G2 half of SetupArtifact (src/trusted_setup.rs:31-35, 64-72) ----------------------------
static int32_t setup_g2_alloc(kzg_ctx* ctx, kzg_setup* s, uint64_t count) {
  if (s->g2_c96 && s->g2_n == count) return KZG_OK;          // regenerated in place: no free / malloc (both synchronise the device)
  if (s->g2_c96) { cudaFree(s->g2_c96); s->g2_c96 = nullptr; s->g2_n = 0; }
  if (count == 0) return KZG_OK;
  if (cudaMalloc(&s->g2_c96, count * 96) != cudaSuccess) { cudaGetLastError(); return fail(ctx, KZG_E_NOMEM, "cudaMalloc(g2)"); }
  s->g2_n = count;
  return KZG_OK;
}

int32_t one_setup_generate_g2(kzg_ctx* ctx, kzg_setup* s, const uint8_t secret_be[32], uint64_t count) {
  if (!ctx || !s || !secret_be) return KZG_E_BAD_ARG;
  if (count > s->n) return fail(ctx, KZG_E_BAD_ARG, "more G2 items than setup items");
  DeviceGuard g(ctx->device);
  KZG_TRY(setup_g2_alloc(ctx, s, count));
  KZG_TRY(setup_generate_g2_dev(ctx, secret_be, s->first, count, s->g2_c96));
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return KZG_OK;
}

int32_t one_setup_import_g2(kzg_ctx* ctx, kzg_setup* s, const uint8_t* g2_c96, uint64_t count) {
  if (!ctx || !s || (!g2_c96 && count)) return KZG_E_BAD_ARG;
  if (count > s->n) return fail(ctx, KZG_E_BAD_ARG, "more G2 items than setup items");
  // Deserialize for G2Point (src/curves.rs:273-343): length fixed by the ABI, blst_p2_uncompress semantics
  // (flags, coordinates < q, Fq2 square root, on-curve; no subgroup check).  A handful of points (the CLI's setup.json)
  // is validated on the host; larger sets on the device, one thread per point (a host square root costs ~0.1 ms).
  DeviceGuard g(ctx->device);
  const bool on_device = count >= 64;
  if (!on_device) {
    for (uint64_t i = 0; i < count; i++) {
      if (kzg_g2_check(g2_c96 + 96 * i) != KZG_OK) {
        ctx->err = "Got error while uncompressing: G2 point " + std::to_string(i);
        return KZG_E_BAD_POINT;
      }
    }
    KZG_TRY(setup_g2_alloc(ctx, s, count));
    if (count) {
      KZG_CUDA_OK(ctx, cudaMemcpyAsync(s->g2_c96, g2_c96, count * 96, cudaMemcpyHostToDevice, ctx->stream));
      KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return KZG_OK;
  }
  void* stage;                                   // validated in scratch first: a rejected import leaves the setup untouched
  KZG_TRY(scratch_get(ctx, SCR_OUT, count * 96, &stage));
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(stage, g2_c96, count * 96, cudaMemcpyHostToDevice, ctx->stream));
  int64_t bad = -1;
  KZG_TRY(validate_g2_points(ctx, reinterpret_cast<const uint8_t*>(stage), count, &bad));
  if (bad >= 0) {
    ctx->err = "Got error while uncompressing: G2 point " + std::to_string(bad);
    return KZG_E_BAD_POINT;
  }
  KZG_TRY(setup_g2_alloc(ctx, s, count));
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(s->g2_c96, stage, count * 96, cudaMemcpyDeviceToDevice, ctx->stream));
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return KZG_OK;
}

int32_t one_setup_export_g2(kzg_ctx* ctx, const kzg_setup* s, uint64_t first, uint64_t count, uint8_t* out_c96) {
  if (!ctx || !s || first > s->g2_n || count > s->g2_n - first) return fail(ctx, KZG_E_BAD_ARG, "export range outside the G2 items");
  if (count == 0) return KZG_OK;
  DeviceGuard g(ctx->device);
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(out_c96, s->g2_c96 + first * 96, count * 96, cudaMemcpyDeviceToHost, ctx->stream));
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return KZG_OK;
}

uint64_t one_setup_g2_len(const kzg_setup* s) { return s ? s->g2_n : 0; }

int32_t one_msm_plan(const kzg_ctx* ctx, const kzg_setup* s, uint64_t n, int32_t plan[5]) {
  if (!ctx || !s || !plan || !s->msm_c) return KZG_E_BAD_ARG;
  plan[0] = s->msm_c; plan[1] = s->msm_w; plan[2] = 1 << (s->msm_c - 1);
  plan[3] = ctx->opt_msm_tree_rounds >= 0 ? ctx->opt_msm_tree_rounds : auto_tree_rounds((uint64_t)s->msm_w * n);   // of a call on its own, first level
  plan[4] = (int32_t)MSM_TAB_PITCH;
  return KZG_OK;
}

void one_setup_destroy(kzg_setup* s) {
  if (!s) return;
  DeviceGuard g(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  if (s->bases) cudaFree(s->bases);
  if (s->msm_tab) cudaFree(s->msm_tab);
  if (s->fix_tab) cudaFree(s->fix_tab);
  if (s->g2_c96) cudaFree(s->g2_c96);
  delete s;
}

int32_t one_setup_prepare_msm(kzg_ctx* ctx, kzg_setup* s, int32_t window_bits) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  return build_msm_table(ctx, s, window_bits);
}

int32_t one_setup_prepare_fixed(kzg_ctx* ctx, kzg_setup* s, uint64_t n_points, int32_t window_bits) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  return build_fixed_table(ctx, s, n_points, window_bits);
}

// ---- single-polynomial operations ----------------------------------------------------------------
int32_t one_msm_g1_partial_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, void* partial_dev) {
  if (!ctx || !s || !partial_dev) return KZG_E_BAD_ARG;
  if (n > s->n) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  DeviceGuard g(ctx->device);
  return msm_any(ctx, s, reinterpret_cast<const uint8_t*>(scalars_dev), n, reinterpret_cast<G1Xyzz*>(partial_dev));
}

int32_t one_msm_g1_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, uint8_t out_c48[48]) {
  if (!ctx || !s || !out_c48) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  void* r;
  KZG_TRY(scratch_get(ctx, SCR_MISC, 1024, &r));
  G1Xyzz* res = reinterpret_cast<G1Xyzz*>(r);
  KZG_TRY(one_msm_g1_partial_dev(ctx, s, scalars_dev, n, res));
  KZG_TRY(finish_one(ctx, res, out_c48));
  ctx->last_lane = 0;
  update_timers(ctx);
  return KZG_OK;
}

int32_t one_msm_g1(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_le32, uint64_t n, uint8_t out_c48[48]) {
  if (!ctx || !s || (!scalars_le32 && n) || !out_c48) return KZG_E_BAD_ARG;
  if (n > s->n) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  DeviceGuard g(ctx->device);
  void* d;
  KZG_TRY(scratch_get(ctx, SCR_SCALARS, n * 32, &d));
  if (n) KZG_CUDA_OK(ctx, cudaMemcpyAsync(d, scalars_le32, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  return one_msm_g1_dev(ctx, s, d, n, out_c48);
}

// ---- several calls in flight ----------------------------------------------------------------------------------------
// Ticket t runs on lane t mod L (L = option "lanes", default and at most KZG_MAX_LANES = 4): lane 0 is the context itself,
// the others are hidden sibling contexts of the same device (own streams, own scratch).  The host->device copy and the
// bulk kernels of one call overlap the latency-bound phases of the others (the inversion chain of every halving round,
// the bucket reduction: a few hundred warps for ~0.6 ms); two lanes are worth 13 % at n = 2^20 and 30 % at n = 2^17, four 18 % and 44 %.
void copy_options(kzg_ctx* dst, const kzg_ctx* src) {
  dst->opt_msm_window = src->opt_msm_window; dst->opt_fixed_window = src->opt_fixed_window;
  dst->opt_reduce_chunk = src->opt_reduce_chunk; dst->opt_msm_sort = src->opt_msm_sort; dst->opt_acc_order = src->opt_acc_order; dst->opt_tree_lean = src->opt_tree_lean; dst->opt_tree_lean_r1 = src->opt_tree_lean_r1; dst->opt_heavy_cap = src->opt_heavy_cap; dst->opt_acc_split = src->opt_acc_split;
  dst->opt_tree_min_rows = src->opt_tree_min_rows; dst->opt_r1_row_minor = src->opt_r1_row_minor; dst->opt_tree_chunk_rows = src->opt_tree_chunk_rows;
  dst->opt_tree_min_pairs = src->opt_tree_min_pairs; dst->opt_inv_top_fan = src->opt_inv_top_fan; dst->opt_inv_top = src->opt_inv_top;
  dst->opt_msm_tree_rounds = src->opt_msm_tree_rounds; dst->opt_acc_min_entries = src->opt_acc_min_entries;
  dst->opt_fixed_table_max_bytes = src->opt_fixed_table_max_bytes; dst->opt_setup_layout = src->opt_setup_layout;
  dst->opt_setup_tree_min = src->opt_setup_tree_min; dst->opt_tree_k = src->opt_tree_k; dst->opt_coop_tail = src->opt_coop_tail; dst->opt_reduce_lk = src->opt_reduce_lk; dst->opt_msm_l2 = src->opt_msm_l2; dst->opt_tree_r1_ool = src->opt_tree_r1_ool; dst->opt_den_side = src->opt_den_side; dst->opt_l2_rounds = src->opt_l2_rounds; dst->opt_acc_minblocks = src->opt_acc_minblocks;
}

static void latch_queue(kzg_ctx* ctx) {
  if (ctx->lanes_in_use) return;
  ctx->lanes_in_use = std::min(KZG_MAX_LANES, std::max(1, ctx->opt_lanes));
  ctx->coalesce_in_use = std::min(KZG_MAX_COALESCE, std::max(1, ctx->opt_msm_coalesce));
  // members of a group that share a device allocate in a phase of their own before anything is launched (reserve_call):
  // a deferred launch would allocate later, behind a peer's spinning exchange kernel
  if (ctx->group_head && ctx->group_head->group_shares_devices) ctx->coalesce_in_use = 1;
}
static inline uint32_t queue_slots(const kzg_ctx* ctx) { return (uint32_t)(ctx->lanes_in_use * ctx->coalesce_in_use); }

// lane l of the context (l = 0: the context itself), created on first use
static int32_t lane_at(kzg_ctx* ctx, int l, kzg_ctx** lane) {
  *lane = ctx;
  if (l == 0) return KZG_OK;
  if (!ctx->lane[l]) {
    int32_t st = one_ctx_create(&ctx->lane[l], ctx->device);
    if (st != KZG_OK) return fail(ctx, st, "could not create another lane of the submission queue");
  }
  copy_options(ctx->lane[l], ctx);
  *lane = ctx->lane[l];
  return KZG_OK;
}

// lane of the group the next ticket joins + result slot + staging buffer of that ticket, allocated on first use
static int32_t pending_slot(kzg_ctx* ctx, uint64_t n, bool host, kzg_ctx** lane_out, int* lane_index_out, kzg_ctx::Pending** q_out) {
  latch_queue(ctx);
  const uint32_t t = ctx->next_ticket;
  const int lane_index = (int)(ctx->groups_launched % (uint32_t)ctx->lanes_in_use);
  kzg_ctx* lane;
  KZG_TRY(lane_at(ctx, lane_index, &lane));
  kzg_ctx::Pending& q = ctx->pending[t % queue_slots(ctx)];
  if (q.busy) return fail(ctx, KZG_E_BAD_ARG, "every lane has a call in flight: wait for the oldest ticket first");
  if (!q.copied) {
    KZG_CUDA_OK(ctx, cudaEventCreateWithFlags(&q.copied, cudaEventDisableTiming));
    KZG_CUDA_OK(ctx, cudaEventCreateWithFlags(&q.done, cudaEventDisableTiming));
    KZG_CUDA_OK(ctx, cudaMalloc(&q.result, 512));
    KZG_CUDA_OK(ctx, cudaHostAlloc(reinterpret_cast<void**>(&q.out_host), 64, cudaHostAllocDefault));
  }
  if (host) {
    const size_t bytes = n ? n * 32 : 32;
    if (q.cap < bytes) {
      if (q.scalars) { KZG_CUDA_OK(ctx, cudaFree(q.scalars)); q.scalars = nullptr; q.cap = 0; }
      KZG_CUDA_OK(ctx, cudaMalloc(&q.scalars, bytes));
      q.cap = bytes;
    }
  }
  *lane_out = lane; *lane_index_out = lane_index; *q_out = &q;
  return KZG_OK;
}

// Everything the NEXT submitted call of n scalars will allocate, allocated now: the lane (streams, events), its result
// slot and staging buffer, the setup's table, and the lane's scratch arenas (sized by a dry run of the sum, whose result
// is dropped).  Device allocations synchronise with the device; a group whose members SHARE a device must therefore not
// allocate between one member's exchange kernel (which spins until its peers publish) and another member's launches —
// abi.cu reserves on every member first, then submits (two phases).  A no-op once the lane has seen this size.
int32_t reserve_call(kzg_ctx* ctx, const kzg_setup* s, const void* scalars, uint64_t n, bool host) {
  if (!ctx || !s) return KZG_E_BAD_ARG;
  if (n > s->n) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  DeviceGuard g(ctx->device);
  kzg_ctx* lane; int lane_index; kzg_ctx::Pending* q;
  KZG_TRY(pending_slot(ctx, n, host, &lane, &lane_index, &q));
  if (n == 0) return KZG_OK;
  if (!s->msm_tab && !(s->fix_tab && n <= s->fix_n)) {      // a table build allocates too: never leave it to the submit
    KZG_TRY(ensure_msm_table(ctx, s));
    KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (lane->reserved_n >= n) return KZG_OK;
  const uint8_t* sc = reinterpret_cast<const uint8_t*>(scalars);
  if (host) {      // whatever the staging buffer holds are valid scalars (any 32 bytes are): the run only sizes the arenas
    KZG_CUDA_OK(ctx, cudaMemsetAsync(q->scalars, 0x5a, n * 32, lane->stream));
    sc = reinterpret_cast<const uint8_t*>(q->scalars);
  }
  void *r, *o;
  KZG_TRY(scratch_get(lane, SCR_MISC, 1024, &r));
  KZG_TRY(scratch_get(lane, SCR_OUT, 64 * KZG_MAX_COALESCE, &o));      // the launch writes its compressed results there
  int32_t st = msm_any(lane, s, sc, n, reinterpret_cast<G1Xyzz*>(r));
  if (st == KZG_OK) st = finalize_points(lane, reinterpret_cast<G1Xyzz*>(r), 1, reinterpret_cast<uint8_t*>(q->result) + 256);
  if (st != KZG_OK) { if (lane != ctx) ctx->err = lane->err; return st; }
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(lane->stream));
  lane->reserved_n = n;
  return KZG_OK;
}

// Launch the open group (1 .. B staged calls with the same setup, length and kind) on the next lane: one bucket problem
// for all of them, then the compression (or the peer exchange) of every result and the read-backs the waits collect.
// A failure is kept in every call of the group and reported by its wait.
static int32_t flush_open_group(kzg_ctx* ctx) {
  kzg_ctx::OpenGroup& og = ctx->open_group;
  const int B = og.count;
  if (B == 0) return KZG_OK;
  DeviceGuard g(ctx->device);
  const uint32_t slots = queue_slots(ctx);
  kzg_ctx::Pending* q[KZG_MAX_COALESCE];
  for (int k = 0; k < B; k++) q[k] = &ctx->pending[(og.first_ticket + (uint32_t)k) % slots];
  const kzg_setup* s = og.s;
  const uint64_t n = og.n;
  const bool sharded = og.sharded;
  og.count = 0;
  const int lane_index = (int)(ctx->groups_launched % (uint32_t)ctx->lanes_in_use);
  kzg_ctx* lane = nullptr;
  auto run = [&]() -> int32_t {
    KZG_TRY(lane_at(ctx, lane_index, &lane));
    if (n && !s->msm_tab && !(s->fix_tab && n <= s->fix_n)) {      // built once, on the main lane, before any lane reads it
      KZG_TRY(ensure_msm_table(ctx, s));
      KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (int k = 0; k < B; k++)
      if (q[k]->host) KZG_CUDA_OK(lane, cudaStreamWaitEvent(lane->stream, q[k]->copied, 0));
    void *r, *o;
    KZG_TRY(scratch_get(lane, SCR_MISC, 1024, &r));                  // B <= 4 partial sums of 192 bytes
    KZG_TRY(scratch_get(lane, SCR_OUT, 64 * KZG_MAX_COALESCE, &o));  // per call 48 compressed bytes (+ the status word of an exchange)
    G1Xyzz* res = reinterpret_cast<G1Xyzz*>(r);
    uint8_t* out_dev = reinterpret_cast<uint8_t*>(o);
    // tuned for throughput instead of the shortest chain when calls ARE in flight together: a fused group, or tickets of
    // earlier groups that nobody has collected yet (a caller that submits and waits one call at a time keeps the short chain)
    int others = 0;
    for (uint32_t k = 0; k < slots; k++) others += (ctx->pending[k].busy && ctx->pending[k].launched) ? 1 : 0;
    lane->throughput_mode = B > 1 || others > 0;
    int32_t st = KZG_OK;
    if (B == 1 || n == 0 || (s->fix_tab && n <= s->fix_n)) {
      for (int k = 0; k < B && st == KZG_OK; k++) st = msm_any(lane, s, q[k]->dev_scalars, n, res + k);
    } else {
      const uint8_t* ptrs[KZG_MAX_COALESCE];
      for (int k = 0; k < B; k++) ptrs[k] = q[k]->dev_scalars;
      // the bucket arrays of one problem hold at most 2^20 buckets: wide windows fuse fewer calls at a time
      const uint64_t half = 1ull << (s->msm_c - 1);
      const int per = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)B, (1ull << 20) / half));
      for (int k0 = 0; k0 < B && st == KZG_OK; k0 += per) {
        const int nb = std::min(per, B - k0);
        st = nb == 1 ? msm_any(lane, s, ptrs[k0], n, res + k0) : msm_bucket_multi(lane, s, ptrs + k0, nb, n, res + k0);
      }
    }
    lane->throughput_mode = false;
    KZG_TRY(st);
    uint32_t rec = 48;
    if (sharded) {
      // collective calls are numbered by their own counter (every rank issues the same sequence of SHARDED calls; plain
      // submits in between do not count); the calls of a group take consecutive numbers, one block of the kernel each
      rec = 64;
      KZG_TRY(exchange_combine(lane, ctx, ctx->p2p_seq + 1, res, out_dev, reinterpret_cast<uint32_t*>(out_dev + 48), B, rec));
      ctx->p2p_seq += (uint32_t)B;
    } else {
      KZG_TRY(finalize_points(lane, res, (uint64_t)B, out_dev));
    }
    for (int k = 0; k < B; k++) {
      KZG_CUDA_OK(lane, cudaMemcpyAsync(q[k]->out_host, out_dev + (size_t)k * rec, rec, cudaMemcpyDeviceToHost, lane->stream));
      KZG_CUDA_OK(lane, cudaEventRecord(q[k]->done, lane->stream));
    }
    return KZG_OK;
  };
  const int32_t st = run();
  if (st != KZG_OK) {
    if (lane && lane != ctx) ctx->err = lane->err;
    if (sharded) ctx->p2p_failed = true;                // the peers may already be waiting for partials that will not come
  }
  for (int k = 0; k < B; k++) { q[k]->launched = true; q[k]->status = st; q[k]->lane_index = lane_index; }
  ctx->last_lane = lane_index;
  ctx->groups_launched++;
  return st;
}

// host == true: `scalars` is a host buffer, copied on the lane's copy stream (device pointers may live on ANY device
// the context's device has peer access to: the kernels read them over NVLink); sharded == true: the result is the
// fold of every rank's partial (peer exchange), else this context's own sum
int32_t submit_common(kzg_ctx* ctx, const kzg_setup* s, const void* scalars, uint64_t n, bool host, bool sharded,
                      int32_t* ticket) {
  if (!ctx || !s || (!scalars && n) || !ticket) return KZG_E_BAD_ARG;
  if (n > s->n) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  if (sharded && !ctx->p2p_connected) return fail(ctx, KZG_E_BAD_ARG, "peer mailboxes not connected (kzg_p2p_init / kzg_p2p_connect)");
  if (sharded && ctx->p2p_failed)
    return fail(ctx, KZG_E_CUDA, "the peer exchange of this context failed earlier (a rank timed out): its sequence numbers are "
                                 "no longer in step, create the contexts again");
  DeviceGuard g(ctx->device);
  latch_queue(ctx);
  kzg_ctx::OpenGroup& og = ctx->open_group;
  // a call that cannot ride with the staged ones sends them off first (their failure, if any, is theirs to report)
  if (og.count > 0 && (og.s != s || og.n != n || og.sharded != sharded)) (void)flush_open_group(ctx);
  const uint32_t t = ctx->next_ticket;
  kzg_ctx* lane; int lane_index; kzg_ctx::Pending* qp;
  KZG_TRY(pending_slot(ctx, n, host, &lane, &lane_index, &qp));
  kzg_ctx::Pending& q = *qp;
  q.dev_scalars = reinterpret_cast<const uint8_t*>(scalars);
  if (host) {
    if (n) KZG_CUDA_OK(ctx, cudaMemcpyAsync(q.scalars, scalars, n * 32, cudaMemcpyHostToDevice, lane->copy_stream));
    KZG_CUDA_OK(ctx, cudaEventRecord(q.copied, lane->copy_stream));
    q.dev_scalars = reinterpret_cast<const uint8_t*>(q.scalars);
  }
  q.busy = true; q.launched = false; q.status = KZG_OK; q.sharded = sharded; q.host = host; q.ticket = t; q.lane_index = lane_index;
  if (og.count == 0) { og.first_ticket = t; og.s = s; og.n = n; og.sharded = sharded; }
  og.count++;
  ctx->next_ticket = t + 1;
  *ticket = (int32_t)t;
  if (og.count >= ctx->coalesce_in_use) {
    const int32_t st = flush_open_group(ctx);
    if (st != KZG_OK) {          // this call never got a ticket: give its slot back; the earlier calls of the group report st from their waits
      q.busy = false;
      ctx->next_ticket = t;
      *ticket = -1;
      return st;
    }
  }
  return KZG_OK;
}

int32_t one_msm_g1_submit(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_le32, uint64_t n, int32_t* ticket) {
  return submit_common(ctx, s, scalars_le32, n, true, false, ticket);
}
int32_t one_msm_g1_submit_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, int32_t* ticket) {
  return submit_common(ctx, s, scalars_dev, n, false, false, ticket);
}
int32_t one_msm_g1_sharded_submit_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, int32_t* ticket) {
  return submit_common(ctx, s, scalars_dev, n, false, true, ticket);
}
int32_t one_msm_g1_sharded_submit(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_le32, uint64_t n, int32_t* ticket) {
  return submit_common(ctx, s, scalars_le32, n, true, true, ticket);
}

uint32_t queue_capacity(kzg_ctx* ctx) { latch_queue(ctx); return queue_slots(ctx); }

// launch whatever is still staged (kzg_ctx_sync, the timers, and anybody else who is about to look at the device)
int32_t flush_submitted(kzg_ctx* ctx) {
  if (!ctx || ctx->open_group.count == 0) return KZG_OK;
  (void)flush_open_group(ctx);      // a failure belongs to the staged calls: their waits report it
  return KZG_OK;
}

int32_t one_msm_g1_wait(kzg_ctx* ctx, int32_t ticket, uint8_t out_c48[48]) {
  if (!ctx || !out_c48 || ticket < 0) return KZG_E_BAD_ARG;
  if (!ctx->lanes_in_use) return fail(ctx, KZG_E_BAD_ARG, "no call in flight under this ticket");
  const uint32_t t = (uint32_t)ticket, F = queue_slots(ctx);
  kzg_ctx::Pending& q = ctx->pending[t % F];
  if (!q.busy || q.ticket != t || ctx->next_ticket - t < 1 || ctx->next_ticket - t > F)
    return fail(ctx, KZG_E_BAD_ARG, "no call in flight under this ticket");
  DeviceGuard g(ctx->device);
  if (!q.launched) (void)flush_open_group(ctx);       // still staged: send the partial group off now
  if (q.status != KZG_OK) {                           // the launch that carried this call failed (ctx->err has the first message)
    q.busy = false;
    return q.status;
  }
  KZG_CUDA_OK(ctx, cudaEventSynchronize(q.done));
  q.busy = false;
  ctx->last_lane = q.lane_index;
  ctx->timers_stale = true;      // the phase timers are read when somebody asks (three event queries per wait add up: eight members per group wait)
  if (q.sharded) {
    uint32_t st;
    memcpy(&st, q.out_host + 48, 4);
    if (st != 0) {
      ctx->p2p_failed = true;
      return fail(ctx, KZG_E_CUDA, st == 1 ? "peer exchange timed out: a rank did not publish its partial sum (every rank of this call fails)"
                                           : "peer exchange abandoned: another rank timed out waiting and gave this call up");
    }
  }
  memcpy(out_c48, q.out_host, 48);
  return KZG_OK;
}

// ---- multi-GPU exchange over peer memory ---------------------------------------------------------------------------
int32_t p2p_alloc_box(kzg_ctx* ctx, int32_t world, int32_t rank) {
  if (world < 1 || world > KZG_MAX_PEERS || rank < 0 || rank >= world) return KZG_E_BAD_ARG;
  if (ctx->p2p_box[ctx->p2p_rank] && ctx->p2p_world) return fail(ctx, KZG_E_BAD_ARG, "peer exchange already initialised on this context");
  void* box = nullptr;
  KZG_CUDA_OK(ctx, cudaMalloc(&box, sizeof(PeerMailbox)));
  KZG_CUDA_OK(ctx, cudaMemset(box, 0, sizeof(PeerMailbox)));
  KZG_CUDA_OK(ctx, cudaDeviceSynchronize());
  ctx->p2p_world = world; ctx->p2p_rank = rank; ctx->p2p_box[rank] = box; ctx->p2p_connected = false;
  ctx->p2p_failed = false; ctx->p2p_seq = 0;
  return KZG_OK;
}

// all ranks in this process: boxes[r] is rank r's mailbox, reachable through plain peer access (cudaDeviceEnablePeerAccess)
int32_t p2p_attach_local(kzg_ctx* ctx, void* const* boxes) {
  if (ctx->p2p_world < 1 || !ctx->p2p_box[ctx->p2p_rank]) return KZG_E_BAD_ARG;
  for (int r = 0; r < ctx->p2p_world; r++) if (r != ctx->p2p_rank) ctx->p2p_box[r] = boxes[r];
  ctx->p2p_ipc = false;
  ctx->p2p_connected = true;
  return KZG_OK;
}

int32_t one_p2p_init(kzg_ctx* ctx, int32_t world, int32_t rank, uint8_t handle_out[KZG_P2P_HANDLE_BYTES]) {
  if (!ctx || !handle_out) return KZG_E_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == KZG_P2P_HANDLE_BYTES, "IPC handle size");
  DeviceGuard g(ctx->device);
  KZG_TRY(p2p_alloc_box(ctx, world, rank));
  cudaIpcMemHandle_t h;
  KZG_CUDA_OK(ctx, cudaIpcGetMemHandle(&h, ctx->p2p_box[rank]));
  memcpy(handle_out, &h, sizeof(h));
  return KZG_OK;
}

int32_t one_p2p_connect(kzg_ctx* ctx, const uint8_t* handles) {
  if (!ctx || !handles || ctx->p2p_world < 1 || !ctx->p2p_box[ctx->p2p_rank]) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  for (int r = 0; r < ctx->p2p_world; r++) {
    if (r == ctx->p2p_rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * KZG_P2P_HANDLE_BYTES, sizeof(h));
    KZG_CUDA_OK(ctx, cudaIpcOpenMemHandle(&ctx->p2p_box[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  ctx->p2p_ipc = true;
  ctx->p2p_connected = true;
  return KZG_OK;
}

int32_t one_msm_g1_sharded_dev(kzg_ctx* ctx, const kzg_setup* s, const void* scalars_dev, uint64_t n, uint8_t out_c48[48]) {
  if (!ctx || !s || !out_c48) return KZG_E_BAD_ARG;
  int32_t t = -1;
  KZG_TRY(one_msm_g1_sharded_submit_dev(ctx, s, scalars_dev, n, &t));
  return one_msm_g1_wait(ctx, t, out_c48);
}

int32_t one_g1_combine_dev(kzg_ctx* ctx, const void* partials_dev, uint64_t count, uint8_t out_c48[48]) {
  if (!ctx || (!partials_dev && count) || !out_c48) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  void* r;
  KZG_TRY(scratch_get(ctx, SCR_MISC, 1024, &r));
  G1Xyzz* res = reinterpret_cast<G1Xyzz*>(r);
  KZG_TRY(sum_points(ctx, reinterpret_cast<const G1Xyzz*>(partials_dev), count, res));
  return finish_one(ctx, res, out_c48);
}

int32_t one_commit(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t n, uint8_t out_c48[48]) {
  if (!ctx || !s || (!coeffs_le32 && n) || !out_c48) return KZG_E_BAD_ARG;
  uint64_t nt = trimmed_len(coeffs_le32, n);
  // (degree() + 1) > setup.len()  with degree() == 0 for the empty / zero polynomial (src/polynomial.rs:95-100,201)
  if (std::max<uint64_t>(nt, 1) > s->n) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  return one_msm_g1(ctx, s, coeffs_le32, nt, out_c48);
}

int32_t one_evaluate_dev(kzg_ctx* ctx, const void* coeffs_dev, uint64_t n, const uint8_t z_le[32], uint8_t y_le[32]) {
  if (!ctx || (!coeffs_dev && n) || !z_le || !y_le) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  void* m;
  KZG_TRY(scratch_get(ctx, SCR_FR_B, 256, &m));
  uint8_t* zd = reinterpret_cast<uint8_t*>(m);
  uint8_t* yd = zd + 64;
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(zd, z_le, 32, cudaMemcpyHostToDevice, ctx->stream));
  KZG_TRY(poly_eval_quotient(ctx, reinterpret_cast<const uint8_t*>(coeffs_dev), 1, n, zd, yd, nullptr));
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(y_le, yd, 32, cudaMemcpyDeviceToHost, ctx->stream));
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return KZG_OK;
}

int32_t one_evaluate(kzg_ctx* ctx, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32], uint8_t y_le[32]) {
  if (!ctx || (!coeffs_le32 && n) || !z_le || !y_le) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  void* d;
  KZG_TRY(scratch_get(ctx, SCR_SCALARS, n * 32, &d));
  if (n) KZG_CUDA_OK(ctx, cudaMemcpyAsync(d, coeffs_le32, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  return one_evaluate_dev(ctx, d, n, z_le, y_le);
}

// shared by one_quotient / one_open: stages coefficients, runs the scan, applies the reference's
// divide_by_root error rules.  On success *q_dev holds nt-1 (or 0) canonical coefficients.
int32_t quotient_core(kzg_ctx* ctx, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32],
                             const uint8_t* y_le, uint8_t y_out[32], uint8_t** q_dev, uint64_t* q_len) {
  uint64_t nt = trimmed_len(coeffs_le32, n);
  void *d, *m, *q;
  KZG_TRY(scratch_get(ctx, SCR_SCALARS, std::max<uint64_t>(nt, 1) * 32, &d));
  KZG_TRY(scratch_get(ctx, SCR_QUOT, std::max<uint64_t>(nt, 1) * 32, &q));
  KZG_TRY(scratch_get(ctx, SCR_FR_B, 256, &m));
  uint8_t* zd = reinterpret_cast<uint8_t*>(m);
  uint8_t* yd = zd + 64;
  if (nt) KZG_CUDA_OK(ctx, cudaMemcpyAsync(d, coeffs_le32, nt * 32, cudaMemcpyHostToDevice, ctx->stream));
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(zd, z_le, 32, cudaMemcpyHostToDevice, ctx->stream));
  KZG_TRY(poly_eval_quotient(ctx, reinterpret_cast<uint8_t*>(d), 1, nt, zd, yd, reinterpret_cast<uint8_t*>(q)));
  uint8_t y_dev_val[32];
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(y_dev_val, yd, 32, cudaMemcpyDeviceToHost, ctx->stream));
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (y_out) memcpy(y_out, y_dev_val, 32);
  bool y_matches = true;
  if (y_le) {
    uint8_t yr[32];
    kzg_fr_reduce(y_le, yr);
    y_matches = memcmp(yr, y_dev_val, 32) == 0;
  }
  if (nt <= 1) {
    // p - y is a constant: zero -> empty quotient, otherwise "Unable to divide a constant polynomial"
    if (!y_matches) return fail(ctx, KZG_E_CONSTANT_POLY, kzg_status_string(KZG_E_CONSTANT_POLY));
    *q_len = 0;
  } else {
    if (!y_matches) return fail(ctx, KZG_E_REMAINDER_NONZERO, kzg_status_string(KZG_E_REMAINDER_NONZERO));
    *q_len = nt - 1;
  }
  *q_dev = reinterpret_cast<uint8_t*>(q);
  return KZG_OK;
}

int32_t one_quotient(kzg_ctx* ctx, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32],
                     const uint8_t* y_le, uint8_t* q_le32, uint64_t* q_len) {
  if (!ctx || (!coeffs_le32 && n) || !z_le || !q_len) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  uint8_t* qd = nullptr;
  KZG_TRY(quotient_core(ctx, coeffs_le32, n, z_le, y_le, nullptr, &qd, q_len));
  if (*q_len && q_le32) {
    KZG_CUDA_OK(ctx, cudaMemcpyAsync(q_le32, qd, *q_len * 32, cudaMemcpyDeviceToHost, ctx->stream));
    KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return KZG_OK;
}

int32_t one_open(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t n, const uint8_t z_le[32],
                 const uint8_t* y_expected_le, uint8_t y_out_le[32], uint8_t proof_c48[48]) {
  if (!ctx || !s || (!coeffs_le32 && n) || !z_le || !proof_c48) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  uint8_t* qd = nullptr;
  uint64_t ql = 0;
  KZG_TRY(quotient_core(ctx, coeffs_le32, n, z_le, y_expected_le, y_out_le, &qd, &ql));
  if (std::max<uint64_t>(ql, 1) > s->n) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  return one_msm_g1_dev(ctx, s, qd, ql, proof_c48);
}

// ---- blob batches ------------------------------------------------------------------------------------
// lane l >= 1 of a context (hidden sibling context: own streams and scratch), created on first use
static int32_t sibling_lane(kzg_ctx* ctx, int l, kzg_ctx** lane) {
  if (l < 1 || l >= KZG_MAX_LANES) return KZG_E_BAD_ARG;
  if (!ctx->lane[l]) {
    int32_t st = one_ctx_create(&ctx->lane[l], ctx->device);
    if (st != KZG_OK) return fail(ctx, st, "could not create another lane of the context");
  }
  copy_options(ctx->lane[l], ctx);
  *lane = ctx->lane[l];
  return KZG_OK;
}

static int32_t batch_points_out(kzg_ctx* ctx, const G1Xyzz* pts, uint64_t blobs, uint8_t* out_c48_dev) {
  void* aff;
  KZG_TRY(scratch_get(ctx, SCR_BATCH_AFF, blobs * sizeof(G1Affine), &aff));
  KZG_TRY(batch_to_affine(ctx, pts, blobs, reinterpret_cast<G1Affine*>(aff)));
  return compress_points(ctx, reinterpret_cast<G1Affine*>(aff), blobs, out_c48_dev);
}

static int32_t msm_batch_any(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_dev, uint64_t blobs, uint64_t n,
                             uint64_t stride, G1Xyzz* out) {
  if (n == 0) return set_infinity(ctx, out, (uint32_t)blobs);   // empty sums: no table is touched (or built)
  if (s->fix_tab && n <= s->fix_n) return msm_fixed_batch(ctx, s, scalars_dev, blobs, n, stride, out);
  KZG_TRY(ensure_msm_table(ctx, s));
  // rows beyond the fixed-base table (long polynomials): up to four at a time as ONE bucket problem (msm_bucket_multi)
  const uint64_t half = 1ull << (s->msm_c - 1);
  const uint64_t per = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)KZG_MAX_COALESCE, (1ull << 20) / half));
  const bool was = ctx->throughput_mode;
  ctx->throughput_mode = was || blobs > 1;
  int32_t st = KZG_OK;
  for (uint64_t b = 0; b < blobs && st == KZG_OK; b += per) {
    const int nb = (int)std::min<uint64_t>(per, blobs - b);
    const uint8_t* ptrs[KZG_MAX_COALESCE];
    for (int k = 0; k < nb; k++) ptrs[k] = scalars_dev + (b + (uint64_t)k) * stride * 32;
    st = msm_bucket_multi(ctx, s, ptrs, nb, n, out + b);
  }
  ctx->throughput_mode = was;
  return st;
}

// Rows are `n` coefficients apart; only the first `used` of a row can be non-zero (the rest is trailing-zero padding that
// Polynomial::try_from would have trimmed, src/polynomial.rs:63-72), so the sums run over `used` terms.
static int32_t commit_rows_dev(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_dev, uint64_t blobs, uint64_t n,
                               uint64_t used, uint8_t* out_c48_dev) {
  void* pts;
  KZG_TRY(scratch_get(ctx, SCR_BATCH_PTS, blobs * sizeof(G1Xyzz), &pts));
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  KZG_TRY(msm_batch_any(ctx, s, coeffs_dev, blobs, used, n, reinterpret_cast<G1Xyzz*>(pts)));
  KZG_TRY(batch_points_out(ctx, reinterpret_cast<G1Xyzz*>(pts), blobs, out_c48_dev));
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->last_lane = 0;
  return KZG_OK;
}

// the quotient of a row has n - 1 coefficients of which the first `used_q` can be non-zero
static int32_t open_rows_dev(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_dev, uint64_t blobs, uint64_t n,
                             uint64_t used_q, const uint8_t* z_dev, uint8_t* y_dev, uint8_t* proofs_c48_dev) {
  void *pts, *q;
  KZG_TRY(scratch_get(ctx, SCR_BATCH_PTS, blobs * sizeof(G1Xyzz), &pts));
  KZG_TRY(scratch_get(ctx, SCR_QUOT, std::max<uint64_t>(blobs * n, 1) * 32, &q));
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  KZG_TRY(poly_eval_quotient(ctx, coeffs_dev, blobs, n, z_dev, y_dev, reinterpret_cast<uint8_t*>(q)));
  KZG_TRY(msm_batch_any(ctx, s, reinterpret_cast<uint8_t*>(q), blobs, used_q, n, reinterpret_cast<G1Xyzz*>(pts)));
  KZG_TRY(batch_points_out(ctx, reinterpret_cast<G1Xyzz*>(pts), blobs, proofs_c48_dev));
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->last_lane = 0;
  return KZG_OK;
}

// commit + open of the same rows: two independent MSM batches, the commitments on the context's own lane and the
// proofs (quotient scan + MSM) on a sibling lane, so that the latency-bound inversion chains of one tree (3 small
// launches per round, ~9 % of a batch) run under the bulk kernels of the other.  Inputs must be ready on ctx->stream;
// on return ctx->stream has joined the sibling lane.
static int32_t commit_open_rows_dev(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_dev, uint64_t blobs, uint64_t n,
                                    uint64_t used, const uint8_t* z_dev, uint8_t* commits_c48_dev, uint8_t* y_dev,
                                    uint8_t* proofs_c48_dev) {
  kzg_ctx* side = ctx;
  if (ctx->opt_lanes > 1) KZG_TRY(sibling_lane(ctx, 1, &side));
  if (side != ctx) {
    KZG_CUDA_OK(ctx, cudaEventRecord(side->ev_copy[0], ctx->stream));          // inputs (and earlier users of the outputs) done
    KZG_CUDA_OK(ctx, cudaStreamWaitEvent(side->stream, side->ev_copy[0], 0));
  }
  int32_t st = open_rows_dev(side, s, coeffs_dev, blobs, n, used ? used - 1 : 0, z_dev, y_dev, proofs_c48_dev);
  if (st != KZG_OK) { if (side != ctx) ctx->err = side->err; return st; }
  KZG_TRY(commit_rows_dev(ctx, s, coeffs_dev, blobs, n, used, commits_c48_dev));
  if (side != ctx) {
    KZG_CUDA_OK(ctx, cudaEventRecord(side->ev_copy[1], side->stream));
    KZG_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, side->ev_copy[1], 0));
  }
  return KZG_OK;
}

// Polynomial::commit's check on rows in device memory (src/polynomial.rs:201-205, trimmed length against the setup):
// *used = number of leading coefficients per row the sums have to cover, at most `limit`
static int32_t rows_used_dev(kzg_ctx* ctx, const uint8_t* coeffs_dev, uint64_t blobs, uint64_t n, uint64_t limit, uint64_t* used) {
  *used = n;
  if (n <= limit) return KZG_OK;
  bool any = false;
  KZG_TRY(any_nonzero_beyond(ctx, coeffs_dev, blobs, n, limit, &any));
  if (any) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  *used = limit;
  return KZG_OK;
}

int32_t one_commit_batch_dev(kzg_ctx* ctx, const kzg_setup* s, const void* coeffs_dev, uint64_t blobs, uint64_t n,
                             void* out_c48_dev) {
  if (!ctx || !s || (!coeffs_dev && blobs * n) || !out_c48_dev) return KZG_E_BAD_ARG;
  if (blobs == 0) return KZG_OK;
  if (s->n == 0) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));   // degree() + 1 >= 1 > 0
  DeviceGuard g(ctx->device);
  uint64_t used;
  KZG_TRY(rows_used_dev(ctx, reinterpret_cast<const uint8_t*>(coeffs_dev), blobs, n, s->n, &used));
  return commit_rows_dev(ctx, s, reinterpret_cast<const uint8_t*>(coeffs_dev), blobs, n, used, reinterpret_cast<uint8_t*>(out_c48_dev));
}

int32_t one_open_batch_dev(kzg_ctx* ctx, const kzg_setup* s, const void* coeffs_dev, uint64_t blobs, uint64_t n,
                           const void* z_dev, void* y_dev, void* proofs_c48_dev) {
  if (!ctx || !s || (!coeffs_dev && blobs * n) || !z_dev || !y_dev || !proofs_c48_dev) return KZG_E_BAD_ARG;
  if (blobs == 0) return KZG_OK;
  if (s->n == 0) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  DeviceGuard g(ctx->device);
  uint64_t used;                              // a quotient is one shorter than its polynomial
  KZG_TRY(rows_used_dev(ctx, reinterpret_cast<const uint8_t*>(coeffs_dev), blobs, n, s->n + 1, &used));
  return open_rows_dev(ctx, s, reinterpret_cast<const uint8_t*>(coeffs_dev), blobs, n, used ? used - 1 : 0,
                       reinterpret_cast<const uint8_t*>(z_dev), reinterpret_cast<uint8_t*>(y_dev), reinterpret_cast<uint8_t*>(proofs_c48_dev));
}

int32_t one_commit_open_batch_dev(kzg_ctx* ctx, const kzg_setup* s, const void* coeffs_dev, uint64_t blobs, uint64_t n,
                                  const void* z_dev, void* commits_c48_dev, void* y_dev, void* proofs_c48_dev) {
  if (!ctx || !s || (!coeffs_dev && blobs * n) || !z_dev || !commits_c48_dev || !y_dev || !proofs_c48_dev) return KZG_E_BAD_ARG;
  if (blobs == 0) return KZG_OK;
  if (s->n == 0) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  DeviceGuard g(ctx->device);
  uint64_t used;
  KZG_TRY(rows_used_dev(ctx, reinterpret_cast<const uint8_t*>(coeffs_dev), blobs, n, s->n, &used));
  return commit_open_rows_dev(ctx, s, reinterpret_cast<const uint8_t*>(coeffs_dev), blobs, n, used, reinterpret_cast<const uint8_t*>(z_dev),
                              reinterpret_cast<uint8_t*>(commits_c48_dev), reinterpret_cast<uint8_t*>(y_dev),
                              reinterpret_cast<uint8_t*>(proofs_c48_dev));
}

// rows per host->device staging chunk of a batch call: the tree path's chunk for large batches, else everything
static uint64_t batch_chunk_rows(const kzg_ctx* ctx, uint64_t blobs) {
  const uint64_t c = ctx->opt_tree_chunk_rows > 0 ? (uint64_t)ctx->opt_tree_chunk_rows : 2048;
  return (blobs >= 2 * c && blobs >= (uint64_t)ctx->opt_tree_min_rows) ? c : blobs;
}

// The first staging chunk is an eighth of the others: its copy is the only one that nothing hides (2048 rows of 4096
// coefficients = 256 MiB, ~10 ms), and its kernels then cover the copy of the first full chunk.
static uint64_t first_chunk_rows(uint64_t chunk, uint64_t blobs, uint64_t r0) {
  return (r0 == 0 && blobs > chunk && chunk >= 64) ? chunk / 8 : chunk;
}

// Polynomial::commit's check on rows in host memory: the longest TRIMMED row must fit the setup (src/polynomial.rs:201-205);
// *used = leading coefficients per row the device has to look at (everything beyond is zero in every row)
static int32_t rows_used_host(kzg_ctx* ctx, const uint8_t* coeffs, uint64_t blobs, uint64_t n, uint64_t limit, uint64_t* used) {
  *used = n;
  if (n <= limit) return KZG_OK;
  uint64_t m = 0;
  for (uint64_t b = 0; b < blobs && m <= limit; b++) {
    const uint8_t* row = coeffs + b * n * 32;
    uint64_t len = n;
    while (len > m && le32_is_zero_mod_r(row + (len - 1) * 32)) len--;     // only interested in len > m
    m = std::max(m, len);
  }
  if (m > limit) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  *used = limit;
  return KZG_OK;
}

// what the three host-buffer batch calls share: stage the rows in chunks on the copy stream, chunk k+1 beside the
// kernels of chunk k.  mode: 1 commit, 2 open, 3 both (one pass over the coefficients).
static int32_t batch_host(kzg_ctx* ctx, const kzg_setup* s, int mode, const uint8_t* coeffs_le32, uint64_t blobs, uint64_t n,
                          const uint8_t* z_le32, uint8_t* commits_c48, uint8_t* y_le32, uint8_t* proofs_c48) {
  if (blobs == 0) return KZG_OK;
  if (s->n == 0) return fail(ctx, KZG_E_DEGREE_TOO_HIGH, kzg_status_string(KZG_E_DEGREE_TOO_HIGH));
  DeviceGuard g(ctx->device);
  uint64_t used;
  KZG_TRY(rows_used_host(ctx, coeffs_le32, blobs, n, (mode & 1) ? s->n : s->n + 1, &used));
  void *d, *o, *zy = nullptr;
  KZG_TRY(scratch_get(ctx, SCR_SCALARS, std::max<uint64_t>(blobs * n, 1) * 32, &d));
  KZG_TRY(scratch_get(ctx, SCR_OUT, blobs * 96, &o));
  uint8_t* cd = reinterpret_cast<uint8_t*>(o);
  uint8_t* pd = cd + blobs * 48;
  uint8_t *zd = nullptr, *yd = nullptr;
  if (mode & 2) {
    KZG_TRY(scratch_get(ctx, SCR_FR_A, blobs * 64, &zy));
    zd = reinterpret_cast<uint8_t*>(zy);
    yd = zd + blobs * 32;
    KZG_CUDA_OK(ctx, cudaMemcpyAsync(zd, z_le32, blobs * 32, cudaMemcpyHostToDevice, ctx->stream));
  }
  const uint64_t chunk = batch_chunk_rows(ctx, blobs);
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev_copy[7], ctx->stream));          // scratch of earlier calls is free
  KZG_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy[7], 0));
  uint64_t k = 0;
  for (uint64_t r0 = 0, rows = 0; r0 < blobs; r0 += rows, k++) {
    rows = std::min(first_chunk_rows(chunk, blobs, r0), blobs - r0);
    uint8_t* dc = reinterpret_cast<uint8_t*>(d) + r0 * n * 32;
    if (n) KZG_CUDA_OK(ctx, cudaMemcpyAsync(dc, coeffs_le32 + r0 * n * 32, rows * n * 32, cudaMemcpyHostToDevice, ctx->copy_stream));
    KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev_copy[k % 7], ctx->copy_stream));
    KZG_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[k % 7], 0));
    if (mode == 1) KZG_TRY(commit_rows_dev(ctx, s, dc, rows, n, used, cd + r0 * 48));
    else if (mode == 2) KZG_TRY(open_rows_dev(ctx, s, dc, rows, n, used ? used - 1 : 0, zd + r0 * 32, yd + r0 * 32, pd + r0 * 48));
    else KZG_TRY(commit_open_rows_dev(ctx, s, dc, rows, n, used, zd + r0 * 32, cd + r0 * 48, yd + r0 * 32, pd + r0 * 48));
  }
  if (mode & 1) KZG_CUDA_OK(ctx, cudaMemcpyAsync(commits_c48, cd, blobs * 48, cudaMemcpyDeviceToHost, ctx->stream));
  if (mode & 2) {
    KZG_CUDA_OK(ctx, cudaMemcpyAsync(y_le32, yd, blobs * 32, cudaMemcpyDeviceToHost, ctx->stream));
    KZG_CUDA_OK(ctx, cudaMemcpyAsync(proofs_c48, pd, blobs * 48, cudaMemcpyDeviceToHost, ctx->stream));
  }
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  update_timers(ctx);
  return KZG_OK;
}

int32_t one_commit_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t blobs, uint64_t n,
                         uint8_t* out_c48) {
  if (!ctx || !s || (!coeffs_le32 && blobs * n) || !out_c48) return KZG_E_BAD_ARG;
  return batch_host(ctx, s, 1, coeffs_le32, blobs, n, nullptr, out_c48, nullptr, nullptr);
}

int32_t one_open_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t blobs, uint64_t n,
                       const uint8_t* z_le32, uint8_t* y_le32, uint8_t* proofs_c48) {
  if (!ctx || !s || (!coeffs_le32 && blobs * n) || !z_le32 || !y_le32 || !proofs_c48) return KZG_E_BAD_ARG;
  return batch_host(ctx, s, 2, coeffs_le32, blobs, n, z_le32, nullptr, y_le32, proofs_c48);
}

// commit + open of every blob with ONE host->device pass over the coefficients (kzg_commit_batch followed by
// kzg_open_batch copies them twice): chunk k+1 is staged on the copy stream while chunk k is committed and opened
int32_t one_commit_open_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* coeffs_le32, uint64_t blobs, uint64_t n,
                              const uint8_t* z_le32, uint8_t* commits_c48, uint8_t* y_le32, uint8_t* proofs_c48) {
  if (!ctx || !s || (!coeffs_le32 && blobs * n) || !z_le32 || !commits_c48 || !y_le32 || !proofs_c48) return KZG_E_BAD_ARG;
  return batch_host(ctx, s, 3, coeffs_le32, blobs, n, z_le32, commits_c48, y_le32, proofs_c48);
}

// ---- diagnostics ---------------------------------------------------------------------------------------
int32_t one_probe_imad_peak(kzg_ctx* ctx, double* wide_mac_per_s) {
  if (!ctx || !wide_mac_per_s) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  return probe_imad(ctx, wide_mac_per_s);
}

int32_t one_probe_fq_mul(kzg_ctx* ctx, double* fq_mul_per_s) {
  if (!ctx || !fq_mul_per_s) return KZG_E_BAD_ARG;
  DeviceGuard g(ctx->device);
  return probe_fq_mul(ctx, fq_mul_per_s);
}

}  // namespace kzg
