// psinfer.cu -- context, schedule and C ABI of libpsinfer.so (see include/psinfer.h).
//
// Replaces object_detect::computeRootPosteriorRot / computePartMarginals / computeRotJointMarginal
// (reference src/libs/libPictStruct/objectdetect_findrot.cpp:470-727, :124-286, :292-456) with a schedule of
// sm_100a kernels over grids that stay resident in HBM.  No CPU fallback exists: every entry point that
// computes needs a CUDA device.
#include "../../include/psinfer.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#ifndef PS_RG
#define PS_RG 8
#endif
#ifndef PS_MBAR_HINT_NS
#define PS_MBAR_HINT_NS 100000u  // suspend-time hint of mbarrier.try_wait (ns)
#endif
#ifndef PS_TMA_MINB
#define PS_TMA_MINB 4  // register budget of k_conv_cols_tma2: 65536 / (4 * 288) -> 56 registers (see launch_conv_cols_tma)
#endif
#include "ps_geometry.hpp"
#include "ps_kernels.cuh"

namespace {

using psk::Affine;

thread_local std::string g_create_error;

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  cudaError_t alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  template <typename T>
  T *as() const { return (T *)p; }
};

// Device-side image of one psg::MessagePlan.
struct DevPlan {
  psg::MessagePlan host;
  DevBuf ints;    // xin | xout | yin | yout | in_shift | out_shift | ytiles | xtiles (xout stays 16-byte aligned whenever W % 4 == 0)
  DevBuf floats;  // rot taps | fx | fy
  DevBuf map;     // int2 [EH][EW], built on first sparse use
  DevBuf mats;    // T31 | T13 (doubles) for the map builder
  DevBuf walks;   // fused Gaussian kernel: int4 walks | x-block masks (bytes)
  int nwalks = 0;
  bool map_ready = false;
  int EP = 0;     // eigen-frame row pitch
  const int *xin() const { return ints.as<int>(); }
  const int *xout(int R, int H, int W) const { return ints.as<int>() + (size_t)R * W; }
  const int *yin(int R, int W) const { return ints.as<int>() + (size_t)R * 2 * W; }
  const int *yout(int R, int H, int W) const { return ints.as<int>() + (size_t)R * (2 * W + H); }
  const int *in_shift(int R, int H, int W) const { return host.in_pure ? ints.as<int>() + (size_t)R * 2 * (W + H) : nullptr; }
  const int *out_shift(int R, int H, int W) const { return host.out_pure ? ints.as<int>() + (size_t)R * 2 * (W + H) + 2 * R : nullptr; }
  const int *ytiles(int R, int H, int W) const { return ints.as<int>() + (size_t)R * 2 * (W + H) + 4 * R; }
  const int *xtiles(int R, int H, int W) const { return ytiles(R, H, W) + host.ytiles.size(); }
  const float *rot_taps() const { return floats.as<float>(); }
  const float *fx() const { return floats.as<float>() + host.rot_taps.size(); }
  const float *fy() const { return floats.as<float>() + host.rot_taps.size() + host.fx.size(); }
};

struct Node {
  int parent = -1;
  int joint = -1;               // joint connecting this node to its parent
  std::vector<int> children;    // in joint-list order (get_incoming_joints, findrot.cpp:59-71)
  std::vector<int> child_joints;
};

}  // namespace

struct ps_ctx {
  ps_config cfg{};
  int R = 0, S = 0, H = 0, W = 0, P = 0, root = -1;
  size_t HW = 0, N = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  std::string err;
  long long launches = 0;
  int num_sms = 148;
  bool disable_tma = false;  // PSINFER_NO_TMA=1: keep the shared-memory staged kernels (A/B testing)
  bool disable_tile_lists = false;  // PSINFER_ALL_TILES=1: filter every eigen-frame tile, used or not (A/B testing)

  // optional per-kernel-class device timing (ps_profile_enable)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Span { int klass; size_t e0, e1; };
  std::vector<Span> spans;

  // resident grids
  DevBuf unary;         // [P][S][N]
  DevBuf unary_backup;  // lazily, for PS_INFER_KEEP_UNARIES
  DevBuf post;          // [keep_all ? S : 1][P][N]
  DevBuf rootmsg;       // [n_root_children][N]: upward messages into the root, then fr_j
  DevBuf tmp[2];        // [N] down-pass inputs (unary + from_root)
  DevBuf bufB;          // [N] rotation-filtered / diag y-pass output
  DevBuf bufU, bufV;    // eigen-frame scratch [R][EHmax][EPmax] (>= N)
  // level-batched schedule: one (B, UT, V) scratch triple per message of a launch
  std::vector<DevBuf> slotB, slotU, slotV;
  size_t slot_elems = 0;
  DevBuf chain_tmp;     // [n_root_children][N]: down-pass inputs of the chains (ping-pong partner of rootmsg)
  // fixed work order of the fused Gaussian launches: per batch (plans, grid) the items of every block, dealt on the host
  std::map<std::string, std::shared_ptr<DevBuf>> gauss_tables;
  DevBuf work_counters; // unsigned [kWorkCounters]: one zeroed work-item counter per fused Gaussian launch
  int work_counter_next = 0;
  bool disable_batch = false;  // PSINFER_NO_BATCH=1: one message per launch through the two-pass Gaussian route (A/B testing)
  DevBuf root_post;     // [S][HW]
  DevBuf maxes;         // int encoded maxima: [0,P) beliefs | [P,P+16) from_root inputs | [P+16,2P+16) tmp of node q | 4 misc
  DevBuf upright_mask;  // uchar [R]
  DevBuf valid_rots;    // int [R]
  int n_valid_rots = 0;
  DevBuf argmax_keys;   // u64 [P]
  DevBuf cand;          // Cand [N] local-max candidates
  DevBuf topk_hist;     // unsigned [65536]
  DevBuf topk_state;    // TopKState [P + 1]
  DevBuf topk_out;      // Cand [(P + 1)][kmax] winners of every grid of one readout
  psk::Cand *host_topk = nullptr;       // pinned mirror of topk_out
  psk::TopKState *host_topk_state = nullptr;
  size_t topk_slots = 0, topk_kmax = 0;
  DevBuf counters;      // unsigned [8]
  DevBuf ingest_batch;              // ps_set_unaries_compact: compact cells of every grid of one call
  // ps_set_unaries_compact: when the whole unary buffer holds nothing but a LOG_ZERO fill plus the collision-free
  // scatter of one lattice (these transforms, this compact grid size), the next image on the same lattice overwrites
  // every lattice cell (LOG_ZERO where its score is 0) and the fill is skipped.  Every other writer of the unaries
  // clears the flag.
  bool lattice_clean = false;
  std::vector<double> lattice_sig;
  DevBuf ingest, ingest_keys;       // ps_set_unary_compact staging: Tig rows + compact cells; order keys [R][H][W]
  DevBuf table_stage, grid_stage;   // ps_add_unary_table(s) / ps_add_unary_grid: stream-ordered staging of host inputs
  DevBuf unary_max;                 // int [P][S]: encoded max of each unary as left by the ingest
  std::vector<unsigned char> unary_max_valid;  // [P][S]
  size_t scratch_elems = 0;

  // legacy POS_GAUSSIAN model (objectdetect_findpos.cpp): 2-D grids on an internal single-slice context
  bool pos_model = false, pending_root_only = false;
  ps_ctx *sub = nullptr;
  DevBuf pos_merged, pos_post, pos_msg;  // [P][HW] merged unaries | [P][HW] upward beliefs | [HW] one message

  // model
  bool joints_set = false;
  std::vector<ps_joint> joints;
  std::vector<Node> nodes;
  // plans[(joint*2 + dir) * S + scale], dir 0 = upward, 1 = downward
  std::vector<std::shared_ptr<DevPlan>> plans;
  // Plans are cached by the bytes of (offset_in, offset_out, C, rot_mean, rot_sigma, scale): the poselet-conditioned
  // model swaps whole joints per image from a finite per-joint table (aux.cpp:76-99), so ps_set_joints is a cache hit
  // after the first image that used a given type.  ps_message shares the cache.
  std::map<std::string, std::shared_ptr<DevPlan>> plan_cache;
  long long plan_cache_hits = 0, plan_cache_misses = 0;

  // CUDA graphs of whole inferences, keyed by (flags, joint set, which leaf maxima the ingest already holds)
  struct InferGraph {
    cudaGraphExec_t exec = nullptr;
    bool failed = false;
    long long launches = 0;
    std::vector<const float *> pending_grids;
    int pending_scaleidx = 0, pending_flags = 0, result_scale = -1;
  };
  std::map<std::string, InferGraph> graphs;
  long long joints_version = 0, graph_replays = 0, last_infer_version = -1;
  bool joints_reused = true;  // the joint set of this inference is the one of the previous inference (or there was none)
  bool disable_graph = false;  // PSINFER_NO_GRAPH=1

  // results.  ps_infer / ps_max_states only enqueue device work; the host part of the readout (decode of the
  // argmax keys, local-maximum selection) runs in finish_result() when a getter needs it.
  bool have_result = false;
  bool result_pending = false;
  int pending_flags = 0;
  int pending_scaleidx = 0;
  std::vector<const float *> pending_grids;
  unsigned long long *host_keys = nullptr;  // pinned [P]
  int result_scale = -1;
  std::vector<float> best_conf;               // [P][7]
  std::vector<std::vector<float>> part_hyps;  // per part rows of 7
  std::vector<float> root_hyps;               // rows of 4
  bool have_local_max = false, have_root_hyps = false;

  ~ps_ctx() {  // shared by ps_destroy and by the error returns of ps_create
    if (sub) ps_destroy(sub);
    if (own_stream) {
      cudaStreamSynchronize(own_stream);
      cudaStreamDestroy(own_stream);
    }
    for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
    for (auto &g : graphs)
      if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    if (host_keys) cudaFreeHost(host_keys);
    if (host_topk) cudaFreeHost(host_topk);
    if (host_topk_state) cudaFreeHost(host_topk_state);
  }
  int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
  float *U(int p, int s) const { return unary.as<float>() + ((size_t)p * S + s) * N; }
  float *POST(int p, int s) const {
    return post.as<float>() + ((size_t)(cfg.keep_all_scales ? s : 0) * P + p) * N;
  }
  int *MAXP(int i) const { return maxes.as<int>() + i; }
};

#define PS_CUDA(ctx, call)                                                                          \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return (ctx)->fail(PS_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// kernel classes for ps_profile_read / DESIGN.md
enum KClass {
  KC_PREP = 0, KC_MAX, KC_MASK, KC_ROTCONV, KC_WARP_DIRECT, KC_WARP_BILINEAR, KC_CONV_ROWS, KC_CONV_COLS,
  KC_WARP_BACK, KC_EPILOGUE, KC_ROOT_COMBINE, KC_ROOT_MARGINAL, KC_ARGMAX, KC_LOCAL_MAX, KC_MISC, KC_GAUSS_XY, KC_COUNT
};
static const char *const kClassNames[KC_COUNT] = {
    "prepare_unary", "grid_max", "mask", "rotconv", "warp_direct", "warp_bilinear", "conv_rows", "conv_cols",
    "warp_back", "epilogue", "root_combine", "root_marginal", "argmax", "local_max", "misc", "gauss_xy"};
constexpr int kWorkCounters = 1024;

static size_t prof_event(ps_ctx *c) {
  if (c->ev_used == c->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->ev_pool.push_back(e);
  }
  cudaEventRecord(c->ev_pool[c->ev_used], c->stream);
  return c->ev_used++;
}

// Launch a kernel of class `klass` on the ctx stream; with profiling on, bracket it with CUDA events.
#define PS_LAUNCH(ctx, klass, ...)                                                                  \
  do {                                                                                              \
    size_t e0_ = 0;                                                                                 \
    if ((ctx)->profiling) e0_ = prof_event(ctx);                                                    \
    __VA_ARGS__;                                                                                    \
    if ((ctx)->profiling) {                                                                         \
      size_t e1_ = prof_event(ctx);                                                                 \
      (ctx)->spans.push_back({(klass), e0_, e1_});                                                  \
    }                                                                                               \
    ++(ctx)->launches;                                                                              \
    cudaError_t e_ = cudaGetLastError();                                                            \
    if (e_ != cudaSuccess) return (ctx)->fail(PS_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e_)); \
  } while (0)

namespace {

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

psg::Grid grid_of(const ps_ctx *c) {
  psg::Grid g;
  g.R = c->R; g.H = c->H; g.W = c->W;
  g.min_rot = c->cfg.min_part_rotation;
  g.max_rot = c->cfg.max_part_rotation;
  return g;
}

double scale_of(const ps_config &cfg, int s) {
  return psg::value_from_index(cfg.min_object_scale, cfg.max_object_scale, cfg.num_scale_steps, s);
}

int upload_plan(ps_ctx *c, DevPlan &dp) {
  const psg::MessagePlan &h = dp.host;
  std::vector<int> ints;
  ints.insert(ints.end(), h.xin.begin(), h.xin.end());
  ints.insert(ints.end(), h.xout.begin(), h.xout.end());
  ints.insert(ints.end(), h.yin.begin(), h.yin.end());
  ints.insert(ints.end(), h.yout.begin(), h.yout.end());
  ints.insert(ints.end(), h.in_shift.begin(), h.in_shift.end());
  ints.insert(ints.end(), h.out_shift.begin(), h.out_shift.end());
  ints.insert(ints.end(), h.ytiles.begin(), h.ytiles.end());
  ints.insert(ints.end(), h.xtiles.begin(), h.xtiles.end());
  std::vector<float> fl;
  fl.insert(fl.end(), h.rot_taps.begin(), h.rot_taps.end());
  fl.insert(fl.end(), h.fx.begin(), h.fx.end());
  fl.insert(fl.end(), h.fy.begin(), h.fy.end());
  PS_CUDA(c, dp.ints.alloc(ints.size() * sizeof(int)));
  PS_CUDA(c, dp.floats.alloc(fl.size() * sizeof(float)));
  PS_CUDA(c, cudaMemcpyAsync(dp.ints.p, ints.data(), ints.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  PS_CUDA(c, cudaMemcpyAsync(dp.floats.p, fl.data(), fl.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  PS_CUDA(c, cudaStreamSynchronize(c->stream));  // host vectors go out of scope
  if (!h.diag) {
    dp.EP = (h.EW + 7) & ~7;
    double m[12];
    memcpy(m, h.T31, sizeof h.T31);
    memcpy(m + 6, h.T13, sizeof h.T13);
    PS_CUDA(c, dp.mats.alloc(sizeof m));
    PS_CUDA(c, cudaMemcpy(dp.mats.p, m, sizeof m, cudaMemcpyHostToDevice));
    const std::vector<int> &wl = c->disable_tile_lists ? h.walks_all : h.walks;
    const std::vector<unsigned char> &ml = c->disable_tile_lists ? h.xmasks_all : h.xmasks;
    dp.nwalks = (int)wl.size() / 4;
    PS_CUDA(c, dp.walks.alloc(wl.size() * sizeof(int) + ml.size()));
    if (!wl.empty()) {
      PS_CUDA(c, cudaMemcpy(dp.walks.p, wl.data(), wl.size() * sizeof(int), cudaMemcpyHostToDevice));
      PS_CUDA(c, cudaMemcpy((char *)dp.walks.p + wl.size() * sizeof(int), ml.data(), ml.size(), cudaMemcpyHostToDevice));
    }
  }
  return PS_OK;
}

int get_plan(ps_ctx *c, const double off_in[2], const double off_out[2], const double C[4], double rot_mean,
             double rot_sigma, double scale, std::shared_ptr<DevPlan> &out, const double *pos_offset = nullptr) {
  double keyv[14] = {off_in[0], off_in[1], off_out[0], off_out[1], C[0], C[1], C[2], C[3], rot_mean, rot_sigma, scale,
                     pos_offset ? 1.0 : 0.0, pos_offset ? pos_offset[0] : 0.0, pos_offset ? pos_offset[1] : 0.0};
  std::string key((const char *)keyv, sizeof keyv);
  auto it = c->plan_cache.find(key);
  if (it != c->plan_cache.end()) {
    ++c->plan_cache_hits;
    out = it->second;
    return PS_OK;
  }
  ++c->plan_cache_misses;
  std::shared_ptr<DevPlan> dp(new DevPlan);
  dp->host = psg::plan_message(grid_of(c), off_in, off_out, C, rot_mean, rot_sigma, scale, pos_offset);
  if (!dp->host.error.empty()) return c->fail(PS_ERR_INVALID, "%s", dp->host.error.c_str());
  int rc = upload_plan(c, *dp);
  if (rc) return rc;
  if (c->plan_cache.size() >= 2048) {  // drop what no joint set references any more
    PS_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto i = c->plan_cache.begin(); i != c->plan_cache.end();)
      i = i->second.use_count() == 1 ? c->plan_cache.erase(i) : std::next(i);
    // work-order tables and captured graphs are keyed by / hold plan addresses, which may now be reused
    c->gauss_tables.clear();
    for (auto &g : c->graphs)
      if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    c->graphs.clear();
  }
  c->plan_cache.emplace(key, dp);
  out = dp;
  return PS_OK;
}

size_t plan_scratch_elems(const ps_ctx *c, const DevPlan &dp) {
  if (dp.host.diag) return c->N;
  const size_t ehp = (dp.host.EH + 7) & ~7;
  return std::max(c->N, std::max((size_t)c->R * dp.host.EH * dp.EP, (size_t)c->R * dp.host.EW * ehp));
}

int ensure_scratch(ps_ctx *c, size_t elems) {
  if (elems <= c->scratch_elems) return PS_OK;
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  PS_CUDA(c, c->bufU.alloc(elems * sizeof(float)));
  PS_CUDA(c, c->bufV.alloc(elems * sizeof(float)));
  // the Gaussian work lists leave cells of the scratch grids unwritten; a later box may carry them into outputs
  // nobody reads -- give them a defined value once (keeps initcheck quiet and NaN bit patterns out of the pipes)
  PS_CUDA(c, cudaMemsetAsync(c->bufU.p, 0, elems * sizeof(float), c->stream));
  PS_CUDA(c, cudaMemsetAsync(c->bufV.p, 0, elems * sizeof(float), c->stream));
  c->scratch_elems = elems;
  return PS_OK;
}

int ensure_direct_map(ps_ctx *c, DevPlan &dp) {
  if (dp.map_ready) return PS_OK;
  const psg::MessagePlan &h = dp.host;
  PS_CUDA(c, dp.map.alloc((size_t)h.EH * h.EW * sizeof(int2)));
  int *ovf = c->counters.as<int>() + 7;
  PS_CUDA(c, cudaMemsetAsync(ovf, 0, sizeof(int), c->stream));
  dim3 b(32, 8), g(cdiv(h.EW, 32), cdiv(h.EH, 8));
  PS_LAUNCH(c, KC_MISC, psk::k_build_direct_map<<<g, b, 0, c->stream>>>(dp.map.as<int2>(), h.EH, h.EW, c->H, c->W, dp.mats.as<double>(),
                                                  dp.mats.as<double>() + 6, ovf));
  int hovf = 0;
  PS_CUDA(c, cudaMemcpyAsync(&hovf, ovf, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  if (hovf)
    return c->fail(PS_ERR_UNSUPPORTED,
                   "TM_DIRECT scatter map: an eigen-frame cell has more than two pre-images (transform.hpp:167-192)");
  dp.map_ready = true;
  return PS_OK;
}


constexpr size_t kSmemBudget = 96 * 1024;  // per block, leaves room for 2 blocks per SM

template <int Rr, int L, int OUT>
int launch_rotconv_t(ps_ctx *c, const psk::RotBatch &rb, int nmsg) {
  constexpr int PX = 128;
  static const bool old_kernel = getenv("PSINFER_ROTCONV3") != nullptr;  // A/B switch
  bool pure = true;
  for (int i = 0; i < nmsg; ++i) pure = pure && rb.a[i].shift_xy;
  const dim3 grid(cdiv(c->HW, PX), nmsg);
  if (pure && !old_kernel && c->cfg.fast_math)
    PS_LAUNCH(c, KC_ROTCONV, psk::k_rotconv4<Rr, L, PX, OUT, true><<<grid, 256, 0, c->stream>>>(
                                 rb, psk::FastDiv((unsigned)c->W), PS_NEGZERO2));
  else if (pure && !old_kernel)
    PS_LAUNCH(c, KC_ROTCONV, psk::k_rotconv4<Rr, L, PX, OUT, false><<<grid, 256, 0, c->stream>>>(
                                 rb, psk::FastDiv((unsigned)c->W), PS_NEGZERO2));
  else
    for (int i = 0; i < nmsg; ++i)
      PS_LAUNCH(c, KC_ROTCONV, psk::k_rotconv3<Rr, L, PX, OUT><<<cdiv(c->HW, PX), 256, 0, c->stream>>>(rb.a[i], PS_NEGZERO2));
  return PS_OK;
}

// Block-cooperative rotation filter for the common rotation counts, generic shared-memory kernel otherwise.
// All messages of the batch share one launch when every translation table is a pure shift.
int launch_rotconv(ps_ctx *c, const psk::RotBatch &rb, int nmsg) {
  int len = 1;
  for (int i = 0; i < nmsg; ++i) len = std::max(len, rb.a[i].mode == 1 ? rb.a[i].len : 1);
  const int R = rb.a[0].R;
#define ROT_CASE(Rr, L, OUT) \
  if (R == Rr && len <= L) return launch_rotconv_t<Rr, L, OUT>(c, rb, nmsg)
  ROT_CASE(8, 7, 4);
  ROT_CASE(12, 11, 6);
  ROT_CASE(24, 7, 6);
  ROT_CASE(24, 15, 6);
  ROT_CASE(24, 23, 6);
  ROT_CASE(48, 7, 8);
  ROT_CASE(48, 15, 8);
  ROT_CASE(48, 23, 8);
  ROT_CASE(48, 47, 8);
#undef ROT_CASE
  for (int i = 0; i < nmsg; ++i) {
    const psk::RotArgs &a = rb.a[i];
    size_t smem = (size_t)a.R * psk::kRotThreads * sizeof(float);
    if (smem > 48 * 1024)
      PS_CUDA(c, cudaFuncSetAttribute(psk::k_rotconv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PS_LAUNCH(c, KC_ROTCONV, psk::k_rotconv<<<cdiv(c->HW, psk::kRotThreads), psk::kRotThreads, smem, c->stream>>>(a));
  }
  return PS_OK;
}
int launch_rotconv(ps_ctx *c, const psk::RotArgs &a) {
  psk::RotBatch rb{};
  rb.a[0] = a;
  return launch_rotconv(c, rb, 1);
}

int launch_conv_rows(ps_ctx *c, const psk::ConvArgs &a, int slices) {
  constexpr int T = 8;
  const int n = (a.len - 1) / 2;
  const int G = (a.cols + T - 1) / T;
  {
    // v3: row pairs, packed arithmetic, persistent blocks with cp.async double buffering
    int S = (G * T + 2 * n + T - 1) / T + 1;
    while (S % 16 != 2) ++S;
    const size_t pair_bytes = (size_t)T * S * sizeof(float2);
    int best = 0;
    double best_eff = 0;
    for (int pairs = 1; pairs <= 32 && 2 * pairs * pair_bytes <= 100 * 1024; ++pairs) {
      int items = pairs * G;
      double eff = (double)items / (((items + 255) / 256) * 256);
      if (eff >= best_eff - 1e-9) { best_eff = eff; best = pairs; }
    }
    if (best > 0 && !c->disable_tma) {
      const int ytiles = (a.rows + 2 * best - 1) / (2 * best);
      const int ntiles = slices * ytiles;
      static const int bps_env = getenv("PSINFER_CONV_BLOCKS") ? atoi(getenv("PSINFER_CONV_BLOCKS")) : 0;  // A/B knob
  const int grid = std::min(ntiles, c->num_sms * (bps_env == 1 ? 1 : 2));
      PS_LAUNCH(c, KC_CONV_ROWS,
                psk::k_conv_rows3<T><<<grid, 256, 2 * best * pair_bytes, c->stream>>>(a, PS_NEGZERO2, best, S, slices, ytiles));
      return PS_OK;
    }
    best = 0; best_eff = 0;
    for (int pairs = 1; pairs <= 32 && pairs * pair_bytes <= 40 * 1024; ++pairs) {
      int items = pairs * G;
      double eff = (double)items / (((items + 255) / 256) * 256);
      if (eff >= best_eff - 1e-9) { best_eff = eff; best = pairs; }
    }
    if (best > 0) {
      size_t smem = best * pair_bytes;
      PS_LAUNCH(c, KC_CONV_ROWS,
                psk::k_conv_rows2<T><<<dim3(cdiv(a.rows, 2 * best), slices), 256, smem, c->stream>>>(a, PS_NEGZERO2, best, S));
      return PS_OK;
    }
  }
  // v1 fallback (very long filters)
  int S = (G * T + 2 * n + T - 1) / T + 1;
  while (S % 8 != 4) ++S;
  size_t row_bytes = (size_t)T * S * sizeof(float);
  if (row_bytes > kSmemBudget)
    return c->fail(PS_ERR_UNSUPPORTED, "row filter of %d taps over %d columns exceeds the shared-memory tile", a.len, a.cols);
  int TY = (int)std::max<size_t>(1, std::min<size_t>(4, kSmemBudget / row_bytes));
  PS_LAUNCH(c, KC_CONV_ROWS,
            psk::k_conv_rows<T><<<dim3(cdiv(a.rows, TY), slices), 256, TY * row_bytes, c->stream>>>(a, TY, S));
  return PS_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encoder() {
  // function-local static with an initialiser: thread-safe (several host threads create contexts at once)
  static const EncodeTiledFn fn = []() -> EncodeTiledFn {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return (EncodeTiledFn)p;
    return nullptr;
  }();
  return fn;
}

// Can the TMA column kernel take a filter of `len` taps over an input of this layout?
bool cols_tma_ok(const ps_ctx *c, const float *in, int len, int pitch, size_t plane) {
  const int nrows = 64 + (len - 1);
  const size_t stage = ((size_t)nrows * 64 * sizeof(float) + 127) & ~(size_t)127;
  return tensor_map_encoder() && !c->disable_tma && nrows <= 256 && 2 * stage <= 104 * 1024 && pitch % 4 == 0 &&
         plane % 4 == 0 && (uintptr_t)in % 16 == 0;
}

// TMA column filter.  Input: [slices][rows][in_pitch] with `cols` valid columns (filter along rows).
// Output: same orientation (transpose_out = 0, out pitch/plane as given) or transposed [slices][cols][out_pitch].
int launch_conv_cols_tma(ps_ctx *c, const float *in, int in_pitch, size_t in_plane, float *out, int out_pitch,
                         size_t out_plane, const float *taps, int len, int rows, int cols, int slices, int transpose_out,
                         const int *tile_list = nullptr, int ntile_list = 0) {
  constexpr int T = 8;
  const int n = (len - 1) / 2;
  const int nrows = 8 * T + 2 * n;
  const size_t stage = ((size_t)nrows * 64 * sizeof(float) + 127) & ~(size_t)127;
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slices};
  cuuint64_t strides[2] = {(cuuint64_t)in_pitch * sizeof(float), (cuuint64_t)in_plane * sizeof(float)};
  cuuint32_t box[3] = {64, (cuuint32_t)nrows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = tensor_map_encoder()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)in, dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return c->fail(PS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  psk::ColsTmaArgs t;
  t.out = out; t.taps = taps; t.len = len; t.rows = rows; t.cols = cols; t.pitch = out_pitch; t.plane = out_plane;
  t.slices = slices; t.ytiles = (rows + 8 * T - 1) / (8 * T); t.xtiles = (cols + 63) / 64;
  t.transpose_out = transpose_out;
  t.tile_list = (tile_list && ntile_list > 0 && !c->disable_tile_lists) ? tile_list : nullptr;
  t.ntile_list = ntile_list;
  const int ntiles = t.slices * (t.tile_list ? ntile_list : t.ytiles * t.xtiles);
  // Two blocks per SM, and never a third.  With images in flight on several streams the memory-bound kernels of other
  // images (resampling, epilogue, rotation filter) run beside the tap loops, and how much room they find decides the
  // bench (round-1 A/B, cfg-2, 8 streams): 88 registers, nothing else constrained: 424 img/s (one small block fits
  // beside two Gaussian blocks); Gaussian blocks that fill the SM's shared memory, nothing co-resident: 380; 67
  // registers and no cap: 383 -- a third Gaussian block moves in, 219 KB of shared memory leave the gathers of the
  // other kernels almost no L1; 67 / 56 registers with the dynamic shared memory padded so that three Gaussian blocks
  // cannot fit: 430.6 / 431.8 -- half of the register file and ~100 KB of L1 stay free for three or four small blocks.
  static const int bps_env = getenv("PSINFER_CONV_BLOCKS") ? atoi(getenv("PSINFER_CONV_BLOCKS")) : 0;  // A/B knob
  const int grid = std::min(ntiles, c->num_sms * (bps_env == 1 ? 1 : 2));
  {
    static const int ns_env = getenv("PSINFER_TMA_STAGES") ? atoi(getenv("PSINFER_TMA_STAGES")) : 0;
    int ns = 2;  // deeper queues (3, 4) measured no faster: the boxes already land a tile ahead
    if (ns_env >= 2) ns = (int)std::min<size_t>(std::min(ns_env, psk::kMaxTmaStages), (104 * 1024) / stage);
    // 3 * (dynamic + 8 KB static + 1 KB reserved) must exceed the 228 KB of an SM: 70 KB of dynamic shared memory at least
    static const size_t smem_env = getenv("PSINFER_CONV_SMEM") ? (size_t)atol(getenv("PSINFER_CONV_SMEM")) : 70 * 1024;
    const size_t smem = std::min<size_t>(std::max(ns * stage, smem_env), 104 * 1024);
    if (c->cfg.fast_math)
      PS_LAUNCH(c, transpose_out ? KC_CONV_ROWS : KC_CONV_COLS,
                psk::k_conv_cols_tma2<T, true><<<grid, 288, smem, c->stream>>>(tm, t, PS_NEGZERO2, ns));
    else
      PS_LAUNCH(c, transpose_out ? KC_CONV_ROWS : KC_CONV_COLS,
                psk::k_conv_cols_tma2<T, false><<<grid, 288, smem, c->stream>>>(tm, t, PS_NEGZERO2, ns));
  }
  return PS_OK;
}

int launch_conv_cols(ps_ctx *c, const psk::ConvArgs &a, int slices) {
  constexpr int T = 8;
  const int n = (a.len - 1) / 2;
  const int nrows = 8 * T + 2 * n;
  const bool aligned = (a.pitch % 2 == 0) && (a.plane % 2 == 0) && ((uintptr_t)a.in % 8 == 0) && ((uintptr_t)a.out % 8 == 0);
  if (aligned && cols_tma_ok(c, a.in, a.len, a.pitch, a.plane))
    return launch_conv_cols_tma(c, a.in, a.pitch, a.plane, a.out, a.pitch, a.plane, a.taps, a.len, a.rows, a.cols, slices, 0);
  const size_t smem = (size_t)nrows * 32 * sizeof(float2);
  if (aligned && smem <= kSmemBudget) {
    PS_LAUNCH(c, KC_CONV_COLS,
              psk::k_conv_cols2<T><<<dim3(cdiv(a.cols, 64), cdiv(a.rows, 8 * T), slices), 256, smem, c->stream>>>(a, PS_NEGZERO2));
    return PS_OK;
  }
  PS_LAUNCH(c, KC_CONV_COLS, psk::k_conv_cols<T><<<dim3(cdiv(a.cols, 128), cdiv(a.rows, T), slices), 128, 0, c->stream>>>(a));
  return PS_OK;
}

template <int NC>
int launch_root_combine_t(ps_ctx *c, const psk::RootArgs &a) {
  if (a.N % 4 == 0)
    PS_LAUNCH(c, KC_ROOT_COMBINE, psk::k_root_combine<NC, 4><<<cdiv(a.N / 4, 256), 256, 0, c->stream>>>(a));
  else
    PS_LAUNCH(c, KC_ROOT_COMBINE, psk::k_root_combine<NC, 1><<<cdiv(a.N, 256), 256, 0, c->stream>>>(a));
  return PS_OK;
}

int launch_root_combine(ps_ctx *c, const psk::RootArgs &a) {
  switch (a.n) {
#define RC_CASE(NC) case NC: return launch_root_combine_t<NC>(c, a)
    RC_CASE(1); RC_CASE(2); RC_CASE(3); RC_CASE(4); RC_CASE(5); RC_CASE(6); RC_CASE(7); RC_CASE(8);
    RC_CASE(9); RC_CASE(10); RC_CASE(11); RC_CASE(12); RC_CASE(13); RC_CASE(14); RC_CASE(15); RC_CASE(16);
#undef RC_CASE
  }
  return c->fail(PS_ERR_UNSUPPORTED, "root with %d children", a.n);
}

// Where the value of one message goes (the addGrid2 calls around computeRotJointMarginal).
struct Sink {
  float *out0 = nullptr;
  const float *acc0 = nullptr;
  const float *add0 = nullptr;
  float *out1 = nullptr;
  const float *add1 = nullptr;
  int *max0 = nullptr;
  int *max1 = nullptr;
  unsigned long long *amax0 = nullptr;
};

// One computeRotJointMarginal (findrot.cpp:292-456) on device buffers.  `in_max` holds enc(max(in)).
int run_message(ps_ctx *c, DevPlan &dp, const float *in, const int *in_max, bool sparse, const Sink &sink) {
  const psg::MessagePlan &h = dp.host;
  const int R = c->R, H = c->H, W = c->W;
  int rc = ensure_scratch(c, plan_scratch_elems(c, dp));
  if (rc) return rc;
  cudaStream_t st = c->stream;

  // stage 1: shift + exp + rotation filter -> bufB
  {
    psk::RotArgs a;
    a.in = in; a.out = c->bufB.as<float>();
    a.xin = dp.xin(); a.yin = dp.yin(R, W);
    a.shift_xy = dp.in_shift(R, H, W);
    a.taps = dp.rot_taps(); a.max_enc = in_max;
    a.R = R; a.H = H; a.W = W;
    a.shift = h.rot_shift; a.mode = h.rot_mode; a.len = (int)h.rot_taps.size();
    int rc2 = launch_rotconv(c, a);
    if (rc2) return rc2;
  }

  psk::EpiArgs e{};
  const float *filtered = nullptr;
  auto conv_rows = [&](const float *src, float *dst, int rows, int cols, int pitch, size_t plane, const float *taps,
                       int len) -> int {
    return launch_conv_rows(c, psk::ConvArgs{src, dst, taps, len, rows, cols, pitch, plane}, R);
  };
  auto conv_cols = [&](const float *src, float *dst, int rows, int cols, int pitch, size_t plane, const float *taps,
                       int len) -> int {
    return launch_conv_cols(c, psk::ConvArgs{src, dst, taps, len, rows, cols, pitch, plane}, R);
  };

  if (h.diag) {
    // gaussFilterDiag2d on the image grid: x then y
    rc = conv_rows(c->bufB.as<float>(), c->bufU.as<float>(), H, W, W, c->HW, dp.fx(), (int)h.fx.size());
    if (rc) return rc;
    rc = conv_cols(c->bufU.as<float>(), c->bufB.as<float>(), H, W, W, c->HW, dp.fy(), (int)h.fy.size());
    if (rc) return rc;
    filtered = c->bufB.as<float>();
    e.general = 0;
  } else {
    // gaussFilter2dOffset: rotate into the eigen-frame, filter there, bilinear read-back
    const int EH = h.EH, EW = h.EW, EP = dp.EP;
    const size_t eplane = (size_t)EH * EP;
    constexpr int RG = PS_RG;
    // Transposed route: the resampler writes the eigen-frame grid transposed ([r][x][y]) so that the x filter is a
    // TMA column filter too; its transposing store restores [r][y][x] for the y filter.
    const int EHP = (EH + 7) & ~7;
    const size_t tplane = (size_t)EW * EHP;
    const bool tr = cols_tma_ok(c, c->bufU.as<float>(), (int)h.fx.size(), EHP, tplane);
    if (sparse) {
      rc = ensure_direct_map(c, dp);
      if (rc) return rc;
      if (R % RG == 0)
        PS_LAUNCH(c, KC_WARP_DIRECT,
                  psk::k_warp_direct2<RG, true><<<dim3(cdiv(tr ? EW : EP, 16), cdiv(EH, 16), R / RG), dim3(16, 16), 0, st>>>(
                      c->bufB.as<float>(), c->bufU.as<float>(), dp.map.as<int2>(), R, c->HW, EH, EW, tr ? EHP : EP, tr));
      else
        PS_LAUNCH(c, KC_WARP_DIRECT,
                  psk::k_warp_direct2<RG, false><<<dim3(cdiv(tr ? EW : EP, 16), cdiv(EH, 16), cdiv(R, RG)), dim3(16, 16), 0, st>>>(
                      c->bufB.as<float>(), c->bufU.as<float>(), dp.map.as<int2>(), R, c->HW, EH, EW, tr ? EHP : EP, tr));
    } else {
      Affine T13;
      memcpy(T13.m, h.T13, sizeof T13.m);
      if (R % RG == 0)
        PS_LAUNCH(c, KC_WARP_BILINEAR,
                  psk::k_resample_bilinear<RG, true><<<dim3(cdiv(tr ? EW : EP, 16), cdiv(EH, 16), R / RG), dim3(16, 16), 0, st>>>(
                      c->bufB.as<float>(), c->bufU.as<float>(), T13, R, H, W, W, c->HW, EH, EW, tr ? EHP : EP,
                      tr ? tplane : eplane, tr));
      else
        PS_LAUNCH(c, KC_WARP_BILINEAR,
                  psk::k_resample_bilinear<RG, false><<<dim3(cdiv(tr ? EW : EP, 16), cdiv(EH, 16), cdiv(R, RG)), dim3(16, 16), 0, st>>>(
                      c->bufB.as<float>(), c->bufU.as<float>(), T13, R, H, W, W, c->HW, EH, EW, tr ? EHP : EP,
                      tr ? tplane : eplane, tr));
    }
    if (tr) {
      rc = launch_conv_cols_tma(c, c->bufU.as<float>(), EHP, tplane, c->bufV.as<float>(), EP, eplane, dp.fx(),
                                (int)h.fx.size(), /*rows=*/EW, /*cols=*/EH, R, /*transpose_out=*/1, dp.xtiles(R, H, W),
                                (int)h.xtiles.size() / 2);
    } else {
      rc = conv_rows(c->bufU.as<float>(), c->bufV.as<float>(), EH, EW, EP, eplane, dp.fx(), (int)h.fx.size());
    }
    if (rc) return rc;
    if (tr && cols_tma_ok(c, c->bufV.as<float>(), (int)h.fy.size(), EP, eplane))
      rc = launch_conv_cols_tma(c, c->bufV.as<float>(), EP, eplane, c->bufU.as<float>(), EP, eplane, dp.fy(),
                                (int)h.fy.size(), EH, EW, R, 0, dp.ytiles(R, H, W), (int)h.ytiles.size() / 2);
    else
      rc = conv_cols(c->bufV.as<float>(), c->bufU.as<float>(), EH, EW, EP, eplane, dp.fy(), (int)h.fy.size());
    if (rc) return rc;
    // Bilinear read-back into the image frame (filter.hpp:367-368), then the generic epilogue.  (A fused
    // read-back + epilogue over source cells was tried in round 1: 63 us per message against 22 + 28 us for the
    // two kernels -- it loses the 128-bit accumulator/operand traffic of k_epilogue2.  Moving only log(D) + M into
    // the read-back, to hide the fp64 log behind its gathers, was measured too: epilogue 0.46 -> 0.32 ms per image,
    // read-back 0.39 -> 0.55 ms, 409 -> 403 images/s -- the log costs the same ~8 us per message wherever it runs.)
    Affine T34;
    memcpy(T34.m, h.T34, sizeof T34.m);
    if (R % RG == 0)
      PS_LAUNCH(c, KC_WARP_BACK,
                psk::k_resample_bilinear<RG, true><<<dim3(cdiv(W, 16), cdiv(H, 16), R / RG), dim3(16, 16), 0, st>>>(
                    c->bufU.as<float>(), c->bufB.as<float>(), T34, R, EH, EW, EP, eplane, H, W, W, c->HW, 0));
    else
      PS_LAUNCH(c, KC_WARP_BACK,
                psk::k_resample_bilinear<RG, false><<<dim3(cdiv(W, 16), cdiv(H, 16), cdiv(R, RG)), dim3(16, 16), 0, st>>>(
                    c->bufU.as<float>(), c->bufB.as<float>(), T34, R, EH, EW, EP, eplane, H, W, W, c->HW, 0));
    filtered = c->bufB.as<float>();
    e.general = 0;
  }

  // stage 3: log, +M, shift, combine
  e.src = filtered;
  e.xout = dp.xout(R, H, W); e.yout = dp.yout(R, H, W);
  e.shift_xy = dp.out_shift(R, H, W);
  e.max_enc = in_max;
  e.R = R; e.H = H; e.W = W;
  e.out0 = sink.out0; e.acc0 = sink.acc0; e.add0 = sink.add0;
  e.out1 = sink.out1; e.add1 = sink.add1;
  e.max0 = sink.max0; e.max1 = sink.max1;
  e.amax0 = sink.amax0;
  {
    const int XG = (W + 3) / 4;
    // the 128-bit paths need 16-byte aligned caller buffers (cudaMalloc'd grids always are)
    const bool al = ((uintptr_t)e.out0 % 16 == 0) && ((uintptr_t)e.acc0 % 16 == 0) && ((uintptr_t)e.add0 % 16 == 0) &&
                    ((uintptr_t)e.out1 % 16 == 0) && ((uintptr_t)e.add1 % 16 == 0) && ((uintptr_t)e.xout % 16 == 0);
    if (!al) {
      PS_LAUNCH(c, KC_EPILOGUE, psk::k_epilogue<<<dim3(cdiv(W, 256), H, R), 256, 0, st>>>(e));
    } else if (!e.general && e.shift_xy && W % 4 == 0 && (uintptr_t)e.src % 16 == 0) {
      psk::EpiBatch eb{};
      eb.a[0] = e;
      PS_LAUNCH(c, KC_EPILOGUE, psk::k_epilogue3<<<dim3(cdiv((size_t)XG * H, 256), R, 1), 256, 0, st>>>(eb, psk::FastDiv((unsigned)XG)));
    } else if (e.general) {
      PS_LAUNCH(c, KC_EPILOGUE, psk::k_epilogue2<true><<<dim3(cdiv((size_t)XG * H, 256), R), 256, 0, st>>>(e, XG));
    } else {
      PS_LAUNCH(c, KC_EPILOGUE, psk::k_epilogue2<false><<<dim3(cdiv((size_t)XG * H, 256), R), 256, 0, st>>>(e, XG));
    }
  }
  return PS_OK;
}

// ---- level-batched schedule ---------------------------------------------------------------------------------------
// The messages of one tree level (findrot.cpp:582-658 upward, :158-236 downward) are independent; those that take the
// eigen-frame route share five launches: rotation filter, resampling into the (transposed) eigen-frame, the fused
// x+y Gaussian, the read-back and the epilogue, each indexed by message.
struct MsgJob {
  DevPlan *dp;
  const float *in;
  const int *in_max;
  bool sparse;
  Sink sink;
};

size_t fused_smem_bytes(const psg::MessagePlan &h, unsigned &stage) {
  const int nx = ((int)h.fx.size() - 1) / 2;
  stage = (unsigned)(64 + 2 * nx) * 64u * sizeof(float);
  return 2 * (size_t)stage + (size_t)64 * (h.lag + 1) * psk::kRingPitch * sizeof(float);
}

constexpr size_t kFusedSmemMax = 200 * 1024;

bool can_batch(const ps_ctx *c, const MsgJob &j) {
  const psg::MessagePlan &h = j.dp->host;
  if (c->disable_batch || c->disable_tma || h.diag || !tensor_map_encoder()) return false;
  if (c->R % PS_RG != 0 || c->W % 4 != 0 || !h.in_pure || !h.out_pure || j.dp->nwalks == 0) return false;
  if ((int)h.fx.size() > 193 || (int)h.fy.size() > psk::kMaxFusedTaps) return false;
  if (h.fx.size() % 2 == 0 || h.fy.size() % 2 == 0) return false;  // k_gauss_xy's tail code takes 2 n + 1 taps
  unsigned stage;
  if (fused_smem_bytes(h, stage) > kFusedSmemMax) return false;
  const Sink &k = j.sink;
  return ((uintptr_t)k.out0 % 16 == 0) && ((uintptr_t)k.acc0 % 16 == 0) && ((uintptr_t)k.add0 % 16 == 0) &&
         ((uintptr_t)k.out1 % 16 == 0) && ((uintptr_t)k.add1 % 16 == 0) && ((uintptr_t)j.in % 16 == 0);
}

int ensure_slots(ps_ctx *c, size_t nslots, size_t elems) {
  if (nslots <= c->slotB.size() && elems <= c->slot_elems) return PS_OK;
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  nslots = std::max(nslots, c->slotB.size());
  elems = std::max(elems, c->slot_elems);
  const bool grow = elems > c->slot_elems;
  const size_t old = c->slotB.size();
  if (nslots > old) {
    c->slotB.resize(nslots);
    c->slotU.resize(nslots);
    c->slotV.resize(nslots);
  }
  for (size_t i = grow ? 0 : old; i < nslots; ++i) {
    if (!c->slotB[i].p) PS_CUDA(c, c->slotB[i].alloc(c->N * sizeof(float)));
    PS_CUDA(c, c->slotU[i].alloc(elems * sizeof(float)));
    PS_CUDA(c, c->slotV[i].alloc(elems * sizeof(float)));
    // cells outside the work lists are never written; give them a defined value once (see ensure_scratch)
    PS_CUDA(c, cudaMemsetAsync(c->slotU[i].p, 0, elems * sizeof(float), c->stream));
    PS_CUDA(c, cudaMemsetAsync(c->slotV[i].p, 0, elems * sizeof(float), c->stream));
  }
  c->slot_elems = elems;
  return PS_OK;
}

int take_work_counter(ps_ctx *c, unsigned **out) {
  if (!c->work_counters.p) {
    PS_CUDA(c, c->work_counters.alloc(kWorkCounters * sizeof(unsigned)));
    c->work_counter_next = kWorkCounters;
  }
  if (c->work_counter_next >= kWorkCounters) {  // stream-ordered behind every launch that still uses an old counter
    PS_CUDA(c, cudaMemsetAsync(c->work_counters.p, 0, kWorkCounters * sizeof(unsigned), c->stream));
    c->work_counter_next = 0;
  }
  *out = c->work_counters.as<unsigned>() + c->work_counter_next++;
  return PS_OK;
}

int run_batch(ps_ctx *c, MsgJob *const *jobs, int n) {
  const int R = c->R, H = c->H, W = c->W;
  constexpr int RG = PS_RG;
  cudaStream_t st = c->stream;
  size_t elems = c->N;
  for (int i = 0; i < n; ++i) elems = std::max(elems, plan_scratch_elems(c, *jobs[i]->dp));
  int rc = ensure_slots(c, (size_t)n, elems);
  if (rc) return rc;
  const bool sparse = jobs[0]->sparse;
  // "Snake" order over the messages of a launch: blocks are handed out in index order, so a stage that walks the
  // messages in the opposite order of the stage before it starts on the data that was written last and is still in L2
  // (a level's intermediates are 5 x 23-37 MB against 126 MB of L2).  Rotation filter and Gaussian ascending, both
  // resampling stages descending, epilogue ascending.  PSINFER_NO_SNAKE=1: every stage ascending (A/B).
  static const bool snake = getenv("PSINFER_NO_SNAKE") == nullptr;

  // stage 1: shift + exp + rotation filter -> B[i]
  {
    psk::RotBatch rb{};
    for (int i = 0; i < n; ++i) {
      const DevPlan &dp = *jobs[i]->dp;
      psk::RotArgs &a = rb.a[i];
      a.in = jobs[i]->in; a.out = c->slotB[i].as<float>();
      a.xin = dp.xin(); a.yin = dp.yin(R, W);
      a.shift_xy = dp.in_shift(R, H, W);
      a.taps = dp.rot_taps(); a.max_enc = jobs[i]->in_max;
      a.R = R; a.H = H; a.W = W;
      a.shift = dp.host.rot_shift; a.mode = dp.host.rot_mode; a.len = (int)dp.host.rot_taps.size();
    }
    if ((rc = launch_rotconv(c, rb, n))) return rc;
  }
  // stage 2: into the eigen-frame, stored transposed [r][ex][ey]
  int maxEW = 0, maxEH = 0;
  for (int i = 0; i < n; ++i) {
    maxEW = std::max(maxEW, jobs[i]->dp->host.EW);
    maxEH = std::max(maxEH, jobs[i]->dp->host.EH);
  }
  if (sparse) {
    psk::DirectBatch db{};
    db.HW = c->HW; db.R = R; db.tr = 1; db.zgroups = R / RG;
    for (int i = 0; i < n; ++i) {
      DevPlan &dp = *jobs[i]->dp;
      if ((rc = ensure_direct_map(c, dp))) return rc;
      psk::DirectMsg &m = db.m[snake ? n - 1 - i : i];
      m.in = c->slotB[i].as<float>(); m.out = c->slotU[i].as<float>(); m.map = dp.map.as<int2>();
      m.EH = dp.host.EH; m.EW = dp.host.EW; m.EP = (dp.host.EH + 7) & ~7;
    }
    PS_LAUNCH(c, KC_WARP_DIRECT, psk::k_warp_direct_b<RG><<<dim3(cdiv(maxEW, 16), cdiv(maxEH, 16), n * (R / RG)), dim3(16, 16), 0, st>>>(db));
  } else {
    psk::ResampleBatch rb{};
    rb.R = R; rb.tr = 1; rb.zgroups = R / RG;
    for (int i = 0; i < n; ++i) {
      const psg::MessagePlan &h = jobs[i]->dp->host;
      const int EHP = (h.EH + 7) & ~7;
      psk::ResampleMsg &m = rb.m[snake ? n - 1 - i : i];
      m.src = c->slotB[i].as<float>(); m.dst = c->slotU[i].as<float>();
      memcpy(m.T.m, h.T13, sizeof m.T.m);
      m.sh = H; m.sw = W; m.spitch = W; m.splane = c->HW;
      m.dh = h.EH; m.dw = h.EW; m.dpitch = EHP; m.dplane = (size_t)h.EW * EHP;
    }
    PS_LAUNCH(c, KC_WARP_BILINEAR, psk::k_resample_bilinear_b<RG><<<dim3(cdiv(maxEW, 16), cdiv(maxEH, 16), n * (R / RG)), dim3(16, 16), 0, st>>>(rb));
  }
  // stage 3: both Gaussian passes, UT[i] -> V[i] ([r][ey][ex])
  {
    psk::TmapBatch tm;
    memset(&tm, 0, sizeof tm);
    psk::GaussBatch gb{};
    gb.nmsg = n; gb.R = R;
    size_t smem = 0;
    unsigned stride = 0;
    int lagmax = 1, items = 0;
    for (int i = 0; i < n; ++i) {
      const DevPlan &dp = *jobs[i]->dp;
      const psg::MessagePlan &h = dp.host;
      const int EHP = (h.EH + 7) & ~7;
      const int nx = ((int)h.fx.size() - 1) / 2;
      cuuint64_t dims[3] = {(cuuint64_t)h.EH, (cuuint64_t)h.EW, (cuuint64_t)R};
      cuuint64_t strides[2] = {(cuuint64_t)EHP * sizeof(float), (cuuint64_t)h.EW * EHP * sizeof(float)};
      cuuint32_t box[3] = {64, (cuuint32_t)(64 + 2 * nx), 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = tensor_map_encoder()(&tm.t[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, c->slotU[i].p, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return c->fail(PS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
      psk::GaussMsg &g = gb.m[i];
      g.out = c->slotV[i].as<float>();
      g.taps_x = dp.fx(); g.taps_y = dp.fy();
      g.walks = dp.walks.as<int>();
      g.masks = (const unsigned char *)(dp.walks.as<int>() + 4 * (size_t)dp.nwalks);
      g.oplane = (size_t)h.EH * dp.EP;
      g.len_x = (int)h.fx.size(); g.len_y = (int)h.fy.size();
      g.EH = h.EH; g.EW = h.EW; g.EP = dp.EP; g.nwalks = dp.nwalks; g.lag = h.lag; g.halo = h.halo;
      g.item0 = items;
      items += dp.nwalks * R;
      unsigned stage;
      fused_smem_bytes(h, stage);
      stride = std::max(stride, (stage + 127u) & ~127u);
      lagmax = std::max(lagmax, h.lag);
    }
    gb.total_items = items;
    gb.stage_stride = stride;
    gb.ring_rows = 64 * (lagmax + 1);
    if ((rc = take_work_counter(c, &gb.counter))) return rc;
    // Footprint of the Gaussian blocks = what else fits on an SM while the taps run (A/B on the B200, cfg-2, see
    // DESIGN.md 5).  Two blocks per SM give the tap loops four warps per scheduler; at 92 registers and two TMA stages
    // they fill the SM and nothing of another image's memory-bound kernels co-resides; one block per SM leaves room but
    // only two warps per scheduler to cover the FFMA2 -> FADD2 latency (fp32 pipe 62 % busy).  Shipped: two blocks of
    // <= 72 registers with ONE TMA stage each (the next box is requested when the x phase ends and lands during the y
    // phase) and the dynamic shared memory padded so that a third block cannot fit.
    // PSINFER_GAUSS_BPS (1 | 2), PSINFER_GAUSS_STAGES (1 | 2), PSINFER_GAUSS_SMEM (minimum dynamic bytes) are the knobs.
    static const int bps_env = getenv("PSINFER_GAUSS_BPS") ? atoi(getenv("PSINFER_GAUSS_BPS")) : 2;
    static const int ns_env = getenv("PSINFER_GAUSS_STAGES") ? atoi(getenv("PSINFER_GAUSS_STAGES")) : 1;
    static const long pad_env = getenv("PSINFER_GAUSS_SMEM") ? atol(getenv("PSINFER_GAUSS_SMEM")) : -1;
    gb.stages = ns_env >= 2 ? 2 : 1;
    smem = (size_t)gb.stages * stride + (size_t)gb.ring_rows * psk::kRingPitch * sizeof(float);
    if (smem > kFusedSmemMax) {  // two stages of the widest box do not fit: one does (can_batch checked the 2-stage size per message)
      gb.stages = 1;
      smem = (size_t)stride + (size_t)gb.ring_rows * psk::kRingPitch * sizeof(float);
    }
    int bps = (bps_env >= 2 && smem <= 108 * 1024) ? 2 : 1;
    // cap the blocks per SM at `bps` for launches of OTHER images too: (bps + 1) blocks must not fit into 227 KB
    const size_t cap_pad = bps == 2 ? 72 * 1024 : 116 * 1024;
    const size_t pad = pad_env >= 0 ? (size_t)pad_env : cap_pad;
    smem = std::min<size_t>(std::max(smem, pad), kFusedSmemMax);
    const int grid = std::min(items, c->num_sms * bps);
    static const unsigned hint_env = getenv("PSINFER_MBAR_HINT") ? (unsigned)atol(getenv("PSINFER_MBAR_HINT")) : PS_MBAR_HINT_NS;
    gb.hint_ns = hint_env;
    // Fixed work order (default): the message-major item list -- the order that keeps a message's boxes in L2 -- dealt to
    // the least loaded block, costs = tap-outputs of the item (list scheduling: what drawing from a counter does at run
    // time, decided here once per batch).  Tables are cached per (plans, grid); a poselet-conditioned run whose joints
    // change with every image fills the cache and then draws from the counter.
    // PSINFER_GAUSS_DYNAMIC=1: work items from an atomic counter instead of the fixed order (A/B)
    static const bool dyn_flag = getenv("PSINFER_GAUSS_DYNAMIC") && atoi(getenv("PSINFER_GAUSS_DYNAMIC")) != 0;
    bool use_table = false;
    if (!dyn_flag) {
      std::string key((const char *)&grid, sizeof grid);
      for (int i = 0; i < n; ++i) {
        const DevPlan *p = jobs[i]->dp;
        key.append((const char *)&p, sizeof p);
      }
      auto it = c->gauss_tables.find(key);
      // build a table only for a joint set that is being reused (or the very first one): a run that swaps joints with every
      // image would pay a host-side build and a stream synchronisation per level for tables it never sees again
      if (it == c->gauss_tables.end() && c->gauss_tables.size() < 256 && c->joints_reused) {
        std::vector<long long> load(grid, 0);
        std::vector<std::vector<int>> per(grid);
        // min-heap over (load, block)
        std::vector<std::pair<long long, int>> heap;
        for (int g = 0; g < grid; ++g) heap.push_back({0, g});
        auto cmp = [](const std::pair<long long, int> &a, const std::pair<long long, int> &b) { return a > b; };
        std::make_heap(heap.begin(), heap.end(), cmp);
        bool fits = R <= 0xfff;
        for (int i = 0; i < n && fits; ++i) {
          const psg::MessagePlan &h = jobs[i]->dp->host;
          const std::vector<int> &wl = c->disable_tile_lists ? h.walks_all : h.walks;
          const std::vector<unsigned char> &ml = c->disable_tile_lists ? h.xmasks_all : h.xmasks;
          const int nw = (int)wl.size() / 4;
          if (nw > 0xfff) fits = false;
          for (int w = 0; w < nw && fits; ++w) {
            const int ng = wl[4 * w + 2], moff = wl[4 * w + 3], nxb = (ng + 7) / 8 + h.lag;
            long long cost = (long long)ng * 8 * 64 * (long long)h.fy.size();
            for (int x = 0; x < nxb; ++x) cost += (long long)__builtin_popcount(ml[moff + x]) * 8 * 64 * (long long)h.fx.size();
            cost += 20000LL * nxb;  // per-step overhead (barriers, stores)
            for (int z = 0; z < R; ++z) {
              std::pop_heap(heap.begin(), heap.end(), cmp);
              std::pair<long long, int> &top = heap.back();
              per[top.second].push_back((i << 24) | (w << 12) | z);
              top.first += cost;
              std::push_heap(heap.begin(), heap.end(), cmp);
            }
          }
        }
        if (fits) {
          std::vector<int> flat(grid + 1, 0);
          for (int g = 0; g < grid; ++g) flat[g + 1] = flat[g] + (int)per[g].size();
          for (int g = 0; g < grid; ++g) flat.insert(flat.end(), per[g].begin(), per[g].end());
          std::shared_ptr<DevBuf> tb(new DevBuf);
          PS_CUDA(c, tb->alloc(flat.size() * sizeof(int)));
          PS_CUDA(c, cudaMemcpyAsync(tb->p, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice, st));
          PS_CUDA(c, cudaStreamSynchronize(st));  // `flat` goes out of scope; happens once per batch shape
          it = c->gauss_tables.emplace(key, tb).first;
        }
      }
      if (it != c->gauss_tables.end()) {
        gb.cta_off = it->second->as<int>();
        gb.cta_items = it->second->as<int>() + grid + 1;
        use_table = true;
      }
    }
    const bool dyn_env = !use_table;

    // The taps are the one fp32-bound stage; everything else is memory-bound.  With several images in flight the block
    // scheduler should hand freed SM slots to a waiting Gaussian launch first, so that it holds its (capped) share of every
    // SM for its whole run and the memory-bound blocks of the other images fill the rest: launch priority, not a separate
    // stream.  PSINFER_GAUSS_PRIO=0 launches at the stream's priority (A/B).
    static const bool prio_env = !(getenv("PSINFER_GAUSS_PRIO") && atoi(getenv("PSINFER_GAUSS_PRIO")) == 0);
    static const int prio_hi = []() {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      return hi;
    }();
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(dyn_env ? 288 : 256);  // the counter-drawn order keeps its producer warp, the fixed order has none
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute la[1];
    la[0].id = cudaLaunchAttributePriority;
    la[0].val.priority = prio_hi;
    lc.attrs = la;
    lc.numAttrs = prio_env ? 1 : 0;
    const psk::u64 nz2 = PS_NEGZERO2;
    if (c->cfg.fast_math && dyn_env)
      PS_LAUNCH(c, KC_GAUSS_XY, cudaLaunchKernelEx(&lc, psk::k_gauss_xy<true, true>, tm, gb, nz2));
    else if (c->cfg.fast_math)
      PS_LAUNCH(c, KC_GAUSS_XY, cudaLaunchKernelEx(&lc, psk::k_gauss_xy<true, false>, tm, gb, nz2));
    else if (dyn_env)
      PS_LAUNCH(c, KC_GAUSS_XY, cudaLaunchKernelEx(&lc, psk::k_gauss_xy<false, true>, tm, gb, nz2));
    else
      PS_LAUNCH(c, KC_GAUSS_XY, cudaLaunchKernelEx(&lc, psk::k_gauss_xy<false, false>, tm, gb, nz2));
  }
  // stage 4: bilinear read-back into the image frame, V[i] -> B[i]
  {
    psk::ResampleBatch rb{};
    rb.R = R; rb.tr = 0; rb.zgroups = R / RG;
    for (int i = 0; i < n; ++i) {
      const DevPlan &dp = *jobs[i]->dp;
      const psg::MessagePlan &h = dp.host;
      psk::ResampleMsg &m = rb.m[snake ? n - 1 - i : i];
      m.src = c->slotV[i].as<float>(); m.dst = c->slotB[i].as<float>();
      memcpy(m.T.m, h.T34, sizeof m.T.m);
      m.sh = h.EH; m.sw = h.EW; m.spitch = dp.EP; m.splane = (size_t)h.EH * dp.EP;
      m.dh = H; m.dw = W; m.dpitch = W; m.dplane = c->HW;
    }
    PS_LAUNCH(c, KC_WARP_BACK, psk::k_resample_bilinear_b<RG><<<dim3(cdiv(W, 16), cdiv(H, 16), n * (R / RG)), dim3(16, 16), 0, st>>>(rb));
  }
  // stage 5: log, +M, shift, combine
  {
    psk::EpiBatch eb{};
    for (int i = 0; i < n; ++i) {
      const DevPlan &dp = *jobs[i]->dp;
      const Sink &sink = jobs[i]->sink;
      psk::EpiArgs &e = eb.a[i];
      e.src = c->slotB[i].as<float>();
      e.general = 0;
      e.xout = dp.xout(R, H, W); e.yout = dp.yout(R, H, W);
      e.shift_xy = dp.out_shift(R, H, W);
      e.max_enc = jobs[i]->in_max;
      e.R = R; e.H = H; e.W = W;
      e.out0 = sink.out0; e.acc0 = sink.acc0; e.add0 = sink.add0;
      e.out1 = sink.out1; e.add1 = sink.add1;
      e.max0 = sink.max0; e.max1 = sink.max1;
      e.amax0 = sink.amax0;
    }
    const int XG = W / 4;
    PS_LAUNCH(c, KC_EPILOGUE, psk::k_epilogue3<<<dim3(cdiv((size_t)XG * H, 256), R, n), 256, 0, st>>>(eb, psk::FastDiv((unsigned)XG)));
  }
  return PS_OK;
}

// All messages of one tree level.  Messages the batched route cannot take (diagonal covariance, unaligned or
// non-shift cases, filters beyond the shared-memory ring) go through run_message one by one.
int run_level(ps_ctx *c, std::vector<MsgJob> &jobs) {
  std::vector<MsgJob *> batch;
  for (MsgJob &j : jobs) {
    if (can_batch(c, j)) {
      batch.push_back(&j);
    } else {
      int rc = run_message(c, *j.dp, j.in, j.in_max, j.sparse, j.sink);
      if (rc) return rc;
    }
  }
  for (size_t i = 0; i < batch.size();) {
    size_t k = i + 1;
    // PSINFER_MAX_BATCH (1..8) caps the messages per launch: an A/B knob between fewer launches and a smaller L2 footprint
    static const int cap_env = getenv("PSINFER_MAX_BATCH") ? atoi(getenv("PSINFER_MAX_BATCH")) : psk::kMaxBatch;
    const size_t cap = (size_t)std::max(1, std::min(cap_env, psk::kMaxBatch));
    while (k < batch.size() && k - i < cap && batch[k]->sparse == batch[i]->sparse) ++k;
    int rc = run_batch(c, &batch[i], (int)(k - i));
    if (rc) return rc;
    i = k;
  }
  return PS_OK;
}

int reset_max(ps_ctx *c, int *slot) {
  PS_LAUNCH(c, KC_MISC, psk::k_set_int<<<1, 32, 0, c->stream>>>(slot, 1, PS_ENC_NEG_INF));
  return PS_OK;
}

int grid_max(ps_ctx *c, const float *g, size_t n, int *slot) {
  int rc = reset_max(c, slot);
  if (rc) return rc;
  PS_LAUNCH(c, KC_MAX, psk::k_grid_max<<<std::min(cdiv(n / 4 + 1, 256), (unsigned)c->num_sms * 16), 256, 0, c->stream>>>(g, n, slot));
  return PS_OK;
}

// Fills the device tables of exp_fast / log_fast (ps_kernels.cuh) with host-libm doubles.
int math_tables_init(ps_ctx *c) {
  // the tables are per device and never change: fill them once per device (other contexts' kernels may be reading them)
  static std::mutex mu;
  static std::vector<char> done(64, 0);
  std::lock_guard<std::mutex> lock(mu);
  if (c->cfg.device < (int)done.size() && done[c->cfg.device]) return PS_OK;
  double2 lt[129];
  for (int j = 0; j <= 128; ++j) {
    double F = 1.0 + j / 128.0;
    lt[j].x = std::ldexp(1.0 / F, -23);
    lt[j].y = j == 128 ? 0.0 : (j >= 54 ? std::log(F * 0.5) : std::log(F));
  }
  static double et[512];
  for (int b = 0; b < 512; ++b) et[b] = (b - 127) * 0.693147180559945309417232121458;
  double xt[64];
  for (int j = 0; j < 64; ++j) xt[j] = std::exp2(j / 64.0);
  PS_CUDA(c, cudaMemcpyToSymbol(psk::d_log_tab, lt, sizeof lt));
  PS_CUDA(c, cudaMemcpyToSymbol(psk::d_eln2_tab, et, sizeof et));
  PS_CUDA(c, cudaMemcpyToSymbol(psk::d_exp_tab, xt, sizeof xt));
  if (c->cfg.device < (int)done.size()) done[c->cfg.device] = 1;
  return PS_OK;
}

int validate_config(const ps_config *cfg, std::string &why) {
  char buf[256];
#define BAD(...)                         \
  do {                                   \
    snprintf(buf, sizeof buf, __VA_ARGS__); \
    why = buf;                           \
    return PS_ERR_INVALID;               \
  } while (0)
  if (cfg->num_parts < 1 || cfg->num_parts > PS_MAX_PARTS) BAD("num_parts %d out of [1,%d]", cfg->num_parts, PS_MAX_PARTS);
  if (cfg->num_rotation_steps < 1 || cfg->num_rotation_steps > 512) BAD("num_rotation_steps %d out of [1,512]", cfg->num_rotation_steps);
  if (cfg->num_scale_steps < 1) BAD("num_scale_steps %d < 1", cfg->num_scale_steps);
  if (cfg->height < 1 || cfg->width < 1) BAD("grid %dx%d is empty", cfg->height, cfg->width);
  if ((size_t)cfg->num_rotation_steps * cfg->height * cfg->width >= (size_t)1 << 31)
    BAD("R*H*W does not fit the reference's int flat index (findrot.cpp:266)");
  if (cfg->min_part_rotation == cfg->max_part_rotation && cfg->num_rotation_steps != 1)
    BAD("min_part_rotation == max_part_rotation needs num_rotation_steps == 1 (partapp_aux.hpp:29-31)");
  if (cfg->min_object_scale == cfg->max_object_scale && cfg->num_scale_steps != 1)
    BAD("min_object_scale == max_object_scale needs num_scale_steps == 1 (partapp_aux.hpp:29-31)");
  if (cfg->strip_border_detections >= 0.5f) BAD("strip_border_detections must be < 0.5 (findrot.cpp:529)");
  if (cfg->roi_save_num_samples < 0) BAD("roi_save_num_samples < 0");
#undef BAD
  return PS_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

const char *ps_version(void) { return "psinfer 0.1 sm_100a"; }

const char *ps_last_error(const ps_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

long long ps_launch_count(const ps_ctx *ctx) { return ctx ? ctx->launches : 0; }

int ps_profile_enable(ps_ctx *c, int on) {
  if (!c) return PS_ERR_INVALID;
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  c->profiling = on != 0;
  c->spans.clear();
  c->ev_used = 0;
  return PS_OK;
}

int ps_profile_read(ps_ctx *c, int cap, const char **names, double *total_ms, long long *launches, int *count) {
  if (!c || !names || !total_ms || !launches || !count) return PS_ERR_INVALID;
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  double ms[KC_COUNT] = {0};
  long long n[KC_COUNT] = {0};
  for (const ps_ctx::Span &sp : c->spans) {
    float t = 0;
    PS_CUDA(c, cudaEventElapsedTime(&t, c->ev_pool[sp.e0], c->ev_pool[sp.e1]));
    ms[sp.klass] += t;
    ++n[sp.klass];
  }
  int k = 0;
  for (int i = 0; i < KC_COUNT && k < cap; ++i)
    if (n[i]) {
      names[k] = kClassNames[i];
      total_ms[k] = ms[i];
      launches[k] = n[i];
      ++k;
    }
  *count = k;
  c->spans.clear();
  c->ev_used = 0;
  return PS_OK;
}

double ps_rot_from_index(const ps_config *cfg, int idx) {
  return psg::value_from_index(cfg->min_part_rotation, cfg->max_part_rotation, cfg->num_rotation_steps, idx);
}
double ps_scale_from_index(const ps_config *cfg, int idx) { return scale_of(*cfg, idx); }
int ps_plan_work_lists(const ps_config *cfg, const double C[4], double scale, int dims[6], double T34[6], int *xlist,
                       int *ylist, int cap) {
  if (!cfg || !C || !dims || !T34 || cap < 0 || (cap > 0 && (!xlist || !ylist))) return PS_ERR_INVALID;
  if (cfg->num_rotation_steps < 1 || cfg->height < 1 || cfg->width < 1 || !(scale > 0)) return PS_ERR_INVALID;
  psg::Grid g;
  g.R = cfg->num_rotation_steps; g.H = cfg->height; g.W = cfg->width;
  g.min_rot = cfg->min_part_rotation; g.max_rot = cfg->max_part_rotation;
  const double zero[2] = {0.0, 0.0};
  const psg::MessagePlan p = psg::plan_message(g, zero, zero, C, 0.0, 0.0, scale);
  if (!p.error.empty() || p.diag) return PS_ERR_INVALID;
  dims[0] = p.EH; dims[1] = p.EW;
  dims[2] = ((int)p.fx.size() - 1) / 2; dims[3] = ((int)p.fy.size() - 1) / 2;
  dims[4] = (int)p.xtiles.size() / 2; dims[5] = (int)p.ytiles.size() / 2;
  for (int k = 0; k < 6; ++k) T34[k] = p.T34[k];
  const auto put = [cap](const std::vector<int> &src, int *dst) {
    for (size_t i = 0; i + 1 < src.size() && (int)(i / 2) < cap; i += 2) {
      dst[3 * (i / 2) + 0] = src[i];
      dst[3 * (i / 2) + 1] = src[i + 1] & 0xfff;
      dst[3 * (i / 2) + 2] = src[i + 1] >> 12;
    }
  };
  put(p.xtiles, xlist);
  put(p.ytiles, ylist);
  return PS_OK;
}

int ps_plan_walks(const ps_config *cfg, const double C[4], double scale, int dims[7], int *walks, int cap,
                  unsigned char *masks, int mask_cap, int *nmasks) {
  if (!cfg || !C || !dims || !nmasks || cap < 0 || mask_cap < 0 || (cap > 0 && !walks) || (mask_cap > 0 && !masks))
    return PS_ERR_INVALID;
  if (cfg->num_rotation_steps < 1 || cfg->height < 1 || cfg->width < 1 || !(scale > 0)) return PS_ERR_INVALID;
  psg::Grid g;
  g.R = cfg->num_rotation_steps; g.H = cfg->height; g.W = cfg->width;
  g.min_rot = cfg->min_part_rotation; g.max_rot = cfg->max_part_rotation;
  const double zero[2] = {0.0, 0.0};
  const psg::MessagePlan p = psg::plan_message(g, zero, zero, C, 0.0, 0.0, scale);
  if (!p.error.empty() || p.diag) return PS_ERR_INVALID;
  dims[0] = p.EH; dims[1] = p.EW;
  dims[2] = ((int)p.fx.size() - 1) / 2; dims[3] = ((int)p.fy.size() - 1) / 2;
  dims[4] = p.halo; dims[5] = p.lag; dims[6] = (int)p.walks.size() / 4;
  for (size_t i = 0; i < p.walks.size() && (int)(i / 4) < cap; ++i) walks[i] = p.walks[i];
  for (size_t i = 0; i < p.xmasks.size() && (int)i < mask_cap; ++i) masks[i] = p.xmasks[i];
  *nmasks = (int)p.xmasks.size();
  return PS_OK;
}

int ps_index_from_rot(const ps_config *cfg, double rot) {
  return psg::index_from_value(cfg->min_part_rotation, cfg->max_part_rotation, cfg->num_rotation_steps, rot);
}

void ps_flip_joint(ps_joint *j) {
  // C <- T*(C*T), T = diag(-1, 1); offsets <- T*offset; rot_mean <- -rot_mean (aux.cpp:102-119)
  const double T[2][2] = {{-1, 0}, {0, 1}};
  double C[2][2] = {{j->C[0], j->C[1]}, {j->C[2], j->C[3]}}, CT[2][2], R[2][2];
  for (int i = 0; i < 2; ++i)
    for (int k = 0; k < 2; ++k) {
      double t = 0;
      for (int l = 0; l < 2; ++l) t += C[i][l] * T[l][k];
      CT[i][k] = t;
    }
  for (int i = 0; i < 2; ++i)
    for (int k = 0; k < 2; ++k) {
      double t = 0;
      for (int l = 0; l < 2; ++l) t += T[i][l] * CT[l][k];
      R[i][k] = t;
    }
  double op[2], oc[2];
  for (int i = 0; i < 2; ++i) {
    double t = 0, u = 0;
    for (int l = 0; l < 2; ++l) {
      t += T[i][l] * j->offset_p[l];
      u += T[i][l] * j->offset_c[l];
    }
    op[i] = t;
    oc[i] = u;
  }
  j->C[0] = R[0][0]; j->C[1] = R[0][1]; j->C[2] = R[1][0]; j->C[3] = R[1][1];
  j->offset_p[0] = op[0]; j->offset_p[1] = op[1];
  j->offset_c[0] = oc[0]; j->offset_c[1] = oc[1];
  if (j->type == PS_JOINT_ROT_GAUSSIAN) j->rot_mean = -j->rot_mean;
}

// ---- conditioning tables: host arithmetic of objectdetect_icps.cpp (float/double mix kept) ---------
static inline float sqf(float t) { return t * t; }  // pow(float,int) / square<float> of the C++98 reference build

void ps_rot_score_table(const ps_config *cfg, double mu_d, double var_d, float *table) {
  float mu = (float)mu_d, var = (float)var_d;  // icps.cpp:245-246
  for (int r = 0; r < cfg->num_rotation_steps; ++r) {
    float rot = (float)(ps_rot_from_index(cfg, r) / 180 * M_PI);
    float score = (float)std::exp(-0.5 * sqf(rot - mu) / var);
    if (score < 1e-4) score = (float)1e-4;
    score = logf(score);
    table[r] = score;
  }
}

void ps_pos_score_table(int H, int W, double mu_x_d, double mu_y_d, double var_x_d, double var_y_d, double root_x,
                        double root_y, float *table) {
  float var_weight = 1.0f;
  float mu_x = (float)mu_x_d, mu_y = (float)mu_y_d;
  float var_x = (float)(var_x_d * var_weight), var_y = (float)(var_y_d * var_weight);
  for (int iy = 0; iy < H; ++iy)
    for (int ix = 0; ix < W; ++ix) {
      float ix_rel = (float)(ix - root_x), iy_rel = (float)(iy - root_y);
      float sx = (float)std::exp(-0.5 * sqf(ix_rel - mu_x) / var_x);
      float sy = (float)std::exp(-0.5 * sqf(iy_rel - mu_y) / var_y);
      float score = sx * sy;
      if (score < 1e-4) score = (float)1e-4;
      table[(size_t)iy * W + ix] = logf(score);
    }
}

void ps_torso_prior_table(int H, int W, double mu_x_d, double mu_y_d, double var_x_d, double var_y_d, float weight,
                          float *table) {
  float mu_x = (float)mu_x_d, mu_y = (float)mu_y_d, var_x = (float)var_x_d, var_y = (float)var_y_d;
  float img_c_x = (float)(0.5 * W), img_c_y = (float)(0.5 * H);
  float var_weight = 1.0f * 1.0f;
  var_x *= var_weight;
  var_y *= var_weight;
  for (int iy = 0; iy < H; ++iy)
    for (int ix = 0; ix < W; ++ix) {
      float ix_rel = img_c_x - ix, iy_rel = img_c_y - iy;
      float sx = (float)std::exp(-0.5 * sqf(ix_rel - mu_x) / var_x);
      float sy = (float)std::exp(-0.5 * sqf(iy_rel - mu_y) / var_y);
      float score = weight * sx * sy;
      if (score < 1e-4) score = (float)1e-4;
      table[(size_t)iy * W + ix] = logf(score);
    }
}

// ---- lifetime -------------------------------------------------------------------------------------

int ps_create(const ps_config *cfg, ps_ctx **out) {
  if (!cfg || !out) {
    g_create_error = "ps_create: null argument";
    return PS_ERR_INVALID;
  }
  *out = nullptr;
  std::string why;
  if (validate_config(cfg, why)) {
    g_create_error = "ps_create: " + why;
    return PS_ERR_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("ps_create: no CUDA device (") + cudaGetErrorString(e) +
                     "); libpsinfer has no CPU fallback";
    return PS_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) {
    g_create_error = "ps_create: device ordinal out of range";
    return PS_ERR_INVALID;
  }
  e = cudaSetDevice(cfg->device);
  if (e != cudaSuccess) {
    g_create_error = std::string("ps_create: cudaSetDevice: ") + cudaGetErrorString(e);
    return PS_ERR_CUDA;
  }
  std::unique_ptr<ps_ctx> c(new ps_ctx);
  c->cfg = *cfg;
  c->R = cfg->num_rotation_steps; c->S = cfg->num_scale_steps;
  c->H = cfg->height; c->W = cfg->width; c->P = cfg->num_parts;
  c->HW = (size_t)c->H * c->W;
  c->N = c->HW * c->R;
  c->root = cfg->root_idx;
  if (c->root < 0) {
    for (int p = 0; p < c->P; ++p)
      if (cfg->is_detect[p] && cfg->is_root[p]) {
        if (c->root >= 0) {
          g_create_error = "ps_create: more than one root part (findrot.cpp:786)";
          return PS_ERR_INVALID;
        }
        c->root = p;
      }
  }
  if (c->root < 0 || c->root >= c->P) {
    g_create_error = "ps_create: root part not found (findrot.cpp:831)";
    return PS_ERR_INVALID;
  }
  c->cfg.root_idx = c->root;

  auto cu = [&](cudaError_t err, const char *what) -> bool {
    if (err == cudaSuccess) return true;
    g_create_error = std::string("ps_create: ") + what + ": " + cudaGetErrorString(err);
    return false;
  };
  if (!cu(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking), "cudaStreamCreate")) return PS_ERR_CUDA;
  c->stream = c->own_stream;
  const size_t G = c->N * sizeof(float);
  if (!cu(c->unary.alloc(G * c->P * c->S), "alloc unaries")) return PS_ERR_CUDA;
  if (!cu(c->post.alloc(G * c->P * (cfg->keep_all_scales ? c->S : 1)), "alloc beliefs")) return PS_ERR_CUDA;
  if (!cu(c->tmp[0].alloc(G), "alloc tmp") || !cu(c->tmp[1].alloc(G), "alloc tmp")) return PS_ERR_CUDA;
  if (!cu(c->bufB.alloc(G), "alloc scratch")) return PS_ERR_CUDA;
  if (!cu(c->bufU.alloc(G), "alloc scratch") || !cu(c->bufV.alloc(G), "alloc scratch")) return PS_ERR_CUDA;
  c->scratch_elems = c->N;
  if (!cu(c->root_post.alloc(c->HW * c->S * sizeof(float)), "alloc root posterior")) return PS_ERR_CUDA;
  if (!cu(c->maxes.alloc((2 * c->P + psk::kMaxRootChildren + 4) * sizeof(int)), "alloc maxima")) return PS_ERR_CUDA;
  if (!cu(c->argmax_keys.alloc(c->P * sizeof(unsigned long long)), "alloc argmax")) return PS_ERR_CUDA;
  if (!cu(c->counters.alloc(8 * sizeof(unsigned)), "alloc counters")) return PS_ERR_CUDA;
  if (!cu(c->unary_max.alloc((size_t)c->P * c->S * sizeof(int)), "alloc unary maxima")) return PS_ERR_CUDA;
  c->unary_max_valid.assign((size_t)c->P * c->S, 0);
  if (!cu(cudaMallocHost((void **)&c->host_keys, c->P * sizeof(unsigned long long)), "alloc pinned keys")) return PS_ERR_CUDA;
  cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, cfg->device);
  c->disable_tma = getenv("PSINFER_NO_TMA") != nullptr;
  c->disable_tile_lists = getenv("PSINFER_ALL_TILES") != nullptr;
  c->disable_batch = getenv("PSINFER_NO_BATCH") != nullptr;
  c->disable_graph = getenv("PSINFER_NO_GRAPH") != nullptr;
  {
    // kernel attributes are per device and never change: set them once per device, under a lock -- several host threads
    // create contexts at once (findObjectDataset's workers) while others are already launching these kernels
    static std::mutex mu;
    static std::vector<char> done(64, 0);
    std::lock_guard<std::mutex> lock(mu);
    if (cfg->device >= (int)done.size() || !done[cfg->device]) {
      if (!cu(cudaFuncSetAttribute(psk::k_gauss_xy<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemMax), "smem attr") ||
          !cu(cudaFuncSetAttribute(psk::k_gauss_xy<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemMax), "smem attr") ||
          !cu(cudaFuncSetAttribute(psk::k_gauss_xy<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemMax), "smem attr") ||
          !cu(cudaFuncSetAttribute(psk::k_gauss_xy<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemMax), "smem attr") ||
          !cu(cudaFuncSetAttribute(psk::k_conv_cols_tma2<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024), "smem attr") ||
          !cu(cudaFuncSetAttribute(psk::k_conv_cols_tma2<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 104 * 1024), "smem attr") ||
          !cu(cudaFuncSetAttribute(psk::k_conv_rows3<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024), "smem attr") ||
          !cu(cudaFuncSetAttribute(psk::k_conv_cols2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget), "smem attr") ||
          !cu(cudaFuncSetAttribute(psk::k_conv_rows2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget), "smem attr") ||
          !cu(cudaFuncSetAttribute(psk::k_conv_rows<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget), "smem attr"))
        return PS_ERR_CUDA;
      if (cfg->device < (int)done.size()) done[cfg->device] = 1;
    }
  }
  // non-detect parts have all-zero unaries in the reference (findrot.cpp:794 resize, never loaded)
  if (!cu(cudaMemsetAsync(c->unary.p, 0, c->unary.bytes, c->stream), "memset")) return PS_ERR_CUDA;

  // upright mask (findrot.cpp:512-515) and valid root rotations (:694-712)
  std::vector<unsigned char> mask(c->R);
  for (int r = 0; r < c->R; ++r) mask[r] = !(std::fabs(ps_rot_from_index(cfg, r)) < 15.0);
  if (!cu(c->upright_mask.alloc(c->R), "alloc mask")) return PS_ERR_CUDA;
  if (!cu(cudaMemcpy(c->upright_mask.p, mask.data(), c->R, cudaMemcpyHostToDevice), "copy mask")) return PS_ERR_CUDA;
  std::vector<int> valid;
  if (cfg->is_upright[c->root]) {
    int k1 = ps_index_from_rot(cfg, -1e-6), k2 = ps_index_from_rot(cfg, 1e-6);
    if (k1 < 0 || k2 < 0) {
      g_create_error = "ps_create: upright root needs 0 degrees inside the rotation range (partapp_aux.hpp:35-38)";
      return PS_ERR_INVALID;
    }
    valid.push_back(k1);
    if (k2 != k1) valid.push_back(k2);
  } else {
    for (int r = 0; r < c->R; ++r) valid.push_back(r);
  }
  c->n_valid_rots = (int)valid.size();
  if (!cu(c->valid_rots.alloc(valid.size() * sizeof(int)), "alloc rots")) return PS_ERR_CUDA;
  if (!cu(cudaMemcpy(c->valid_rots.p, valid.data(), valid.size() * sizeof(int), cudaMemcpyHostToDevice), "copy rots"))
    return PS_ERR_CUDA;
  if (!cu(cudaStreamSynchronize(c->stream), "sync")) return PS_ERR_CUDA;
  if (math_tables_init(c.get())) {
    g_create_error = "ps_create: " + c->err;
    return PS_ERR_CUDA;
  }
  *out = c.release();
  return PS_OK;
}

void ps_destroy(ps_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  cudaStreamSynchronize(ctx->stream);
  delete ctx;
}

int ps_set_stream(ps_ctx *c, void *s) {
  if (!c) return PS_ERR_INVALID;
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return PS_OK;
}

int ps_synchronize(ps_ctx *c) {
  if (!c) return PS_ERR_INVALID;
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PS_OK;
}

// ---- model ----------------------------------------------------------------------------------------

int ps_set_joints(ps_ctx *c, const ps_joint *joints, int nj) {
  if (!c || !joints) return PS_ERR_INVALID;
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  if (nj != c->P - 1) return c->fail(PS_ERR_INVALID, "need num_parts-1 = %d joints, got %d (aux.cpp:129)", c->P - 1, nj);
  std::vector<Node> nodes(c->P);
  // "currently no models with heterogeneous joints are supported" (aux.cpp:346): all ROT_GAUSSIAN -> the
  // computeRootPosteriorRot path, all POS_GAUSSIAN -> the legacy 2-D computeRootPosterior (findpos.cpp:172-334)
  const bool pos_model = nj > 0 && joints[0].type == PS_JOINT_POS_GAUSSIAN;
  for (int j = 0; j < nj; ++j) {
    const ps_joint &q = joints[j];
    if (q.type != (pos_model ? PS_JOINT_POS_GAUSSIAN : PS_JOINT_ROT_GAUSSIAN))
      return c->fail(PS_ERR_UNSUPPORTED, "joint %d: joints must be all ROT_GAUSSIAN or all POS_GAUSSIAN (aux.cpp:346, findrot.cpp:766)", j);
    if (q.child_idx < 0 || q.child_idx >= c->P || q.parent_idx < 0 || q.parent_idx >= c->P || q.child_idx == q.parent_idx)
      return c->fail(PS_ERR_INVALID, "joint %d: part index out of range (aux.cpp:126-127)", j);
    if (nodes[q.child_idx].parent >= 0) return c->fail(PS_ERR_INVALID, "part %d has two parents", q.child_idx);
    nodes[q.child_idx].parent = q.parent_idx;
    nodes[q.child_idx].joint = j;
    nodes[q.parent_idx].children.push_back(q.child_idx);
    nodes[q.parent_idx].child_joints.push_back(j);
  }
  if (nodes[c->root].parent >= 0) return c->fail(PS_ERR_INVALID, "root part %d has a parent", c->root);
  for (int p = 0; p < c->P; ++p) {
    // every part must reach the root (a tree); non-root parts may have at most one child (findrot.cpp:207-210)
    int q = p, steps = 0;
    while (q != c->root && q >= 0 && steps <= c->P) {
      q = nodes[q].parent;
      ++steps;
    }
    if (q != c->root) return c->fail(PS_ERR_INVALID, "part %d is not connected to the root", p);
    if (!pos_model && p != c->root && nodes[p].children.size() > 1)
      return c->fail(PS_ERR_INVALID, "part %d has %zu children; only the root may branch (findrot.cpp:210)", p,
                     nodes[p].children.size());
  }
  if ((int)nodes[c->root].children.size() > psk::kMaxRootChildren)
    return c->fail(PS_ERR_UNSUPPORTED, "root has more than %d children", psk::kMaxRootChildren);
  if (pos_model) {  // no per-message plans up front: ps_pos_message's plan cache fills as the messages run
    for (int j = 0; j < nj; ++j)
      if (!(joints[j].C[0] * joints[j].C[3] - joints[j].C[1] * joints[j].C[2] > 0))
        return c->fail(PS_ERR_INVALID, "joint %d: covariance must have a positive determinant (aux.cpp:133)", j);
    c->plans.clear();
    c->nodes = std::move(nodes);
    c->joints.assign(joints, joints + nj);
    c->joints_set = true;
    c->pos_model = true;
    c->have_result = false;
    c->result_pending = false;
    ++c->joints_version;
    return PS_OK;
  }
  c->pos_model = false;

  std::vector<std::shared_ptr<DevPlan>> plans((size_t)nj * 2 * c->S);
  size_t need = c->N;
  for (int j = 0; j < nj; ++j)
    for (int dir = 0; dir < 2; ++dir)
      for (int s = 0; s < c->S; ++s) {
        const ps_joint &q = joints[j];
        std::shared_ptr<DevPlan> dp;
        // upward: (offset_c, offset_p, +rot_mean) findrot.cpp:630-635; downward: (offset_p, offset_c, -rot_mean) :174-179
        int rc = dir == 0 ? get_plan(c, q.offset_c, q.offset_p, q.C, q.rot_mean, q.rot_sigma, scale_of(c->cfg, s), dp)
                          : get_plan(c, q.offset_p, q.offset_c, q.C, -q.rot_mean, q.rot_sigma, scale_of(c->cfg, s), dp);
        if (rc) {
          std::string why = c->err;
          return c->fail(rc, "joint %d: %s", j, why.c_str());
        }
        need = std::max(need, plan_scratch_elems(c, *dp));
        plans[((size_t)j * 2 + dir) * c->S + s] = dp;
      }
  int rc = ensure_scratch(c, need);
  if (rc) return rc;
  size_t nroot = nodes[c->root].children.size();
  if (c->rootmsg.bytes < nroot * c->N * sizeof(float)) {
    PS_CUDA(c, cudaStreamSynchronize(c->stream));
    PS_CUDA(c, c->rootmsg.alloc(nroot * c->N * sizeof(float)));
  }
  bool deep = false;
  for (int q : nodes[c->root].children) deep = deep || !nodes[q].children.empty();
  if (deep && c->chain_tmp.bytes < nroot * c->N * sizeof(float)) {
    PS_CUDA(c, cudaStreamSynchronize(c->stream));
    PS_CUDA(c, c->chain_tmp.alloc(nroot * c->N * sizeof(float)));
  }
  if (plans != c->plans || (int)c->joints.size() != nj ||
      memcmp(c->joints.data(), joints, sizeof(ps_joint) * (size_t)nj) != 0)
    ++c->joints_version;  // graphs captured for another joint set carry its plans' pointers
  c->plans = std::move(plans);
  c->nodes = std::move(nodes);
  c->joints.assign(joints, joints + nj);
  c->joints_set = true;
  c->have_result = false;
  c->result_pending = false;
  return PS_OK;
}

// ---- unaries --------------------------------------------------------------------------------------

int ps_set_unary(ps_ctx *c, int part, int scale, const float *src, int mem_kind, int raw) {
  if (!c || !src) return PS_ERR_INVALID;
  if (part < 0 || part >= c->P || scale < 0 || scale >= c->S) return c->fail(PS_ERR_INVALID, "part/scale out of range");
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  float *dst = c->U(part, scale);
  c->lattice_clean = false;
  c->unary_max_valid[(size_t)part * c->S + scale] = 0;
  PS_CUDA(c, cudaMemcpyAsync(dst, src, c->N * sizeof(float),
                             mem_kind == PS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
  if (raw) {
    PS_LAUNCH(c, KC_PREP, psk::k_prepare_unary<<<std::min(cdiv(c->N / 4 + 1, 256), (unsigned)c->num_sms * 16), 256, 0, c->stream>>>(dst, c->N, nullptr));
  }
  return PS_OK;
}

static int set_unary_compact_impl(ps_ctx *c, int part, int scale, const float *cells, int gh, int gw, const double *Tig,
                                  int mem_kind, int raw);
int ps_set_unary_compact(ps_ctx *c, int part, int scale, const float *cells, int gh, int gw, const double *Tig,
                         int mem_kind) {
  return set_unary_compact_impl(c, part, scale, cells, gh, gw, Tig, mem_kind, 0);
}
int ps_set_unary_compact_raw(ps_ctx *c, int part, int scale, const float *cells, int gh, int gw, const double *Tig,
                             int mem_kind) {
  return set_unary_compact_impl(c, part, scale, cells, gh, gw, Tig, mem_kind, 1);
}
int ps_log_unary(ps_ctx *c, int part, int scale) {
  if (!c) return PS_ERR_INVALID;
  if (part < 0 || part >= c->P || scale < 0 || scale >= c->S) return c->fail(PS_ERR_INVALID, "part/scale out of range");
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  int *mslot = c->unary_max.as<int>() + (size_t)part * c->S + scale;
  c->lattice_clean = false;
  PS_LAUNCH(c, KC_MISC, psk::k_set_int<<<1, 32, 0, c->stream>>>(mslot, 1, PS_ENC_NEG_INF));
  PS_LAUNCH(c, KC_PREP, psk::k_prepare_unary<<<std::min(cdiv(c->N / 4 + 1, 256), (unsigned)c->num_sms * 16), 256, 0, c->stream>>>(
                            c->U(part, scale), c->N, mslot));
  c->unary_max_valid[(size_t)part * c->S + scale] = 1;
  return PS_OK;
}
static int set_unary_compact_impl(ps_ctx *c, int part, int scale, const float *cells, int gh, int gw, const double *Tig,
                                  int mem_kind, int raw) {
  if (!c || !cells || !Tig) return PS_ERR_INVALID;
  if (part < 0 || part >= c->P || scale < 0 || scale >= c->S) return c->fail(PS_ERR_INVALID, "part/scale out of range");
  if (gh < 1 || gw < 1 || (size_t)gh * gw >= ((size_t)1 << 31)) return c->fail(PS_ERR_INVALID, "compact grid size out of range");
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  c->lattice_clean = false;
  const size_t ncell = (size_t)c->R * gh * gw;
  // staging: cells | Tig rows (doubles first for alignment)
  const size_t need = (size_t)c->R * 6 * sizeof(double) + ncell * sizeof(float);
  if (c->ingest.bytes < need) {
    PS_CUDA(c, cudaStreamSynchronize(c->stream));
    PS_CUDA(c, c->ingest.alloc(need));
  }
  if (c->ingest_keys.bytes < c->N * sizeof(int)) PS_CUDA(c, c->ingest_keys.alloc(c->N * sizeof(int)));
  // rows 0,1 of each 3x3 (map_point only reads those, homogeneous_coord.h:79-80)
  psk::TigRows rows;
  double *dT = nullptr;
  float *dcells = (float *)(c->ingest.as<double>() + (size_t)c->R * 6);
  // TM_DIRECT maps forward with Tig; TM_BILINEAR gathers through T13 = prod(inverse(Tig), identity)
  // (transform.hpp:131,196; the product with the identity changes nothing)
  std::vector<double> t((size_t)c->R * 6);
  for (int r = 0; r < c->R; ++r) {
    if (c->cfg.interpolate) {
      psg::M3 T;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T.a[i][j] = Tig[(size_t)r * 9 + i * 3 + j];
      const psg::M3 inv = psg::inverse(T);
      for (int k = 0; k < 6; ++k) t[(size_t)r * 6 + k] = inv.a[k / 3][k % 3];
    } else {
      for (int k = 0; k < 6; ++k) t[(size_t)r * 6 + k] = Tig[(size_t)r * 9 + k];
    }
  }
  if (c->R <= psk::kMaxIngestRot) {
    for (size_t k = 0; k < t.size(); ++k) rows.m[k] = t[k];
  } else {
    // the staging rows are shared by every (part, scale) call and c->stream does not synchronise with the legacy
    // stream: wait for the previous call's scatter kernel before its transforms are overwritten, copy in stream order
    dT = c->ingest.as<double>();
    PS_CUDA(c, cudaStreamSynchronize(c->stream));
    PS_CUDA(c, cudaMemcpyAsync(dT, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PS_CUDA(c, cudaStreamSynchronize(c->stream));  // `t` goes out of scope
  }
  const float *src_cells = cells;
  if (mem_kind == PS_MEM_HOST) {
    PS_CUDA(c, cudaMemcpyAsync(dcells, cells, ncell * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    src_cells = dcells;
  }
  psk::IngestArgs a;
  a.cells = src_cells; a.Tig = dT; a.keys = c->ingest_keys.as<int>(); a.out = c->U(part, scale);
  a.R = c->R; a.gh = gh; a.gw = gw; a.H = c->H; a.W = c->W;
  a.raw = raw;
  int *mslot = c->unary_max.as<int>() + (size_t)part * c->S + scale;
  // collision-free? smallest singular value of every 2x2 linear part must exceed sqrt(2) (plus a safety margin)
  bool collision_free = true;
  for (int r = 0; r < c->R && collision_free; ++r) {
    const double a00 = Tig[r * 9 + 0], a01 = Tig[r * 9 + 1], a10 = Tig[r * 9 + 3], a11 = Tig[r * 9 + 4];
    const double s1 = a00 * a00 + a01 * a01 + a10 * a10 + a11 * a11, det = a00 * a11 - a01 * a10;
    const double disc = std::sqrt(std::max(0.0, s1 * s1 - 4 * det * det));
    const double smin2 = 0.5 * (s1 - disc);  // squared smallest singular value
    collision_free = smin2 > 2.0 * 1.01;
  }
  if (c->cfg.interpolate || !collision_free)  // the direct path resets the slot inside its fill
    PS_LAUNCH(c, KC_MISC, psk::k_set_int<<<1, 32, 0, c->stream>>>(mslot, 1, PS_ENC_NEG_INF));
  if (c->cfg.interpolate) {
    PS_LAUNCH(c, KC_PREP, psk::k_ingest_bilinear<<<dim3(cdiv(c->HW, 256), c->R), 256, 0, c->stream>>>(
                              a, rows, psk::FastDiv((unsigned)c->W), mslot));
  } else if (collision_free) {
    PS_LAUNCH(c, KC_PREP, psk::k_fill<<<std::min(cdiv(c->N, 4096), (unsigned)c->num_sms * 8), 256, 0, c->stream>>>(
                              a.out, c->N, raw ? 0.0f : psk::kLogZero, mslot, PS_ENC_NEG_INF));
    PS_LAUNCH(c, KC_PREP, psk::k_ingest_scatter_direct<<<dim3(cdiv((size_t)gh * gw, 256), c->R), 256, 0, c->stream>>>(a, rows, mslot));
  } else {
    PS_CUDA(c, cudaMemsetAsync(c->ingest_keys.p, 0, c->N * sizeof(int), c->stream));
    PS_LAUNCH(c, KC_PREP, psk::k_ingest_scatter<<<dim3(cdiv((size_t)gh * gw, 256), c->R), 256, 0, c->stream>>>(a, rows));
    PS_LAUNCH(c, KC_PREP, psk::k_ingest_sweep<<<cdiv((c->N + 3) / 4, 256), 256, 0, c->stream>>>(a, mslot));
  }
  c->unary_max_valid[(size_t)part * c->S + scale] = raw ? 0 : 1;  // a raw grid is not what the messages read
  return PS_OK;
}

int ps_set_unaries_compact(ps_ctx *c, int n, const int *parts, const int *scales, const float *const *cells, int gh, int gw,
                           const double *Tig, int mem_kind) {
  if (!c || !parts || !scales || !cells || !Tig || n < 1) return PS_ERR_INVALID;
  if (gh < 1 || gw < 1 || (size_t)gh * gw >= ((size_t)1 << 31)) return c->fail(PS_ERR_INVALID, "compact grid size out of range");
  for (int i = 0; i < n; ++i)
    if (parts[i] < 0 || parts[i] >= c->P || scales[i] < 0 || scales[i] >= c->S || !cells[i])
      return c->fail(PS_ERR_INVALID, "part/scale out of range");
  // the shared-lattice fast path needs the collision-free direct scatter with by-value transforms
  bool fast = !c->cfg.interpolate && c->R <= psk::kMaxIngestRot;
  for (int r = 0; r < c->R && fast; ++r) {
    const double a00 = Tig[r * 9 + 0], a01 = Tig[r * 9 + 1], a10 = Tig[r * 9 + 3], a11 = Tig[r * 9 + 4];
    const double s1 = a00 * a00 + a01 * a01 + a10 * a10 + a11 * a11, det = a00 * a11 - a01 * a10;
    const double disc = std::sqrt(std::max(0.0, s1 * s1 - 4 * det * det));
    fast = 0.5 * (s1 - disc) > 2.0 * 1.01;  // squared smallest singular value, as in set_unary_compact_impl
  }
  if (!fast) {
    for (int i = 0; i < n; ++i)
      if (int rc = ps_set_unary_compact(c, parts[i], scales[i], cells[i], gh, gw, Tig, mem_kind)) return rc;
    return PS_OK;
  }
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  const size_t ncell = (size_t)c->R * gh * gw;
  psk::TigRows rows;
  for (int r = 0; r < c->R; ++r)
    for (int k = 0; k < 6; ++k) rows.m[(size_t)r * 6 + k] = Tig[(size_t)r * 9 + k];
  if (mem_kind == PS_MEM_HOST && c->ingest_batch.bytes < (size_t)n * ncell * sizeof(float)) {
    PS_CUDA(c, cudaStreamSynchronize(c->stream));
    PS_CUDA(c, c->ingest_batch.alloc((size_t)n * ncell * sizeof(float)));
  }
  // one fill for everything when the call covers the whole unary buffer, else one per grid
  bool whole = n == c->P * c->S;
  if (whole) {
    std::vector<unsigned char> seen((size_t)c->P * c->S, 0);
    for (int i = 0; i < n; ++i) seen[(size_t)parts[i] * c->S + scales[i]] = 1;
    for (unsigned char v : seen) whole = whole && v;
  }
  // Same lattice as the image before and nothing else has written the unaries since: every lattice cell is overwritten
  // (LOG_ZERO where the score is 0), every other cell still holds the fill -- the 4 P S R H W-byte fill is skipped.
  std::vector<double> sig((size_t)c->R * 6 + 2);
  for (size_t k = 0; k < (size_t)c->R * 6; ++k) sig[k] = rows.m[k];
  sig[(size_t)c->R * 6] = gh;
  sig[(size_t)c->R * 6 + 1] = gw;
  static const bool no_skip = getenv("PSINFER_ALWAYS_FILL") != nullptr;  // A/B switch
  const bool overwrite = whole && c->lattice_clean && !no_skip && sig.size() == c->lattice_sig.size() &&
                         memcmp(sig.data(), c->lattice_sig.data(), sig.size() * sizeof(double)) == 0;
  c->lattice_clean = false;  // until every launch below has been issued
  if (whole) {
    if (!overwrite)
      PS_LAUNCH(c, KC_PREP, psk::k_fill<<<std::min(cdiv((size_t)n * c->N, 4096), (unsigned)c->num_sms * 16), 256, 0, c->stream>>>(
                                c->unary.as<float>(), (size_t)n * c->N, psk::kLogZero));
    PS_LAUNCH(c, KC_MISC, psk::k_set_int<<<cdiv(n, 128), 128, 0, c->stream>>>(c->unary_max.as<int>(), n, PS_ENC_NEG_INF));
  }
  psk::IngestArgs a;
  a.cells = nullptr; a.Tig = nullptr; a.keys = nullptr; a.out = nullptr;
  a.R = c->R; a.gh = gh; a.gw = gw; a.H = c->H; a.W = c->W;
  a.raw = 0;
  for (int i0 = 0; i0 < n; i0 += psk::kMaxIngestBatch) {
    const int nb = std::min(psk::kMaxIngestBatch, n - i0);
    psk::IngestBatch b{};
    for (int k = 0; k < nb; ++k) {
      const int i = i0 + k;
      int *mslot = c->unary_max.as<int>() + (size_t)parts[i] * c->S + scales[i];
      if (!whole)
        PS_LAUNCH(c, KC_PREP, psk::k_fill<<<std::min(cdiv(c->N, 4096), (unsigned)c->num_sms * 8), 256, 0, c->stream>>>(
                                  c->U(parts[i], scales[i]), c->N, psk::kLogZero, mslot, PS_ENC_NEG_INF));
      const float *src = cells[i];
      if (mem_kind == PS_MEM_HOST) {
        float *d = c->ingest_batch.as<float>() + (size_t)i * ncell;
        PS_CUDA(c, cudaMemcpyAsync(d, cells[i], ncell * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        src = d;
      }
      b.cells[k] = src;
      b.out[k] = c->U(parts[i], scales[i]);
      b.max_dst[k] = mslot;
      c->unary_max_valid[(size_t)parts[i] * c->S + scales[i]] = 1;
    }
    if (overwrite)
      PS_LAUNCH(c, KC_PREP, psk::k_ingest_scatter_direct_b<true><<<dim3(cdiv((size_t)gh * gw, 256), c->R, nb), 256, 0, c->stream>>>(a, b, rows));
    else
      PS_LAUNCH(c, KC_PREP, psk::k_ingest_scatter_direct_b<false><<<dim3(cdiv((size_t)gh * gw, 256), c->R, nb), 256, 0, c->stream>>>(a, b, rows));
  }
  if (whole) {
    c->lattice_sig.swap(sig);
    c->lattice_clean = true;
  }
  return PS_OK;
}

int ps_get_unary(ps_ctx *c, int part, int scale, float *dst, int mem_kind) {
  if (!c || !dst) return PS_ERR_INVALID;
  if (part < 0 || part >= c->P || scale < 0 || scale >= c->S) return c->fail(PS_ERR_INVALID, "part/scale out of range");
  PS_CUDA(c, cudaMemcpyAsync(dst, c->U(part, scale), c->N * sizeof(float),
                             mem_kind == PS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PS_OK;
}

// Stream-ordered staging for host tables / grids: the copy, the kernel that reads it and the next call's copy are all
// on ctx->stream, and cudaMemcpyAsync from pageable memory returns once the source has been staged -- no host
// synchronisation is needed (growing the buffer is the exception).
static int stage_host(ps_ctx *c, DevBuf &buf, size_t &used, const float *src, size_t n, const float **dev) {
  if (buf.bytes < (used + n) * sizeof(float)) return c->fail(PS_ERR_INVALID, "staging buffer too small");
  float *d = buf.as<float>() + used;
  PS_CUDA(c, cudaMemcpyAsync(d, src, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  used += n;
  *dev = d;
  return PS_OK;
}

int ps_add_unary_tables(ps_ctx *c, int part, int n, const float *const *tables, const int *kinds, const float *weights,
                        int mem_kind) {
  if (!c || !tables || !kinds || !weights) return PS_ERR_INVALID;
  if (part < 0 || part >= c->P) return c->fail(PS_ERR_INVALID, "part out of range");
  if (n < 1 || n > psk::kMaxTables) return c->fail(PS_ERR_INVALID, "1 to %d tables per call", psk::kMaxTables);
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  size_t need = 0;
  for (int k = 0; k < n; ++k) {
    if (!tables[k]) return PS_ERR_INVALID;
    if (kinds[k] < 0 || kinds[k] > 2) return c->fail(PS_ERR_INVALID, "table_kind must be 0, 1 or 2");
    need += kinds[k] == 0 ? (size_t)c->R : c->HW;
  }
  for (int s2 = 0; s2 < c->S; ++s2) c->unary_max_valid[(size_t)part * c->S + s2] = 0;
  c->lattice_clean = false;
  psk::TableArgs a{};
  a.n = n;
  if (mem_kind == PS_MEM_HOST && c->table_stage.bytes < need * sizeof(float)) {
    PS_CUDA(c, cudaStreamSynchronize(c->stream));
    PS_CUDA(c, c->table_stage.alloc(need * sizeof(float)));
  }
  size_t used = 0;
  for (int k = 0; k < n; ++k) {
    a.kind[k] = kinds[k];
    a.weight[k] = weights[k];
    a.table[k] = tables[k];
    if (mem_kind == PS_MEM_HOST)
      if (int rc = stage_host(c, c->table_stage, used, tables[k], kinds[k] == 0 ? (size_t)c->R : c->HW, &a.table[k])) return rc;
  }
  for (int s = 0; s < c->S; ++s)
    PS_LAUNCH(c, KC_MISC, psk::k_add_tables<<<dim3(std::min(cdiv(c->HW, 256), 1024u), c->R), 256, 0, c->stream>>>(c->U(part, s), c->R, c->HW, a));
  return PS_OK;
}

int ps_add_unary_table(ps_ctx *c, int part, const float *table, int kind, float weight) {
  if (!c || !table) return PS_ERR_INVALID;
  return ps_add_unary_tables(c, part, 1, &table, &kind, &weight, PS_MEM_HOST);
}

int ps_add_unary_grid(ps_ctx *c, int part, const float *grid, int num_rot, int mode, float weight, int mem_kind) {
  if (!c || !grid) return PS_ERR_INVALID;
  if (part < 0 || part >= c->P) return c->fail(PS_ERR_INVALID, "part out of range");
  if (num_rot != 1 && num_rot != c->R) return c->fail(PS_ERR_INVALID, "num_rot must be 1 or num_rotation_steps");
  if (mode < 0 || mode > 1) return c->fail(PS_ERR_INVALID, "mode must be 0 (log-domain add) or 1 (raw DPM scores)");
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  for (int s2 = 0; s2 < c->S; ++s2) c->unary_max_valid[(size_t)part * c->S + s2] = 0;
  const float *dg = grid;
  if (mem_kind == PS_MEM_HOST) {
    const size_t n = (size_t)num_rot * c->HW;
    if (c->grid_stage.bytes < n * sizeof(float)) {
      PS_CUDA(c, cudaStreamSynchronize(c->stream));
      PS_CUDA(c, c->grid_stage.alloc(n * sizeof(float)));
    }
    size_t used = 0;
    if (int rc = stage_host(c, c->grid_stage, used, grid, n, &dg)) return rc;
  }
  for (int s = 0; s < c->S; ++s)
    PS_LAUNCH(c, KC_MISC, psk::k_add_grid<<<dim3(std::min(cdiv(c->HW, 256), 1024u), c->R), 256, 0, c->stream>>>(
                              c->U(part, s), c->R, c->HW, dg, num_rot, mode, weight));
  return PS_OK;
}

// ---- readout helpers ------------------------------------------------------------------------------

namespace {

void fill_hyp(const ps_config &cfg, float *row, int scaleidx, int rotidx, int x, int y, float score) {
  // PartHyp ctor (objectdetect.h:91-97): m_scale, m_rot are floats; toVect :139-160
  row[0] = (float)scaleidx;
  row[1] = (float)ps_scale_from_index(&cfg, scaleidx);
  row[2] = (float)rotidx;
  row[3] = (float)ps_rot_from_index(&cfg, rotidx);
  row[4] = (float)x;
  row[5] = (float)y;
  row[6] = score;
}

int ensure_topk(ps_ctx *c, size_t slots, size_t kmax, size_t ncand) {
  if (c->cand.bytes < ncand * sizeof(psk::Cand)) {
    PS_CUDA(c, cudaStreamSynchronize(c->stream));
    PS_CUDA(c, c->cand.alloc(ncand * sizeof(psk::Cand)));
  }
  if (!c->topk_hist.p) PS_CUDA(c, c->topk_hist.alloc(65536 * sizeof(unsigned)));
  if (slots > c->topk_slots || kmax > c->topk_kmax) {
    PS_CUDA(c, cudaStreamSynchronize(c->stream));
    slots = std::max(slots, c->topk_slots);
    kmax = std::max(kmax, c->topk_kmax);
    PS_CUDA(c, c->topk_state.alloc(slots * sizeof(psk::TopKState)));
    PS_CUDA(c, c->topk_out.alloc(slots * std::max<size_t>(kmax, 1) * sizeof(psk::Cand)));
    PS_CUDA(c, cudaMemset(c->topk_out.p, 0, c->topk_out.bytes));  // unused winner slots are copied back too
    PS_CUDA(c, cudaMemset(c->topk_state.p, 0, c->topk_state.bytes));
    if (c->host_topk) cudaFreeHost(c->host_topk);
    if (c->host_topk_state) cudaFreeHost(c->host_topk_state);
    PS_CUDA(c, cudaMallocHost((void **)&c->host_topk, slots * std::max<size_t>(kmax, 1) * sizeof(psk::Cand)));
    PS_CUDA(c, cudaMallocHost((void **)&c->host_topk_state, slots * sizeof(psk::TopKState)));
    c->topk_slots = slots;
    c->topk_kmax = kmax;
  }
  return PS_OK;
}

// findLocalMax (aux.cpp:193-261) of a device grid [D0][H][W], asynchronously: candidates -> radix select of the
// max_n best -> winners in topk_out[slot].  Call decode_local_max(slot) after the stream has been synchronised.
int enqueue_local_max(ps_ctx *c, const float *g, int D0, int H, int W, int max_n, int slot) {
  const size_t n = (size_t)D0 * H * W;
  unsigned *cnt = c->counters.as<unsigned>();
  psk::TopKState *st = c->topk_state.as<psk::TopKState>() + slot;
  unsigned *hist = c->topk_hist.as<unsigned>();
  psk::Cand *cand = c->cand.as<psk::Cand>();
  psk::Cand *out = c->topk_out.as<psk::Cand>() + (size_t)slot * std::max<size_t>(c->topk_kmax, 1);
  PS_CUDA(c, cudaMemsetAsync(cnt, 0, sizeof(unsigned), c->stream));
  PS_LAUNCH(c, KC_LOCAL_MAX, psk::k_local_max<<<dim3(cdiv(W, 256), H, D0), 256, 0, c->stream>>>(g, D0, H, W, cand, (unsigned)n, cnt));
  PS_LAUNCH(c, KC_LOCAL_MAX, psk::k_topk_init<<<1, 1024, 0, c->stream>>>(st, cnt, (unsigned)n, (unsigned)max_n, hist));
  const unsigned blocks = std::min(cdiv(n, 256 * 8), (unsigned)c->num_sms * 8);
  for (int shift = 48; shift >= 0; shift -= 16) {
    PS_LAUNCH(c, KC_LOCAL_MAX, psk::k_topk_hist<<<blocks, 256, 0, c->stream>>>(cand, st, shift, hist));
    PS_LAUNCH(c, KC_LOCAL_MAX, psk::k_topk_scan<<<1, 1024, 0, c->stream>>>(st, shift, hist));
  }
  PS_LAUNCH(c, KC_LOCAL_MAX, psk::k_topk_compact<<<blocks, 256, 0, c->stream>>>(cand, st, out, (unsigned)std::max(max_n, 0)));
  return PS_OK;
}

// Copies the winners of slots [0, nslots) to pinned memory (asynchronous).
int fetch_local_max(ps_ctx *c, int nslots) {
  const size_t km = std::max<size_t>(c->topk_kmax, 1);
  PS_CUDA(c, cudaMemcpyAsync(c->host_topk, c->topk_out.p, (size_t)nslots * km * sizeof(psk::Cand), cudaMemcpyDeviceToHost, c->stream));
  PS_CUDA(c, cudaMemcpyAsync(c->host_topk_state, c->topk_state.p, (size_t)nslots * sizeof(psk::TopKState), cudaMemcpyDeviceToHost, c->stream));
  return PS_OK;
}

// Host ordering of one slot's winners: the reference's scan order (dim0, x, y) when nothing was cut, descending
// score otherwise (ties by scan order; std::sort leaves them unspecified).  Rows of (d0, x, y, score).
void decode_local_max(ps_ctx *c, int slot, int H, int W, std::vector<float> &rows) {
  const size_t km = std::max<size_t>(c->topk_kmax, 1);
  const psk::TopKState &st = c->host_topk_state[slot];
  const unsigned n = std::min(st.out_count, st.k);
  std::vector<psk::Cand> cand(c->host_topk + (size_t)slot * km, c->host_topk + (size_t)slot * km + n);
  std::sort(cand.begin(), cand.end(), [](const psk::Cand &a, const psk::Cand &b) { return a.key < b.key; });
  if (st.count > st.k)
    std::stable_sort(cand.begin(), cand.end(), [](const psk::Cand &a, const psk::Cand &b) { return a.score > b.score; });
  rows.resize(cand.size() * 4);
  for (size_t i = 0; i < cand.size(); ++i) {
    unsigned key = cand[i].key;
    rows[4 * i + 0] = (float)(key / ((unsigned)H * W));
    rows[4 * i + 1] = (float)((key / H) % W);
    rows[4 * i + 2] = (float)(key % H);
    rows[4 * i + 3] = cand[i].score;
  }
}

// Synchronous convenience used by ps_find_local_max and ps_max_states' slow path.
int finish_result(ps_ctx *c);
int local_max_device(ps_ctx *c, const float *g, int D0, int H, int W, int max_n, std::vector<float> &rows) {
  // slot 0 and the pinned winner buffers are shared with a pending ps_infer / ps_max_states readout: decode that first
  if (c->result_pending)
    if (int frc = finish_result(c)) return frc;
  int rc = ensure_topk(c, 1, (size_t)std::max(max_n, 1), (size_t)D0 * H * W);
  if (rc) return rc;
  if ((rc = enqueue_local_max(c, g, D0, H, W, max_n, 0))) return rc;
  if ((rc = fetch_local_max(c, 1))) return rc;
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  decode_local_max(c, 0, H, W, rows);
  return PS_OK;
}

// argmax of grids g[p] for all parts (findrot.cpp:261-277 / :93-98): device part, asynchronous
int enqueue_readout(ps_ctx *c, const std::vector<const float *> &grids, int scaleidx, int flags, bool keys_ready = false) {
  const int P = c->P;
  if (!keys_ready) {
    PS_CUDA(c, cudaMemsetAsync(c->argmax_keys.p, 0, P * sizeof(unsigned long long), c->stream));
    for (int p = 0; p < P; ++p)
      PS_LAUNCH(c, KC_ARGMAX,
                psk::k_argmax<<<std::min(cdiv(c->N / 4 + 1, 256), (unsigned)c->num_sms * 16), 256, 0, c->stream>>>(
                    grids[p], c->N, c->argmax_keys.as<unsigned long long>() + p));
  }
  PS_CUDA(c, cudaMemcpyAsync(c->host_keys, c->argmax_keys.p, P * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                             c->stream));
  if (flags & (PS_INFER_LOCAL_MAX | PS_INFER_ROOT_HYPS)) {
    // local maxima of every part (slot p) and of the root posterior (slot P): selected on the device, one copy back
    const size_t biggest = std::max(c->N, (size_t)c->S * c->HW);
    int rc = ensure_topk(c, (size_t)P + 1, (size_t)std::max(c->cfg.roi_save_num_samples, 1000), biggest);
    if (rc) return rc;
    if (flags & PS_INFER_LOCAL_MAX)
      for (int p = 0; p < P; ++p)
        if ((rc = enqueue_local_max(c, grids[p], c->R, c->H, c->W, c->cfg.roi_save_num_samples, p))) return rc;
    if (flags & PS_INFER_ROOT_HYPS)
      if ((rc = enqueue_local_max(c, c->root_post.as<float>(), c->S, c->H, c->W, 1000, P))) return rc;
    if ((rc = fetch_local_max(c, P + 1))) return rc;
  }
  c->pending_grids = grids;
  c->pending_scaleidx = scaleidx;
  c->pending_flags = flags;
  c->result_pending = true;
  c->have_result = false;
  return PS_OK;
}

// host part of the readout: decode argmax keys, then (optionally) local maxima (findrot.cpp:277-283, :1037-1038)
int finish_result(ps_ctx *c) {
  if (!c->result_pending) return c->have_result ? PS_OK : c->fail(PS_ERR_STATE, "no result: call ps_infer first");
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  c->result_pending = false;
  const int P = c->P;
  const int scaleidx = c->pending_scaleidx;
  c->best_conf.assign((size_t)P * PS_HYP_VEC, 0.f);
  c->part_hyps.assign(P, std::vector<float>());
  if (c->pending_root_only) {  // POS_GAUSSIAN: computeRootPosterior has no downward pass, hence no part estimates
    c->have_local_max = false;
    c->have_root_hyps = false;
    if (c->pending_flags & PS_INFER_ROOT_HYPS) {
      decode_local_max(c, P, c->H, c->W, c->root_hyps);
      c->have_root_hyps = true;
    }
    c->have_result = true;
    return PS_OK;
  }
  for (int p = 0; p < P; ++p) {
    unsigned long long key = c->host_keys[p];
    if (key == 0) return c->fail(PS_ERR_INVALID, "part %d: no finite maximum (findrot.cpp:273 assert)", p);
    unsigned idx = ~(unsigned)(key & 0xffffffffu);
    float val = psk::dec_f((int)((unsigned)(key >> 32) ^ 0x80000000u));
    int rot = idx / (unsigned)c->HW, rem = idx % (unsigned)c->HW;
    int y = rem / c->W, x = rem % c->W;
    fill_hyp(c->cfg, &c->best_conf[(size_t)p * PS_HYP_VEC], scaleidx, rot, x, y, val);
    c->part_hyps[p].assign(c->best_conf.begin() + (size_t)p * PS_HYP_VEC,
                           c->best_conf.begin() + (size_t)(p + 1) * PS_HYP_VEC);
  }
  const bool local_max = c->pending_flags & PS_INFER_LOCAL_MAX;
  if (local_max)
    for (int p = 0; p < P; ++p) {
      std::vector<float> rows;
      decode_local_max(c, p, c->H, c->W, rows);
      for (size_t i = 0; i < rows.size() / 4; ++i) {
        float h[PS_HYP_VEC];
        // findLocalMax wrapper tags scaleidx 0 (aux.cpp:305)
        fill_hyp(c->cfg, h, 0, (int)rows[4 * i], (int)rows[4 * i + 1], (int)rows[4 * i + 2], rows[4 * i + 3]);
        c->part_hyps[p].insert(c->part_hyps[p].end(), h, h + PS_HYP_VEC);
      }
    }
  c->have_local_max = local_max;
  c->have_root_hyps = false;
  if (c->pending_flags & PS_INFER_ROOT_HYPS) {
    decode_local_max(c, P, c->H, c->W, c->root_hyps);
    c->have_root_hyps = true;
  }
  c->have_result = true;
  return PS_OK;
}

}  // namespace

// ---- inference ------------------------------------------------------------------------------------

// Enqueues one whole inference on ctx->stream (no host synchronisation once buffers, plans and scatter maps exist).
static int infer_enqueue(ps_ctx *c, int flags) {
  const int P = c->P, S = c->S, R = c->R, root = c->root;
  const size_t N = c->N;
  const bool sparse = flags & PS_INFER_SPARSE;
  cudaStream_t st = c->stream;
  c->have_result = false;
  c->work_counter_next = kWorkCounters;  // the first fused launch resets the counter pool: every inference (and every replay of its graph) starts from zeroed counters
  int rc;

  if (flags & PS_INFER_KEEP_UNARIES) {
    if (c->unary_backup.bytes < c->unary.bytes) PS_CUDA(c, c->unary_backup.alloc(c->unary.bytes));
    PS_CUDA(c, cudaMemcpyAsync(c->unary_backup.p, c->unary.p, c->unary.bytes, cudaMemcpyDeviceToDevice, st));
  }
  // upright masks and the border strip rewrite unary cells in place: the buffer is "fill + one lattice" again only if
  // it is restored at the end (see lattice_clean)
  const bool lattice_was_clean = c->lattice_clean;
  {
    bool masks = c->cfg.strip_border_detections > 0 && !(flags & PS_INFER_NO_BORDER_STRIP);
    for (int p = 0; p < P; ++p) masks = masks || c->cfg.is_upright[p];
    if (masks) c->lattice_clean = false;
  }

  const Node &rn = c->nodes[root];
  const int nrc = (int)rn.children.size();

  for (int s = 0; s < S; ++s) {
    // upright masking of this scale (findrot.cpp:509-523)
    for (int p = 0; p < P; ++p)
      if (c->cfg.is_upright[p]) {
        PS_LAUNCH(c, KC_MASK, psk::k_mask_slices<<<dim3(std::min(cdiv(c->HW, 256), 512u), R), 256, 0, st>>>(c->U(p, s), R, c->HW,
                                                                                 c->upright_mask.as<unsigned char>()));
      }
    // border strip of the root, all scales, every iteration (findrot.cpp:528-551; idempotent)
    if (c->cfg.strip_border_detections > 0 && !(flags & PS_INFER_NO_BORDER_STRIP)) {
      int sw = (int)(c->cfg.strip_border_detections * c->W);
      if (sw > 0)
        for (int s2 = 0; s2 < S; ++s2) {
          PS_LAUNCH(c, KC_MASK, psk::k_strip_border<<<cdiv((size_t)R * c->H, 8), dim3(32, 8), 0, st>>>(c->U(root, s2), R * c->H, c->W, sw));
        }
    }

    // every max slot is written at most once per scale: one reset for all of them
    PS_LAUNCH(c, KC_MISC, psk::k_set_int<<<cdiv(2 * P + psk::kMaxRootChildren + 4, 128), 128, 0, st>>>(
                              c->maxes.as<int>(), 2 * P + psk::kMaxRootChildren + 4, PS_ENC_NEG_INF));
    const bool last_scale = s == S - 1;
    if (last_scale) PS_CUDA(c, cudaMemsetAsync(c->argmax_keys.p, 0, P * sizeof(unsigned long long), st));
    unsigned long long *keys = last_scale ? c->argmax_keys.as<unsigned long long>() : nullptr;

    // ---------------- upward pass (findrot.cpp:582-658) ----------------
    // Every root child heads a chain; chains are independent, so level t carries the t-th message (counted from the
    // leaf) of every chain that long, all in the same launches (run_level).
    // `belief[p]` is the grid MSG_up reads for part p, `bmax[p]` its encoded maximum.
    std::vector<const float *> belief(P, nullptr);
    std::vector<std::vector<int>> chains(nrc);
    size_t maxlen = 0;
    for (int ci = 0; ci < nrc; ++ci) {
      // collect the chain root-child -> ... -> leaf
      std::vector<int> &chain = chains[ci];
      for (int q = rn.children[ci]; q >= 0; q = c->nodes[q].children.empty() ? -1 : c->nodes[q].children[0])
        chain.push_back(q);
      maxlen = std::max(maxlen, chain.size());
      const int leaf = chain.back();
      // leaf belief: 0 + unary (is_detect) or all zeros
      if (c->cfg.is_detect[leaf]) {
        belief[leaf] = c->U(leaf, s);
      } else {
        PS_LAUNCH(c, KC_MISC, psk::k_fill<<<std::min(cdiv(N, 1024), 2048u), 256, 0, st>>>(c->POST(leaf, s), N, 0.0f));
        belief[leaf] = c->POST(leaf, s);
      }
      const bool masked = c->cfg.is_upright[leaf] || (leaf == root && c->cfg.strip_border_detections > 0);
      if (c->cfg.is_detect[leaf] && !masked && c->unary_max_valid[(size_t)leaf * S + s]) {
        // the ingest already folded max(unary): copy the scalar instead of sweeping 23 MB
        PS_CUDA(c, cudaMemcpyAsync(c->MAXP(leaf), c->unary_max.as<int>() + (size_t)leaf * S + s, sizeof(int),
                                   cudaMemcpyDeviceToDevice, st));
      } else if ((rc = grid_max(c, belief[leaf], N, c->MAXP(leaf)))) {
        return rc;
      }
    }
    for (size_t t = 0; t < maxlen; ++t) {
      std::vector<MsgJob> jobs;
      for (int ci = 0; ci < nrc; ++ci) {
        const std::vector<int> &chain = chains[ci];
        if (chain.size() <= t) continue;
        const int child = chain[chain.size() - 1 - t];
        const int parent = c->nodes[child].parent;
        MsgJob job;
        job.dp = c->plans[((size_t)c->nodes[child].joint * 2 + 0) * S + s].get();
        job.in = belief[child];
        job.in_max = c->MAXP(child);
        job.sparse = sparse;
        if (parent == root) {
          job.sink.out0 = c->rootmsg.as<float>() + (size_t)ci * N;  // combined later, in joint order
        } else {
          job.sink.out0 = c->POST(parent, s);
          if (c->cfg.is_detect[parent]) job.sink.add0 = c->U(parent, s);  // :652-654
          job.sink.max0 = c->MAXP(parent);
          belief[parent] = job.sink.out0;
        }
        jobs.push_back(job);
      }
      if ((rc = run_level(c, jobs))) return rc;
    }
    // root: post = sum of messages (joint order) + unary; fr_j = sum_{i != j} + unary (:637-654, :169)
    {
      psk::RootArgs a{};
      a.n = nrc;
      for (int j = 0; j < nrc; ++j) {
        a.m[j] = c->rootmsg.as<float>() + (size_t)j * N;
        a.fr_max[j] = c->MAXP(P + j);
      }
      a.amax = keys ? keys + root : nullptr;
      if (!c->cfg.is_detect[root])
        return c->fail(PS_ERR_INVALID, "root part must have is_detect set (findrot.cpp:785)");
      a.unary = c->U(root, s);
      a.post = c->POST(root, s);
      a.N = N;
      if (nrc > 0) {
        if ((rc = launch_root_combine(c, a))) return rc;
      } else {
        PS_CUDA(c, cudaMemcpyAsync(a.post, a.unary, N * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (keys)
          PS_LAUNCH(c, KC_ARGMAX, psk::k_argmax<<<std::min(cdiv(N / 4 + 1, 256), (unsigned)c->num_sms * 16), 256, 0, st>>>(a.post, N, keys + root));
      }
    }

    // ---------------- downward pass (computePartMarginals, findrot.cpp:158-236) ----------------
    // Level t carries the message into the t-th node of every chain.  A chain's inputs alternate between its root
    // message buffer (fr_j, consumed by level 0) and its chain_tmp grid.
    for (size_t t = 0; t < maxlen; ++t) {
      std::vector<MsgJob> jobs;
      for (int ci = 0; ci < nrc; ++ci) {
        const std::vector<int> &chain = chains[ci];
        if (chain.size() <= t) continue;
        const int q = chain[t];
        const Node &nq = c->nodes[q];
        float *const bufs[2] = {c->rootmsg.as<float>() + (size_t)ci * N, c->chain_tmp.as<float>() + (size_t)ci * N};
        MsgJob job;
        job.dp = c->plans[((size_t)nq.joint * 2 + 1) * S + s].get();
        job.in = bufs[t & 1];
        job.in_max = t == 0 ? c->MAXP(P + ci) : c->MAXP(P + psk::kMaxRootChildren + chain[t - 1]);
        job.sparse = false;
        // post[q] += from_root[q] (:201).  For a leaf, post[q] is still "0 + unary" held in the unary itself.
        job.sink.out0 = c->POST(q, s);
        job.sink.acc0 = belief[q];
        job.sink.amax0 = keys ? keys + q : nullptr;  // post[q] is final here: fold the readout's argmax into this write
        if (t + 1 < chain.size()) {
          // tmp = unary[q] + from_root[q] (:221-222), input of the next message
          job.sink.out1 = bufs[(t + 1) & 1];
          job.sink.add1 = c->U(q, s);
          job.sink.max1 = c->MAXP(P + psk::kMaxRootChildren + q);
        }
        jobs.push_back(job);
      }
      if ((rc = run_level(c, jobs))) return rc;
    }

    // root rotation-marginal of this scale (findrot.cpp:694-726)
    PS_LAUNCH(c, KC_ROOT_MARGINAL, psk::k_root_marginal<<<cdiv(c->HW, 256), 256, 0, st>>>(c->POST(root, s), c->HW, c->valid_rots.as<int>(),
                                                          c->n_valid_rots, c->root_post.as<float>() + (size_t)s * c->HW));
  }

  // per-part readout: the reference redoes it for every scale and keeps the last (findrot.cpp:257-259)
  std::vector<const float *> grids(P);
  for (int p = 0; p < P; ++p) grids[p] = c->POST(p, S - 1);
  if ((rc = enqueue_readout(c, grids, S - 1, flags, /*keys_ready=*/true))) return rc;
  c->result_scale = S - 1;
  if (flags & PS_INFER_KEEP_UNARIES) {
    PS_CUDA(c, cudaMemcpyAsync(c->unary.p, c->unary_backup.p, c->unary.bytes, cudaMemcpyDeviceToDevice, st));
    c->lattice_clean = lattice_was_clean;
  }
  return PS_OK;
}

// The legacy POS_GAUSSIAN model: mergeRotations + computeRootPosterior (objectdetect_findpos.cpp:118-170, :172-334).
// Rotations are summed out of every unary first, the upward pass then runs on 2-D grids with computePosJointMarginal
// (ps_pos_message on an internal single-slice context); there is no downward pass: the result is the root posterior of
// every scale and, with PS_INFER_ROOT_HYPS, its local maxima.
static int infer_pos(ps_ctx *c, int flags) {
  const int P = c->P, S = c->S, root = c->root;
  const size_t HW = c->HW;
  const bool sparse = flags & PS_INFER_SPARSE;
  cudaStream_t st = c->stream;
  c->have_result = false;
  if (!c->sub) {
    ps_config sc = c->cfg;
    sc.num_parts = 2;
    sc.num_rotation_steps = 1;
    sc.num_scale_steps = 1;
    sc.min_object_scale = sc.max_object_scale = 1.0f;
    sc.root_idx = 0;
    sc.keep_all_scales = 0;
    memset(sc.is_detect, 0, sizeof sc.is_detect);
    memset(sc.is_upright, 0, sizeof sc.is_upright);
    memset(sc.is_root, 0, sizeof sc.is_root);
    sc.is_detect[0] = sc.is_detect[1] = 1;
    sc.is_root[0] = 1;
    sc.strip_border_detections = 0;
    if (ps_create(&sc, &c->sub) != PS_OK) return c->fail(PS_ERR_CUDA, "internal 2-D context: %s", ps_last_error(nullptr));
  }
  if (int rc = ps_set_stream(c->sub, st)) return c->fail(rc, "internal 2-D context: %s", ps_last_error(c->sub));
  if (c->pos_merged.bytes < (size_t)P * HW * sizeof(float)) {
    PS_CUDA(c, cudaStreamSynchronize(st));
    PS_CUDA(c, c->pos_merged.alloc((size_t)P * HW * sizeof(float)));
    PS_CUDA(c, c->pos_post.alloc((size_t)P * HW * sizeof(float)));
    PS_CUDA(c, c->pos_msg.alloc(HW * sizeof(float)));
  }
  float *merged = c->pos_merged.as<float>(), *post = c->pos_post.as<float>(), *msg = c->pos_msg.as<float>();
  for (int s = 0; s < S; ++s) {
    const double scale = scale_of(c->cfg, s);
    for (int p = 0; p < P; ++p)  // mergeRotationsSum of every detected part (:146-165)
      if (c->cfg.is_detect[p])
        PS_LAUNCH(c, KC_MISC, psk::k_merge_rotations<<<cdiv(HW, 256), 256, 0, st>>>(c->U(p, s), c->R, HW, merged + (size_t)p * HW));
    PS_CUDA(c, cudaMemsetAsync(post, 0, (size_t)P * HW * sizeof(float), st));  // log_part_posterior starts at 0 (:214)
    // post-order walk with the reference's explicit stack (:223-301): a node waits for its first uncomputed child
    std::vector<char> uniform(P, 1), computed(P, 0);
    std::vector<int> stack(1, root);
    while (!stack.empty()) {
      const int cur = stack.back();
      stack.pop_back();
      bool can = true;
      for (size_t j = 0; j < c->joints.size() && can; ++j)
        if (c->joints[j].parent_idx == cur && !computed[c->joints[j].child_idx]) {
          can = false;
          stack.push_back(cur);
          stack.push_back(c->joints[j].child_idx);
        }
      if (!can) continue;
      for (size_t j = 0; j < c->joints.size(); ++j) {  // children in joint order (:252-257)
        const ps_joint &q = c->joints[j];
        if (q.parent_idx != cur || uniform[q.child_idx]) continue;  // a branch without detected parts sends nothing (:263)
        if (int rc = ps_pos_message(c->sub, post + (size_t)q.child_idx * HW, msg, PS_MEM_DEVICE, q.offset_p, q.C, scale, sparse ? 1 : 0))
          return c->fail(rc, "message of joint %zu: %s", j, ps_last_error(c->sub));
        PS_LAUNCH(c, KC_MISC, psk::k_add2<<<cdiv(HW, 256), 256, 0, st>>>(post + (size_t)cur * HW, msg, HW));
        uniform[cur] = 0;
      }
      if (c->cfg.is_detect[cur]) {  // :293-296
        PS_LAUNCH(c, KC_MISC, psk::k_add2<<<cdiv(HW, 256), 256, 0, st>>>(post + (size_t)cur * HW, merged + (size_t)cur * HW, HW));
        uniform[cur] = 0;
      }
      computed[cur] = 1;
    }
    PS_CUDA(c, cudaMemcpyAsync(c->root_post.as<float>() + (size_t)s * HW, post + (size_t)root * HW, HW * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
  }
  c->launches += c->sub->launches;
  c->sub->launches = 0;
  if (flags & PS_INFER_ROOT_HYPS) {  // findLocalMax(root_part_posterior, 1000), findpos.cpp:416-417
    int rc = ensure_topk(c, (size_t)P + 1, (size_t)std::max(c->cfg.roi_save_num_samples, 1000), std::max(c->N, (size_t)S * HW));
    if (rc) return rc;
    if ((rc = enqueue_local_max(c, c->root_post.as<float>(), S, c->H, c->W, 1000, P))) return rc;
    if ((rc = fetch_local_max(c, P + 1))) return rc;
  }
  c->pending_grids.clear();
  c->pending_scaleidx = S - 1;
  c->pending_flags = flags & PS_INFER_ROOT_HYPS;
  c->pending_root_only = true;
  c->result_pending = true;
  c->result_scale = S - 1;
  return PS_OK;
}

// computeRootPosteriorRot + computePartMarginals (findrot.cpp:470-727, :124-286).  The schedule of one inference is a
// fixed sequence of launches over ctx-owned buffers as long as the joints, the flags and the way each leaf's maximum is
// obtained stay the same, so it is captured into a CUDA graph the second time such an inference is requested and
// replayed from then on (PSINFER_NO_GRAPH=1 keeps the eager launches; profiling always runs eagerly).
int ps_infer(ps_ctx *c, int flags) {
  if (!c) return PS_ERR_INVALID;
  if (!c->joints_set) return c->fail(PS_ERR_STATE, "ps_infer before ps_set_joints");
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  if (c->pos_model) return infer_pos(c, flags);
  c->pending_root_only = false;
  c->joints_reused = c->last_infer_version < 0 || c->last_infer_version == c->joints_version;
  c->last_infer_version = c->joints_version;
  if (c->disable_graph || c->profiling) return infer_enqueue(c, flags);
  std::string key((const char *)&flags, sizeof flags);
  key.append((const char *)&c->joints_version, sizeof c->joints_version);
  key.append((const char *)c->unary_max_valid.data(), c->unary_max_valid.size());
  auto it = c->graphs.find(key);
  if (it == c->graphs.end()) {  // first sight: eager, so that every lazy allocation and scatter map exists afterwards
    if (c->graphs.size() >= 16) {
      // keys seen once (a joint set that changes with every image never comes back) cost nothing to forget; graphs that
      // were instantiated may still be running
      for (auto i = c->graphs.begin(); i != c->graphs.end();) i = i->second.exec ? std::next(i) : c->graphs.erase(i);
      if (c->graphs.size() >= 16) {
        PS_CUDA(c, cudaStreamSynchronize(c->stream));
        for (auto &g : c->graphs) cudaGraphExecDestroy(g.second.exec);
        c->graphs.clear();
      }
    }
    c->graphs[key] = ps_ctx::InferGraph();
    return infer_enqueue(c, flags);
  }
  ps_ctx::InferGraph &g = it->second;
  if (g.failed) return infer_enqueue(c, flags);
  if (!g.exec) {
    const long long l0 = c->launches;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      g.failed = true;
      return infer_enqueue(c, flags);
    }
    const int rc = infer_enqueue(c, flags);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
    if (rc != PS_OK || ce != cudaSuccess || !graph ||
        cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      g.exec = nullptr;
      g.failed = true;
      c->launches = l0;
      c->result_pending = false;
      if (rc != PS_OK) return rc;
      return infer_enqueue(c, flags);  // something on the path cannot be captured: stay eager for this key
    }
    cudaGraphDestroy(graph);
    g.launches = c->launches - l0;
    g.pending_grids = c->pending_grids;
    g.pending_scaleidx = c->pending_scaleidx;
    g.pending_flags = c->pending_flags;
    g.result_scale = c->result_scale;
    c->launches = l0;
  }
  PS_CUDA(c, cudaGraphLaunch(g.exec, c->stream));
  c->launches += g.launches;
  ++c->graph_replays;
  c->pending_grids = g.pending_grids;
  c->pending_scaleidx = g.pending_scaleidx;
  c->pending_flags = g.pending_flags;
  c->result_scale = g.result_scale;
  c->result_pending = true;
  c->have_result = false;
  return PS_OK;
}

int ps_max_states(ps_ctx *c, int flags) {
  if (!c) return PS_ERR_INVALID;
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  std::vector<const float *> grids(c->P);
  for (int p = 0; p < c->P; ++p) grids[p] = c->U(p, 0);
  int rc = enqueue_readout(c, grids, 0, flags & PS_INFER_LOCAL_MAX);
  if (rc) return rc;
  c->result_scale = -2;  // marginals are not available after getMaxStates
  return PS_OK;
}

int ps_get_best_conf(ps_ctx *c, float *out) {
  if (!c || !out) return PS_ERR_INVALID;
  if (int rc = finish_result(c)) return rc;
  if (c->pending_root_only)
    return c->fail(PS_ERR_STATE, "no part estimates: the POS_GAUSSIAN path has no downward pass (findpos.cpp:172-334)");
  memcpy(out, c->best_conf.data(), c->best_conf.size() * sizeof(float));
  return PS_OK;
}

int ps_get_part_hyps(ps_ctx *c, int part, float *out, int cap, int *count) {
  if (!c || !out || !count) return PS_ERR_INVALID;
  if (int rc = finish_result(c)) return rc;
  if (part < 0 || part >= c->P) return c->fail(PS_ERR_INVALID, "part out of range");
  int n = std::min((int)(c->part_hyps[part].size() / PS_HYP_VEC), cap);
  memcpy(out, c->part_hyps[part].data(), (size_t)n * PS_HYP_VEC * sizeof(float));
  *count = n;
  return PS_OK;
}

int ps_get_marginal(ps_ctx *c, int part, int scale, float *dst, int mem_kind) {
  if (!c || !dst) return PS_ERR_INVALID;
  if (int rc = finish_result(c)) return rc;
  if (c->result_scale < 0) return c->fail(PS_ERR_STATE, "no marginals: call ps_infer first");
  if (part < 0 || part >= c->P || scale < 0 || scale >= c->S) return c->fail(PS_ERR_INVALID, "part/scale out of range");
  if (c->pending_root_only) return c->fail(PS_ERR_STATE, "no marginals on the POS_GAUSSIAN path");
  if (!c->cfg.keep_all_scales && scale != c->S - 1)
    return c->fail(PS_ERR_STATE, "only the last scale is resident; create the ctx with keep_all_scales");
  PS_CUDA(c, cudaMemcpyAsync(dst, c->POST(part, scale), c->N * sizeof(float),
                             mem_kind == PS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PS_OK;
}

int ps_get_root_posterior(ps_ctx *c, float *dst, int mem_kind) {
  if (!c || !dst) return PS_ERR_INVALID;
  if (int rc = finish_result(c)) return rc;
  if (c->result_scale < 0) return c->fail(PS_ERR_STATE, "no root posterior: call ps_infer first");
  PS_CUDA(c, cudaMemcpyAsync(dst, c->root_post.p, c->HW * c->S * sizeof(float),
                             mem_kind == PS_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PS_OK;
}

int ps_get_root_hyps(ps_ctx *c, float *out, int cap, int *count) {
  if (!c || !out || !count) return PS_ERR_INVALID;
  if (int rc = finish_result(c)) return rc;
  if (!c->have_root_hyps) return c->fail(PS_ERR_STATE, "no root hypotheses: ps_infer with PS_INFER_ROOT_HYPS");
  int n = std::min((int)(c->root_hyps.size() / 4), cap);
  memcpy(out, c->root_hyps.data(), (size_t)n * 4 * sizeof(float));
  *count = n;
  return PS_OK;
}

// ---- test seams -----------------------------------------------------------------------------------

int ps_message(ps_ctx *c, const float *child, float *parent, int mem_kind, const double off_in[2],
               const double off_out[2], const double C[4], double rot_mean, double rot_sigma, double scale, int sparse) {
  if (!c || !child || !parent || !off_in || !off_out || !C) return PS_ERR_INVALID;
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  std::shared_ptr<DevPlan> plan;
  int prc = get_plan(c, off_in, off_out, C, rot_mean, rot_sigma, scale, plan);
  if (prc) return prc;
  DevPlan &dp = *plan;
  const float *din = child;
  float *dout = parent;
  if (mem_kind == PS_MEM_HOST) {
    PS_CUDA(c, cudaMemcpyAsync(c->tmp[0].p, child, c->N * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    din = c->tmp[0].as<float>();
    dout = c->tmp[1].as<float>();
  }
  int *mx = c->MAXP(2 * c->P + psk::kMaxRootChildren + 2);
  int rc = grid_max(c, din, c->N, mx);
  if (rc) return rc;
  Sink sink;
  sink.out0 = dout;
  {
    std::vector<MsgJob> jobs(1);
    jobs[0].dp = &dp; jobs[0].in = din; jobs[0].in_max = mx; jobs[0].sparse = sparse != 0; jobs[0].sink = sink;
    if ((rc = run_level(c, jobs))) return rc;
  }
  if (mem_kind == PS_MEM_HOST)
    PS_CUDA(c, cudaMemcpyAsync(parent, dout, c->N * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PS_OK;
}

int ps_pos_message(ps_ctx *c, float *child, float *parent, int mem_kind, const double offset[2], const double C[4],
                   double scale, int sparse) {
  if (!c || !child || !parent || !offset || !C) return PS_ERR_INVALID;
  if (!(scale > 0)) return c->fail(PS_ERR_INVALID, "scale must be > 0 (objectdetect_findpos.cpp:68)");
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  const double zero[2] = {0.0, 0.0};
  std::shared_ptr<DevPlan> plan;
  int prc = get_plan(c, zero, zero, C, 0.0, 0.0, scale, plan, offset);
  if (prc) return prc;
  float *din = child, *dout = parent;
  if (mem_kind == PS_MEM_HOST) {
    PS_CUDA(c, cudaMemcpyAsync(c->tmp[0].p, child, c->N * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    din = c->tmp[0].as<float>();
    dout = c->tmp[1].as<float>();
  }
  // computeExpGrid has no max shift on this path (:76): M = +0, so exp(x + -0) = exp(x) and log(d) + 0 = log(d)
  int *mx = c->MAXP(2 * c->P + psk::kMaxRootChildren + 2);
  PS_LAUNCH(c, KC_MISC, psk::k_set_int<<<1, 32, 0, c->stream>>>(mx, 1, psk::enc_f(0.0f)));
  Sink sink;
  sink.out0 = dout;
  std::vector<MsgJob> jobs(1);
  jobs[0].dp = plan.get(); jobs[0].in = din; jobs[0].in_max = mx; jobs[0].sparse = sparse != 0; jobs[0].sink = sink;
  int rc = run_level(c, jobs);
  if (rc) return rc;
  // the reference leaves log(exp(child)) in its first argument (:88)
  PS_LAUNCH(c, KC_MISC, psk::k_exp_log<<<std::min(cdiv(c->N, 256), (unsigned)c->num_sms * 16), 256, 0, c->stream>>>(din, c->N));
  if (mem_kind == PS_MEM_HOST) {
    PS_CUDA(c, cudaMemcpyAsync(parent, dout, c->N * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    PS_CUDA(c, cudaMemcpyAsync(child, din, c->N * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  }
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PS_OK;
}

int ps_get_plan_info(ps_ctx *c, int joint, int downward, int scale, int out[10]) {
  if (!c || !out) return PS_ERR_INVALID;
  if (!c->joints_set) return c->fail(PS_ERR_STATE, "ps_get_plan_info before ps_set_joints");
  if (joint < 0 || joint >= (int)c->joints.size() || scale < 0 || scale >= c->S)
    return c->fail(PS_ERR_INVALID, "joint/scale out of range");
  const psg::MessagePlan &h = c->plans[((size_t)joint * 2 + (downward ? 1 : 0)) * c->S + scale]->host;
  out[0] = h.diag ? 1 : 0;
  out[1] = h.diag ? c->H : h.EH;
  out[2] = h.diag ? c->W : h.EW;
  out[3] = h.rot_mode == 1 ? (int)h.rot_taps.size() : 0;
  out[4] = (int)h.fx.size();
  out[5] = (int)h.fy.size();
  out[6] = h.rot_shift;
  out[7] = (h.in_pure ? 1 : 0) | (h.out_pure ? 2 : 0);
  const bool lists = !h.diag && !c->disable_tile_lists && !c->disable_tma;
  const bool fused = lists && !c->disable_batch && h.fcells_y;  // the level-batched route filters the walks' cells
  out[8] = fused ? (int)h.fcells_x : lists && h.xcells ? (int)h.xcells : out[1] * out[2];
  out[9] = fused ? (int)h.fcells_y : lists && h.ycells ? (int)h.ycells : out[1] * out[2];
  return PS_OK;
}

int ps_selftest_math(ps_ctx *c, unsigned first_bits, unsigned long long count, unsigned long long out[4]) {
  if (!c || !out) return PS_ERR_INVALID;
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  DevBuf d;
  PS_CUDA(c, d.alloc(4 * sizeof(unsigned long long)));
  PS_CUDA(c, cudaMemsetAsync(d.p, 0, 4 * sizeof(unsigned long long), c->stream));
  PS_LAUNCH(c, KC_MISC, psk::k_selftest_math<<<c->num_sms * 16, 256, 0, c->stream>>>(first_bits, count, d.as<unsigned long long>()));
  PS_CUDA(c, cudaMemcpyAsync(out, d.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PS_OK;
}

int ps_eval_math(ps_ctx *c, int op, unsigned first_bits, unsigned count, float *out_host) {
  if (!c || !out_host || op < 0 || op > 3) return PS_ERR_INVALID;
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  DevBuf d;
  PS_CUDA(c, d.alloc((size_t)count * sizeof(float)));
  PS_LAUNCH(c, KC_MISC, psk::k_eval_math<<<c->num_sms * 16, 256, 0, c->stream>>>(op, first_bits, count, d.as<float>()));
  PS_CUDA(c, cudaMemcpyAsync(out_host, d.p, (size_t)count * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  PS_CUDA(c, cudaStreamSynchronize(c->stream));
  return PS_OK;
}

int ps_find_local_max(ps_ctx *c, const float *grid, int mem_kind, int d0, int h, int w, int max_n, float *out,
                      int *count) {
  if (!c || !grid || !out || !count || d0 < 1 || h < 1 || w < 1 || max_n < 0) return PS_ERR_INVALID;
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  size_t n = (size_t)d0 * h * w;
  DevBuf tmp;
  const float *dg = grid;
  if (mem_kind == PS_MEM_HOST) {
    PS_CUDA(c, tmp.alloc(n * sizeof(float)));
    PS_CUDA(c, cudaMemcpyAsync(tmp.p, grid, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    dg = tmp.as<float>();
  }
  std::vector<float> rows;
  int rc = local_max_device(c, dg, d0, h, w, max_n, rows);
  if (rc) return rc;
  memcpy(out, rows.data(), rows.size() * sizeof(float));
  *count = (int)(rows.size() / 4);
  return PS_OK;
}

int ps_unary_local_max(ps_ctx *c, int part, int scale, int max_n, float *out, int *count) {
  if (!c || !out || !count || max_n < 0) return PS_ERR_INVALID;
  if (part < 0 || part >= c->P || scale < 0 || scale >= c->S) return c->fail(PS_ERR_INVALID, "part/scale out of range");
  PS_CUDA(c, cudaSetDevice(c->cfg.device));
  std::vector<float> rows;
  int rc = local_max_device(c, c->U(part, scale), c->R, c->H, c->W, max_n, rows);
  if (rc) return rc;
  memcpy(out, rows.data(), rows.size() * sizeof(float));
  *count = (int)(rows.size() / 4);
  return PS_OK;
}

}  // extern "C"
