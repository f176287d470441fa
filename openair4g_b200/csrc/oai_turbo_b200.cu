// Host side of the B200 LTE turbo-decoding engine: GPU context, batch plans, and the C ABI
// declared in include/oai_turbo_b200.h.  No torch, no CPU compute path: every entry point
// launches the sm_100a kernels in td16_map.cuh / td16_xchg.cuh (and friends) or fails.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <tuple>
#include <utility>
#include <algorithm>
#include <chrono>

#include "../../include/oai_turbo_b200.h"
#include "td_common.cuh"
#include "td16_map.cuh"
#include "td16_xchg.cuh"
#include "td16_sw.cuh"
#include "rm_kernels.cuh"
#include "td8_kernels.cuh"
#include "tx_kernels.cuh"
#include "tb_kernels.cuh"
#include "ul_kernels.cuh"

namespace oai {

// host_pack.cpp: int16 -> int8 range check + pack on the caller's CPU cores (the narrow input feed)
int host_pack_threads();
int host_pack_i16_to_i8(const int16_t* src, int8_t* dst, size_t n);

static const uint16_t kQpp[188][2] = {
#include "qpp_table.inc"
};

static std::atomic<unsigned long long> g_launches{0};
static const bool g_use_graphs = [] { const char* e = getenv("OAI_TURBO_NO_GRAPHS"); return !(e && e[0] == '1'); }();
// tracked fast passes beyond the a-priori guard (td16_map.cuh); OAI_TURBO_NO_TRACK=1 restores guard-or-exact
static const int g_track = [] { const char* e = getenv("OAI_TURBO_NO_TRACK"); return (e && e[0] == '1') ? 0 : 1; }();
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  fprintf(stderr, "[oai_turbo_b200] ERROR: %s\n", g_err);
  return code;
}
#define CU(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) return fail(-100, "%s -> %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// K -> row of the 188-entry table (same bucketing as dlsch_decoding.c:314-325); -1 if illegal
static int qpp_index(int K) {
  if (K < 40 || K > 6144 || (K & 7)) return -1;
  int kb = K >> 3;
  if (kb <= 64) return kb - 5;
  if (kb <= 128) return (K & 15) ? -1 : 59 + ((kb - 64) >> 1);
  if (kb <= 256) return (K & 31) ? -1 : 91 + ((kb - 128) >> 2);
  return (K & 63) ? -1 : 123 + ((kb - 256) >> 3);
}
static int qpp_K(int idx) {
  if (idx < 60) return 40 + 8 * idx;
  if (idx < 92) return 512 + 16 * (idx - 59);
  if (idx < 124) return 1024 + 32 * (idx - 91);
  return 2048 + 64 * (idx - 123);
}

// ---- per-device context: read-only tables ------------------------------------------
struct DevCtx {
  int dev = -1;
  uint16_t* pi_pool = nullptr;        // H tables of all 188 K (natural position -> C4 halfword index)
  uint16_t* t_pool = nullptr;         // T tables (layout order -> C4 index of the QPP image)
  uint32_t pi_off[188], t_off[188];
  uint16_t* qpp_pool = nullptr;       // plain QPP tables pi[i] (8-bit decoder kernels)
  uint32_t qpp_off[188];
  uint16_t* t8_pool = nullptr;        // 8-bit decoder: T8 tables (C8 byte index -> C8 byte index of the QPP image)
  uint32_t t8_off[188];
  u32* crc_xp = nullptr;              // [4][32][CRC_NM] powers of x mod the CRC polynomials
  u32* sw_tab = nullptr;              // sliding-window mode (td16_sw.cuh): interleaver tables in the window layout, all K
  u32* sw_tab_off = nullptr;          // [769] word offset of K's table, indexed by K >> 3
  bool ok = false;
  unsigned gen = 0;                   // bumped when the tables are (re)built: cached launch graphs hold table pointers
  // rate-dematching prefix-count tables, one per (K, F) seen so far: cnt[i] = number of circular-buffer slots in [0, i)
  // that carry a bit (generate_dummy_w's NULL map, lte_rate_matching.c:293-382); built on the host at first use
  uint16_t* rm_tab = nullptr;
  uint32_t rm_tab_used = 0;
  std::unordered_map<uint32_t, uint32_t> rm_tab_off;     // (K << 8 | F) -> halfword offset
};
constexpr uint32_t RM_TAB_CAP = 64u * 18560u;            // halfwords: 64 tables of the largest size (2.4 MB)
static DevCtx g_ctx[16];
static std::mutex g_ctx_mu;

// offset of the (K, F) prefix-count table in c->rm_tab, building it if needed; 0xffffffff when the pool is full (the
// kernel then derives the ranks itself).  The caller holds a DevGuard on c->dev.
static uint32_t rm_table(DevCtx* c, uint32_t K, uint32_t F) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  const uint32_t key = (K << 8) | (F & 0xffu);
  auto it = c->rm_tab_off.find(key);
  if (it != c->rm_tab_off.end()) return it->second;
  const uint32_t D = K + 4, RTC = (D + 31) / 32, Kpi = 32 * RTC, ND = Kpi - D, n = 3 * Kpi + 1;
  const uint32_t need = (n + 7) & ~7u;
  if (!c->rm_tab) {
    if (cudaMalloc(&c->rm_tab, sizeof(uint16_t) * RM_TAB_CAP) != cudaSuccess) { cudaGetLastError(); c->rm_tab = nullptr; return 0xffffffffu; }
  }
  if (c->rm_tab_used + need > RM_TAB_CAP) return 0xffffffffu;
  std::vector<uint16_t> cnt(need, 0);
  const uint32_t magic = 0xffffffffu / RTC + 1;
  uint32_t run = 0;
  for (uint32_t i = 0; i < 3 * Kpi; ++i) { cnt[i] = (uint16_t)run; run += dummy_is_null(i, RTC, Kpi, ND, F, magic) ? 0u : 1u; }
  for (uint32_t i = 3 * Kpi; i < need; ++i) cnt[i] = (uint16_t)run;
  const uint32_t off = c->rm_tab_used;
  if (cudaMemcpy(c->rm_tab + off, cnt.data(), sizeof(uint16_t) * need, cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); return 0xffffffffu; }
  c->rm_tab_used += need;
  c->rm_tab_off[key] = off;
  return off;
}

static u32 gf_xtimes(u32 r, u32 poly, int w) {
  u32 top = r & (1u << (w - 1));
  r = (r << 1) & ((w == 32) ? 0xffffffffu : ((1u << w) - 1));
  return top ? (r ^ poly) : r;
}

// sets the current device for the lifetime of the object and puts the caller's device back afterwards: no entry point
// of the library changes the calling thread's current device
struct DevGuard {
  int prev = -1;
  bool switched = false;
  int enter(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) return -1;
    if (dev >= 0 && dev != prev) { if (cudaSetDevice(dev) != cudaSuccess) return -1; switched = true; }
    return 0;
  }
  ~DevGuard() { if (switched) cudaSetDevice(prev); }
};

static int ctx_get(int dev, DevCtx** out) {
  if (dev < 0) CU(cudaGetDevice(&dev));
  if (dev >= 16) return fail(-2, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  DevCtx& c = g_ctx[dev];
  if (!c.ok) {
    DevGuard guard;
    if (guard.enter(dev)) return fail(-100, "cannot select CUDA device %d", dev);
    std::vector<uint16_t> pool, tpool, qpool, t8pool;
    for (int i = 0; i < 188; ++i) {
      const int K = qpp_K(i), W = K / 8, A = c4_words(W) * 2;
      c.pi_off[i] = (uint32_t)pool.size();
      c.t_off[i] = (uint32_t)tpool.size();
      c.qpp_off[i] = (uint32_t)qpool.size();
      const uint64_t f1 = kQpp[i][0], f2 = kQpp[i][1];
      std::vector<uint16_t> H(K), T(A);
      for (int j = 0; j < K; ++j) H[j] = (uint16_t)c4_hw(j % W, j / W);
      for (int h = 0; h < A; ++h) T[h] = (uint16_t)h;                       // padding maps to itself
      for (uint64_t j = 0; j < (uint64_t)K; ++j) {
        const uint64_t pj = (f1 * j + f2 * j * j) % (uint64_t)K;           // pi(j), 36.212 5.1.3.2.3
        T[H[j]] = H[pj];
        qpool.push_back((uint16_t)pj);
      }
      c.t8_off[i] = (uint32_t)t8pool.size();
      if (K >= 256 && (K & 15) == 0) {                                       // the 8-bit decoder's domain
        const int W8 = K / 16, A8 = c8_bytes(W8);
        std::vector<uint16_t> T8(A8);
        for (int h = 0; h < A8; ++h) T8[h] = (uint16_t)h;
        for (uint64_t j = 0; j < (uint64_t)K; ++j) {
          const uint64_t pj = (f1 * j + f2 * j * j) % (uint64_t)K;
          T8[h8((int)(j % W8), (int)(j / W8))] = (uint16_t)h8((int)(pj % W8), (int)(pj / W8));
        }
        t8pool.insert(t8pool.end(), T8.begin(), T8.end());                   // A8 is a multiple of 128
      }
      pool.insert(pool.end(), H.begin(), H.end());
      while (pool.size() & 7) pool.push_back(0);
      tpool.insert(tpool.end(), T.begin(), T.end());                         // A is a multiple of 32
    }
    CU(cudaMalloc(&c.pi_pool, pool.size() * sizeof(uint16_t)));
    CU(cudaMemcpy(c.pi_pool, pool.data(), pool.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&c.qpp_pool, qpool.size() * sizeof(uint16_t)));
    CU(cudaMemcpy(c.qpp_pool, qpool.data(), qpool.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&c.t8_pool, t8pool.size() * sizeof(uint16_t)));
    CU(cudaMemcpy(c.t8_pool, t8pool.data(), t8pool.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&c.t_pool, tpool.size() * sizeof(uint16_t)));
    CU(cudaMemcpy(c.t_pool, tpool.data(), tpool.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    // powers of x modulo the four CRC polynomials (crc_byte.c:53-57): entry [t][r][m] = x^(w + r + 32 m) mod P,
    // the weight of a 32-bit message word whose last bit is r + 32 m bits before the message end
    std::vector<u32> xp(4 * 32 * CRC_NM);
    const u32 polys[4] = {0x864cfbu, 0x800063u, 0x1021u, 0x9Bu};
    const int ws[4] = {24, 24, 16, 8};
    for (int t = 0; t < 4; ++t) {
      u32 r = 1;
      for (int i = 0; i < ws[t]; ++i) r = gf_xtimes(r, polys[t], ws[t]);    // x^w
      std::vector<u32> pw(32 * CRC_NM + 32);
      for (size_t e = 0; e < pw.size(); ++e) { pw[e] = r; r = gf_xtimes(r, polys[t], ws[t]); }   // x^(w+e)
      for (int rr = 0; rr < 32; ++rr)
        for (int m = 0; m < CRC_NM; ++m) xp[(t * 32 + rr) * CRC_NM + m] = pw[rr + 32 * m];
    }
    {
      // sliding-window mode: entry [o][t] = shared-memory halfword indices of pi(j) for j = (2t) WL + o and
      // (2t+1) WL + o, the thread's two windows at step o
      std::vector<u32> swt, swo(769, 0);
      for (int i = 0; i < 188; ++i) {
        const int K = qpp_K(i), NW = sw_windows(K), WL = K / NW, LPB = NW / 2;
        const uint64_t f1 = kQpp[i][0], f2 = kQpp[i][1];
        swo[K >> 3] = (u32)swt.size();
        auto idx = [&](uint64_t j) -> u32 {                // 2 x (step' << 6 | ((lane' + step') & 31) << 1 | half)  (sw_idx)
          const uint64_t pj = (f1 * j + f2 * j * j) % (uint64_t)K;
          const u32 o = (u32)(pj % WL), w = (u32)(pj / WL);
          return 2u * ((o << 6) | ((((w >> 1) + o) & 31u) << 1) | (w & 1u));      // byte offset of the int16 element
        };
        for (int o = 0; o < WL; ++o)
          for (int t = 0; t < LPB; ++t) swt.push_back(idx((uint64_t)(2 * t) * WL + o) | (idx((uint64_t)(2 * t + 1) * WL + o) << 16));
        auto nat = [&](uint64_t j) -> u32 { return (u32)((f1 * j + f2 * j * j) % (uint64_t)K); };      // then pi(j) itself, same order
        for (int o = 0; o < WL; ++o)
          for (int t = 0; t < LPB; ++t) swt.push_back(nat((uint64_t)(2 * t) * WL + o) | (nat((uint64_t)(2 * t + 1) * WL + o) << 16));
      }
      CU(cudaMalloc(&c.sw_tab, swt.size() * sizeof(u32)));
      CU(cudaMemcpy(c.sw_tab, swt.data(), swt.size() * sizeof(u32), cudaMemcpyHostToDevice));
      CU(cudaMalloc(&c.sw_tab_off, swo.size() * sizeof(u32)));
      CU(cudaMemcpy(c.sw_tab_off, swo.data(), swo.size() * sizeof(u32), cudaMemcpyHostToDevice));
      CU(cudaFuncSetAttribute(k_turbo_sw, cudaFuncAttributeMaxDynamicSharedMemorySize, sw_smem_bytes(6144)));
    }
    CU(cudaMalloc(&c.crc_xp, xp.size() * sizeof(u32)));
    CU(cudaMemcpy(c.crc_xp, xp.data(), xp.size() * sizeof(u32), cudaMemcpyHostToDevice));
    CU(cudaFuncSetAttribute(k_map16<MAP_SEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAP_SMEM_BYTES));
    CU(cudaFuncSetAttribute(k_demux16_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 6208 * (int)sizeof(int16_t)));
    c.dev = dev;
    c.ok = true;
    ++c.gen;
  }
  *out = &c;
  return 0;
}

// free_td16 / free_td8: releases the tables of every device (the next call rebuilds them)
static void ctx_release_all() {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  for (DevCtx& c : g_ctx) {
    if (!c.ok) continue;
    cudaFree(c.pi_pool); cudaFree(c.t_pool); cudaFree(c.qpp_pool); cudaFree(c.t8_pool); cudaFree(c.crc_xp);
    cudaFree(c.sw_tab); cudaFree(c.sw_tab_off); c.sw_tab = c.sw_tab_off = nullptr;
    if (c.rm_tab) cudaFree(c.rm_tab);
    c.rm_tab = nullptr; c.rm_tab_used = 0; c.rm_tab_off.clear();
    c.pi_pool = c.t_pool = c.qpp_pool = c.t8_pool = nullptr; c.crc_xp = nullptr;
    c.ok = false;
  }
  cudaGetLastError();
}

// ---- a batch of code blocks resident on one GPU ---------------------------------------
constexpr int CKPT_S = MAP_CKPT_STEPS;
constexpr int GUARD_B = 2954;     // 11 B + 256 <= 32767: the reference cannot saturate (DESIGN.md "fast-path guard"); our own no-wrap rule then allows M = B+1 <= 2499
constexpr int MAX_PARTS = 8;      // pipeline stages of one host batch (copy of part i+1 overlaps the decode of part i)
constexpr int MIN_PART_BLOCKS = 2960;   // smaller parts leave the MAP kernel latency-bound (measured: 16 parts 25 ms, 8 parts 18.6 ms per 23680 blocks)
constexpr int PART_STREAMS = 3;   // the parts of a pipelined batch are decoded on this many compute streams in turn: a part fills
                                  // a fraction of the SMs (one K=6144 wave of the MAP kernel is 14 208 blocks), consecutive
                                  // parts overlap on the GPU instead of queueing behind each other

// Narrow input feed: measured pack throughput decides whether it stays on.  Below ~1.3x the link rate packing cannot win
// (several ranks sharing the host's cores, a small VM); the feed then pauses for a while and is probed again.
static std::atomic<int> g_pack_pause{0};
static double pack_min_gbs() { const char* e = getenv("OAI_TURBO_PACK_MIN_GBS"); return (e && *e) ? atof(e) : 70.0; }

// optional per-launch CUDA-event timing (bench.py's roofline leg): class 0 demux, 1 map, 2 x1, 3 x2
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> ev;          // pairs
  std::vector<int> cls;
  size_t used = 0;
  double ms[4] = {0, 0, 0, 0};
  long count[4] = {0, 0, 0, 0};
  void begin(int c, cudaStream_t st) {
    if (!on) return;
    if (used + 2 > ev.size()) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); ev.push_back(a); ev.push_back(b); }
    cls.push_back(c);
    cudaEventRecord(ev[used], st);
  }
  void end(cudaStream_t st) {
    if (!on) return;
    cudaEventRecord(ev[used + 1], st);
    used += 2;
  }
  void collect() {                        // call after the stream is synchronised
    for (size_t i = 0; i < used; i += 2) {
      float t = 0;
      if (cudaEventElapsedTime(&t, ev[i], ev[i + 1]) == cudaSuccess) { ms[cls[i / 2]] += t; count[cls[i / 2]]++; }
    }
    used = 0; cls.clear();
  }
  void reset() { collect(); for (int i = 0; i < 4; ++i) { ms[i] = 0; count[i] = 0; } }
  ~Profiler() { for (auto e : ev) cudaEventDestroy(e); }
};

// Either launches on a stream or appends nodes to a CUDA graph (a linear chain).  The graph is built explicitly, not by
// stream capture: a capture in progress makes device-wide calls of OTHER threads fail (cudaDeviceSynchronize returns
// "operation not permitted when stream is capturing"), which a library with concurrent callers cannot impose.
struct Launcher {
  cudaStream_t st = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphNode_t last = nullptr;
  bool ok = true;
  template <class... P, class... A>
  void run(void (*k)(P...), dim3 grid, dim3 block, size_t smem, A&&... a) {
    if (!graph) { k<<<grid, block, smem, st>>>(std::forward<A>(a)...); return; }
    std::tuple<typename std::decay<P>::type...> vals(std::forward<A>(a)...);
    void* ptrs[sizeof...(P)];
    fill(ptrs, vals, std::index_sequence_for<P...>{});
    cudaKernelNodeParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.func = (void*)k; kp.gridDim = grid; kp.blockDim = block; kp.sharedMemBytes = (unsigned)smem; kp.kernelParams = ptrs;
    cudaGraphNode_t node;
    if (cudaGraphAddKernelNode(&node, graph, last ? &last : nullptr, last ? 1 : 0, &kp) != cudaSuccess) { ok = false; return; }
    last = node;
  }
  template <class T, size_t... I>
  static void fill(void** ptrs, T& vals, std::index_sequence<I...>) { ((ptrs[I] = (void*)&std::get<I>(vals)), ...); }
  void zero_ints(int* dst, int count) {
    if (!graph) { cudaMemsetAsync(dst, 0, count * sizeof(int), st); return; }
    cudaMemsetParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.dst = dst; mp.value = 0; mp.elementSize = 4; mp.width = count; mp.height = 1; mp.pitch = 0;
    cudaGraphNode_t node;
    if (cudaGraphAddMemsetNode(&node, graph, last ? &last : nullptr, last ? 1 : 0, &mp) != cudaSuccess) { ok = false; return; }
    last = node;
  }
};

struct Batch {
  Profiler prof;
  DevCtx* ctx = nullptr;
  int cap = 0, n = 0, A = 0, max_iter = 0, max_K = 0;
  int cur_max_K = 6144;        // largest K of the batch that is loaded (set_meta): sizes the exchange CTAs
  int sw_mode = 0;             // the loaded batch is decoded in the optional sliding-window mode (td16_sw.cuh)
  int cur_sw_smem = 0;         // its shared-memory need (largest window length of the batch)
  long slot_hw = 0, ckpt_words = 0;
  CbMeta* d_meta = nullptr;
  CbState* d_state = nullptr;
  int16_t* d_ws = nullptr;
  u32* d_ckpt = nullptr;
  int* d_batch_max = nullptr;
  int* d_active = nullptr;     // compacted indices of the running blocks (per pipeline part: relative to the part)
  int* d_nactive = nullptr;    // one counter per part
  std::vector<CbMeta> h_meta;
  // cached CUDA graphs of the launch sequence for small batches
  static constexpr int GRAPH_MAX_BLOCKS = 1024;
  struct GraphKey {
    const void *in, *out, *status, *fe_rm, *fe_w, *fe_harq;
    int lo, n, part, max_iter, max_K;
    unsigned gen;                      // DevCtx::gen the graph was built against
    int in8 = 0;
    int cur_K = 0;                     // largest K of the loaded batch (sizes the exchange CTAs)
    int sw = 0;                        // sliding-window mode (and its shared-memory size)
    bool operator==(const GraphKey& o) const {
      if (sw != o.sw) return false;
      return in == o.in && out == o.out && status == o.status && fe_rm == o.fe_rm && fe_w == o.fe_w && fe_harq == o.fe_harq &&
             lo == o.lo && n == o.n && part == o.part && max_iter == o.max_iter && max_K == o.max_K && gen == o.gen && in8 == o.in8 && cur_K == o.cur_K;
    }
  };
  struct GraphEntry { GraphKey key; cudaGraphExec_t exec; int launches; };
  std::vector<GraphEntry> graphs;
  void drop_graphs() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }

  int alloc(DevCtx* c, int ncb, int Kmax) {
    ctx = c;
    cap = ncb;
    int W = Kmax / 8;
    max_K = Kmax;
    A = c4_words(W) * 2;
    slot_hw = (long)ARR_COUNT * A;
    ckpt_words = (long)((W + CKPT_S - 1) / CKPT_S + 1) * 32;
    CU(cudaMalloc(&d_meta, sizeof(CbMeta) * ncb));
    CU(cudaMalloc(&d_state, sizeof(CbState) * ncb));
    CU(cudaMalloc(&d_ws, sizeof(int16_t) * slot_hw * ncb));
    CU(cudaMalloc(&d_ckpt, sizeof(u32) * ckpt_words * ncb));
    CU(cudaMalloc(&d_batch_max, sizeof(int) * MAX_PARTS));
    CU(cudaMalloc(&d_active, sizeof(int) * ncb * 4));            // two lists (current / next) of 2 n entries per part (k_compact)
    CU(cudaMalloc(&d_nactive, sizeof(int) * MAX_PARTS * 8));     // per part: two lists x three class counters (+ pad)
    CU(cudaMemset(d_ws, 0, sizeof(int16_t) * slot_hw * ncb));
    CU(cudaMemset(d_state, 0, sizeof(CbState) * ncb));
    // cudaMemset on device memory is asynchronous and runs on the legacy default stream, which does not order with the
    // non-blocking streams the decodes use: without this wait the zeroing could land AFTER the first kernels' writes
    CU(cudaDeviceSynchronize());
    return 0;
  }
  void release() {
    drop_graphs();
    cudaFree(d_meta); cudaFree(d_state); cudaFree(d_ws); cudaFree(d_ckpt); cudaFree(d_batch_max); cudaFree(d_active); cudaFree(d_nactive);
    d_meta = nullptr; d_state = nullptr; d_ws = nullptr; d_ckpt = nullptr; d_batch_max = nullptr; d_active = nullptr; d_nactive = nullptr;
  }
  // staged: page-locked copy of m.data() (asynchronous transfer), or nullptr: from h_meta (pageable; the call then blocks
  // until the stream reaches the copy)
  int set_meta(const std::vector<CbMeta>& m, cudaStream_t st, const void* staged = nullptr) {
    h_meta = m;
    n = (int)m.size();
    max_iter = 0;
    cur_max_K = 40;
    cur_sw_smem = 0;
    for (auto& x : m) { if (x.flags & 1) max_iter = std::max<int>(max_iter, x.max_iter); cur_max_K = std::max<int>(cur_max_K, x.K); cur_sw_smem = std::max(cur_sw_smem, sw_smem_bytes(x.K)); }
    CU(cudaMemcpyAsync(d_meta, staged ? staged : (const void*)h_meta.data(), sizeof(CbMeta) * n, cudaMemcpyHostToDevice, st));
    return 0;
  }
  // enqueue the whole 16-bit decode of blocks [lo, lo+cnt) (cnt < 0: all); returns #kernels launched or <0.
  // `part` selects the batch-maximum cell, so that parts of one batch can run as independent pipeline stages.
  // fe_rm != nullptr: block i (of the whole batch) is rm block i and k_demux16 reads its input out of the dematched
  // circular buffers (fused sub-block deinterleaving) instead of in_dev
  // in8: in_dev points at int8 soft bits (one byte per element, same element offsets)
  int decode16(const int16_t* in_dev, uint8_t* out_dev, uint8_t* status_dev, cudaStream_t st, int lo = 0, int cnt = -1,
               int part = 0, const RmBlock* fe_rm = nullptr, const int16_t* fe_w = nullptr, const int16_t* fe_harq = nullptr,
               int in8 = 0) {
    const int n = (cnt < 0) ? this->n : cnt;
    if (n <= 0) return 0;
    // Small batches are launch-latency bound (one K=40 block: 31 launches, 0.24 ms): their launch sequence -- static for a
    // given block count, iteration limit and set of pointers, early exits are decided on the device -- is built once
    // as a CUDA graph and replayed.
    if (n <= GRAPH_MAX_BLOCKS && !prof.on && g_use_graphs) {
      const GraphKey key{in_dev, out_dev, status_dev, fe_rm, fe_w, fe_harq, lo, n, part, max_iter, max_K, ctx->gen, in8, cur_max_K, sw_mode ? cur_sw_smem : 0};
      GraphEntry* ge = nullptr;
      for (auto& g : graphs) if (g.key == key) { ge = &g; break; }
      if (!ge) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        int l = -1;
        if (cudaGraphCreate(&graph, 0) == cudaSuccess) {
          Launcher rec;
          rec.graph = graph;
          l = enqueue16(in_dev, out_dev, status_dev, rec, lo, n, part, fe_rm, fe_w, fe_harq, false, in8);
          if (!rec.ok || l < 0 || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) exec = nullptr;
          cudaGraphDestroy(graph);
        }
        cudaGetLastError();
        if (exec) {
          if (graphs.size() >= 16) { cudaGraphExecDestroy(graphs.front().exec); graphs.erase(graphs.begin()); }
          graphs.push_back(GraphEntry{key, exec, l});
          ge = &graphs.back();
        }
      }
      if (ge) {
        if (cudaGraphLaunch(ge->exec, st) != cudaSuccess) return fail(-101, "graph launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        g_launches += ge->launches;
        return ge->launches;
      }
    }
    Launcher direct;
    direct.st = st;
    return enqueue16(in_dev, out_dev, status_dev, direct, lo, n, part, fe_rm, fe_w, fe_harq, true, in8);
  }
  int enqueue16(const int16_t* in_dev, uint8_t* out_dev, uint8_t* status_dev, Launcher& L, int lo, int n, int part,
                const RmBlock* fe_rm, const int16_t* fe_w, const int16_t* fe_harq, bool count, int in8 = 0) {
    cudaStream_t st = L.st;                             // profiler events (never recorded into a graph: prof.on excludes graphs)
    int launches = 0;
    CbMeta* d_meta = this->d_meta + lo;
    CbState* d_state = this->d_state + lo;
    int16_t* d_ws = this->d_ws + (long)lo * slot_hw;
    u32* d_ckpt = this->d_ckpt + (long)lo * ckpt_words;
    int* d_batch_max = this->d_batch_max + part;
    int* act[2] = {this->d_active + 2 * lo, this->d_active + 2 * cap + 2 * lo};      // packed lists of running blocks
    int* nact[2] = {this->d_nactive + 8 * part, this->d_nactive + 8 * part + 4};
    int cur = 0;
    if (status_dev) status_dev += lo;
    if (sw_mode) {
      // optional sliding-window mode: the whole decode of a block group is one warp of one launch (td16_sw.cuh)
      if (fe_rm || in8) return fail(-3, "sliding-window mode takes the decoder input y as int16");
      SwArgs s{};
      s.meta = d_meta; s.state = d_state; s.nblk = n; s.in_base = in_dev; s.out_base = out_dev; s.status_out = status_dev;
      s.tab_pool = ctx->sw_tab; s.tab_off = ctx->sw_tab_off; s.crc_xp = ctx->crc_xp;
      prof.begin(1, st);
      L.run(k_turbo_sw, dim3(n), dim3(32), (size_t)cur_sw_smem, s);
      prof.end(st);
      if (count) g_launches += 1;
      if (!L.ok) return -101;
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return fail(-101, "kernel launch failed: %s", cudaGetErrorString(e));
      return 1;
    }
    XchgArgs x{};
    x.meta = d_meta; x.state = d_state; x.ws = d_ws; x.slot_hw = slot_hw; x.A = A; x.nblk = n;
    x.pi_pool = ctx->pi_pool; x.t_pool = ctx->t_pool; x.crc_xp = ctx->crc_xp; x.in_base = in_dev; x.out_base = out_dev;
    x.status_out = status_dev; x.iter = 0; x.guard_b = GUARD_B; x.batch_max = d_batch_max;
    x.active = nullptr; x.nactive = nullptr; x.nactive_next = nullptr;                           // k_demux16 sees all blocks
    x.rm = fe_rm ? fe_rm + lo : nullptr; x.w_pool = fe_w; x.harq_pool = fe_harq; x.in8 = in8;
    L.zero_ints(nact[0], 8);                              // list counters of both lists + the retry flag of this part
    L.zero_ints(d_batch_max, 1);
    MapArgs mp{};
    mp.meta = d_meta; mp.state = d_state; mp.ws = d_ws; mp.slot_hw = slot_hw; mp.A = A;
    mp.ckpt = d_ckpt; mp.ckpt_words = ckpt_words; mp.nblk = n; mp.guard_b = GUARD_B; mp.batch_max = d_batch_max;
    // packs the running blocks into list c (its counter is zero: memset above / k_x1_16) and makes it current
    auto compact_into = [&](int c) {
      L.run(k_compact, dim3((n + COMPACT_THREADS - 1) / COMPACT_THREADS), dim3(COMPACT_THREADS), 0, (const CbState*)d_state, n, act[c], nact[c], (int)GUARD_B, (int)g_track);
      ++launches;
      mp.active = act[c]; mp.nactive = nact[c]; x.active = act[c]; x.nactive = nact[c]; x.nactive_next = nact[1 - c];
    };
    int map_seq = 0;
    const int map_grid = ((n + 14) * 4 + MAP_THREADS - 1) / MAP_THREADS;    // + the padding between the three classes of the list
    const size_t map_smem = MAP_SMEM_BYTES;
    auto map = [&](int sys_arr, int par_arr, int out_arr, int term, int iter, int upd) {
      mp.sys_arr = sys_arr; mp.par_arr = par_arr; mp.out_arr = out_arr; mp.term = term; mp.iter = iter; mp.upd = upd;
      mp.track = g_track; mp.retry = 0; mp.force = 0; mp.retry_flag = nact[1] + 3; mp.seq = ++map_seq;
      prof.begin(1, st);
      L.run(k_map16<MAP_SEG>, dim3(map_grid), dim3(MAP_THREADS), map_smem, mp);
      if (g_track) {
        // the retry launch: blocks whose tracked pass failed its range certificate repeat the pass on the exact policy
        // (every other warp returns at once: a few microseconds when nothing failed)
        mp.retry = 1;
        L.run(k_map16<MAP_SEG>, dim3(map_grid), dim3(MAP_THREADS), map_smem, mp);
        mp.retry = 0;
        ++launches;
      }
      prof.end(st);
      ++launches;
    };
    prof.begin(0, st);
    const int xth = xchg_threads_for(cur_max_K);
    auto launch_x1 = [&]() {
      if (xth == 64) L.run(k_x1_16<64>, dim3(n), dim3(64), A * sizeof(int16_t), x);
      else if (xth == 128) L.run(k_x1_16<128>, dim3(n), dim3(128), A * sizeof(int16_t), x);
      else L.run(k_x1_16<256>, dim3(n), dim3(256), A * sizeof(int16_t), x);
    };
    auto launch_x2 = [&]() {
      if (xth == 64) L.run(k_x2_16<64>, dim3(n), dim3(64), A * sizeof(int16_t), x);
      else if (xth == 128) L.run(k_x2_16<128>, dim3(n), dim3(128), A * sizeof(int16_t), x);
      else L.run(k_x2_16<256>, dim3(n), dim3(256), A * sizeof(int16_t), x);
    };
    if (fe_rm) L.run(k_demux16_t<true>, dim3(n), dim3(XCHG_THREADS), (3 * A + 3 * 32 * ((max_K + 4 + 31) / 32)) * sizeof(int16_t), x);
    else if (xth == 64) L.run(k_demux16_t<false, 64>, dim3(n), dim3(64), 3 * A * sizeof(int16_t), x);
    else if (xth == 128) L.run(k_demux16_t<false, 128>, dim3(n), dim3(128), 3 * A * sizeof(int16_t), x);
    else L.run(k_demux16_t<false>, dim3(n), dim3(XCHG_THREADS), 3 * A * sizeof(int16_t), x);
    prof.end(st);
    ++launches;
    compact_into(cur);
    map(ARR_S0, ARR_P1, ARR_EXT, 0, 1, 0);                       // reference :1199
    for (int it = 1; it <= max_iter; ++it) {                    // reference :1201
      x.iter = it;
      prof.begin(2, st);
      launch_x1();
      prof.end(st);
      map(ARR_SYS, ARR_P2, ARR_EXT2, 1, it, 0);                  // :1236
      prof.begin(3, st);
      launch_x2();
      prof.end(st);
      launches += 2;
      if (it < max_iter) {
        if (it > 1) { cur = 1 - cur; compact_into(cur); }        // the CRC check (iterations >= 2) may have retired blocks
        map(ARR_SYS, ARR_P1, ARR_EXT, 0, it + 1, 1);             // :1354-1375 (feedback fused)
      }
    }
    if (count) g_launches += launches;
    if (!L.ok) return -101;                               // graph construction failed: the caller launches directly instead
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-101, "kernel launch failed: %s", cudaGetErrorString(e));
    return launches;
  }
};


// ---- 8-bit decoder batch (TD8): int8 per-position arrays + full alpha/beta arrays in HBM ----
struct Batch8 {
  Profiler prof;
  DevCtx* ctx = nullptr;
  int cap = 0, n = 0, A = 0, max_iter = 0;
  long slot_b = 0, ck_words = 0;
  CbMeta* d_meta = nullptr;
  CbState* d_state = nullptr;
  int8_t* d_ws = nullptr;
  u32* d_ck = nullptr;
  std::vector<CbMeta> h_meta;
  int alloc(DevCtx* c, int ncb, int Kmax) {
    ctx = c; cap = ncb;
    A = (Kmax + 127) & ~127;              // C8 layout: whole 8-step chunks of 128 bytes
    slot_b = (long)A8_COUNT * A;
    ck_words = ckpt8_words(Kmax / 16);
    CU(cudaMalloc(&d_meta, sizeof(CbMeta) * ncb));
    CU(cudaMalloc(&d_state, sizeof(CbState) * ncb));
    CU(cudaMalloc(&d_ws, slot_b * ncb));
    CU(cudaMalloc(&d_ck, sizeof(u32) * ck_words * ncb));
    CU(cudaMemset(d_ws, 0, slot_b * ncb));
    CU(cudaMemset(d_state, 0, sizeof(CbState) * ncb));
    CU(cudaDeviceSynchronize());                       // see Batch::alloc
    return 0;
  }
  void release() {
    drop_graphs();
    cudaFree(d_meta); cudaFree(d_state); cudaFree(d_ws); cudaFree(d_ck);
    d_meta = nullptr; d_state = nullptr; d_ws = nullptr; d_ck = nullptr; cap = 0;
  }
  int set_meta(const std::vector<CbMeta>& m, cudaStream_t st, const void* staged = nullptr) {
    h_meta = m;
    n = (int)m.size();
    max_iter = 0;
    for (auto& x : m) if (x.flags & 1) max_iter = std::max<int>(max_iter, x.max_iter);
    CU(cudaMemcpyAsync(d_meta, staged ? staged : (const void*)h_meta.data(), sizeof(CbMeta) * n, cudaMemcpyHostToDevice, st));
    return 0;
  }
  std::vector<Batch::GraphEntry> graphs;                      // small batches: cached launch graphs, see Batch::decode16
  void drop_graphs() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
    graphs.clear();
  }
  int decode8(const int16_t* in_dev, uint8_t* out_dev, uint8_t* status_dev, cudaStream_t st) {
    if (n <= 0) return 0;
    if (n <= Batch::GRAPH_MAX_BLOCKS && !prof.on && g_use_graphs) {
      const Batch::GraphKey key{in_dev, out_dev, status_dev, nullptr, nullptr, nullptr, 0, n, 0, max_iter, 0, ctx->gen};
      Batch::GraphEntry* ge = nullptr;
      for (auto& g : graphs) if (g.key == key) { ge = &g; break; }
      if (!ge) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        int l = -1;
        if (cudaGraphCreate(&graph, 0) == cudaSuccess) {
          Launcher rec;
          rec.graph = graph;
          l = enqueue8(in_dev, out_dev, status_dev, rec, false);
          if (!rec.ok || l < 0 || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) exec = nullptr;
          cudaGraphDestroy(graph);
        }
        cudaGetLastError();
        if (exec) {
          if (graphs.size() >= 16) { cudaGraphExecDestroy(graphs.front().exec); graphs.erase(graphs.begin()); }
          graphs.push_back(Batch::GraphEntry{key, exec, l});
          ge = &graphs.back();
        }
      }
      if (ge) {
        if (cudaGraphLaunch(ge->exec, st) != cudaSuccess) return fail(-101, "graph launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        g_launches += ge->launches;
        return ge->launches;
      }
    }
    Launcher direct;
    direct.st = st;
    return enqueue8(in_dev, out_dev, status_dev, direct, true);
  }
  int enqueue8(const int16_t* in_dev, uint8_t* out_dev, uint8_t* status_dev, Launcher& L, bool count) {
    cudaStream_t st = L.st;
    int launches = 0;
    Td8Args a{};
    a.meta = d_meta; a.state = d_state; a.ws = d_ws; a.slot_b = slot_b; a.A = A; a.ck = d_ck; a.ck_words = ck_words;
    a.nblk = n; a.qpp = ctx->qpp_pool; a.t8 = ctx->t8_pool; a.crc_xp = ctx->crc_xp; a.in_base = in_dev; a.out_base = out_dev;
    a.status_out = status_dev; a.iter = 0; a.sys_arr = a.par_arr = a.out_arr = 0;
    const int map_grid = (n * 8 + MAP8_THREADS - 1) / MAP8_THREADS;
    auto map = [&](int sys_arr, int par_arr, int out_arr, int iter) {
      a.sys_arr = sys_arr; a.par_arr = par_arr; a.out_arr = out_arr; a.iter = iter;
      prof.begin(1, st);
      L.run(k_map8, dim3(map_grid), dim3(MAP8_THREADS), MAP8_SMEM_BYTES, a);
      prof.end(st);
      ++launches;
    };
    prof.begin(0, st);
    L.run(k_demux8, dim3(n), dim3(XCHG_THREADS), 3 * A, a);
    prof.end(st);
    ++launches;
    map(A8_S0, A8_P1, A8_EXT, 1);                                // TD8:1325
    for (int it = 1; it <= max_iter; ++it) {                    // TD8:1327
      a.iter = it;
      prof.begin(2, st);
      L.run(k_x1_8, dim3(n), dim3(XCHG_THREADS), A, a);
      prof.end(st);
      map(A8_SYS, A8_P2, A8_EXT2, it);                           // TD8:1386
      prof.begin(3, st);
      L.run(k_x2_8, dim3(n), dim3(XCHG_THREADS), 2 * A, a);
      prof.end(st);
      launches += 2;
      if (it < max_iter) map(A8_SYS, A8_P1, A8_EXT, it + 1);     // TD8:1634
    }
    if (count) g_launches += launches;
    if (!L.ok) return -101;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-101, "kernel launch failed: %s", cudaGetErrorString(e));
    return launches;
  }
};

static bool td8_domain(int K) { return K >= 256 && (K & 15) == 0; }

static int make_meta(DevCtx* c, int K, int max_it, int crc, int F, int dec, long in_off, long out_off, CbMeta* m,
                     bool llr8 = false) {
  int idx = qpp_index(K);
  if (idx < 0) return -1;
  m->K = (uint16_t)K; m->W = (uint16_t)(K >> 3);
  m->max_iter = (uint8_t)max_it; m->crc_type = (uint8_t)crc; m->F = (uint8_t)F; m->flags = dec ? 1 : 0;
  m->pi_off = llr8 ? c->qpp_off[idx] : c->pi_off[idx];
  m->t_off = llr8 ? c->t8_off[idx] : c->t_off[idx];
  m->in_off_lo = (uint32_t)((unsigned long long)in_off & 0xffffffffu);
  m->in_off_hi = (uint32_t)((unsigned long long)in_off >> 32);
  m->out_off = (uint32_t)out_off;
  return 0;
}



// ---- rate-matching parameters (lte_rate_matching.c:719-735) ---------------------------------
struct RmParams { uint32_t RTC, Kpi, ND, Ncb, k0, E; };
static int rm_params(uint32_t K, uint32_t G, uint8_t C, uint32_t Nsoft, uint8_t Mdlharq, uint8_t Kmimo, uint8_t rvidx,
                     uint8_t Qm, uint8_t Nl, uint8_t r, uint32_t RTC_in, RmParams* o) {
  if (Kmimo == 0 || Mdlharq == 0 || C == 0 || Qm == 0 || Nl == 0) return -1;       // :713-717
  const uint32_t D = K + 4;
  o->RTC = RTC_in ? RTC_in : (D >> 5) + ((D & 31) ? 1 : 0);
  o->Kpi = o->RTC << 5;
  o->ND = o->Kpi - D;
  const uint32_t Nir = Nsoft / Kmimo / (Mdlharq < 8 ? Mdlharq : 8);
  o->Ncb = std::min<uint32_t>(Nir / C, 3 * o->Kpi);
  const uint32_t Gp = G / Nl / Qm, GpmodC = Gp % C;
  o->E = (r < (uint32_t)(C - GpmodC)) ? Nl * Qm * (Gp / C) : Nl * Qm * ((GpmodC == 0 ? 0 : 1) + (Gp / C));
  const uint32_t Ncbmod = o->Ncb % (o->RTC << 3);
  o->k0 = o->RTC * (2 + (rvidx * (((Ncbmod == 0) ? 0 : 1) + (o->Ncb / (o->RTC << 3))) * 2));
  return 0;
}

}  // namespace oai
// device-resident HARQ soft buffers (include/oai_turbo_b200.h section 2)
struct oai_turbo_harq_pool {
  int dev = -1;
  uint32_t n_slots = 0, slot_hw = 0;     // slot size in int16: 3 * Kpi(max_K)
  int16_t* d = nullptr;
};
namespace oai {

// per-thread, per-device scratch for the single-call reference entry points of the front end: the stream and both
// buffers belong to the device that was current when they were created, so a thread that drives several GPUs gets
// one set per device; everything is released when the thread exits
struct Scratch {
  cudaStream_t st = nullptr;
  void* h = nullptr; void* d = nullptr; size_t cap = 0, dcap = 0;
  // host_bytes: size of the page-locked mirror (defaults to the device size; smaller when part of the device area is
  // never copied, e.g. the intermediate d of the TX batch).  The current device must be the scratch's device.
  int ensure(size_t bytes, size_t host_bytes = (size_t)-1) {
    DevCtx* c;
    int rc = ctx_get(-1, &c);
    if (rc) return rc;
    if (!st) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    if (host_bytes == (size_t)-1) host_bytes = bytes;
    if (host_bytes > cap) {
      if (h) cudaFreeHost(h);
      h = nullptr; cap = 0;
      const size_t want = host_bytes + (host_bytes >> 2) + 4096;
      CU(cudaMallocHost(&h, want));
      cap = want;
    }
    if (bytes > dcap) {
      if (d) cudaFree(d);
      d = nullptr; dcap = 0;
      const size_t want = bytes + (bytes >> 2) + 4096;
      CU(cudaMalloc(&d, want));
      dcap = want;
    }
    return 0;
  }
  void release() {
    if (h) cudaFreeHost(h);
    if (d) cudaFree(d);
    if (st) cudaStreamDestroy(st);
    h = d = nullptr; st = nullptr; cap = dcap = 0;
  }
};
struct ScratchSet {
  Scratch per_dev[16];
  ~ScratchSet() { for (Scratch& s : per_dev) s.release(); cudaGetLastError(); }
};
static thread_local ScratchSet t_scratch_set;
// scratch of the calling thread on the CURRENT device
static Scratch& scratch_here() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) dev = 0;
  return t_scratch_set.per_dev[dev];
}

// Page-locked staging for the metadata arrays that are copied LATE in submit(): the uplink front-end list (behind the
// allocations' soft-bit copy) and the transport-block lists (behind the whole decode).  cudaMemcpyAsync from PAGEABLE
// memory blocks the host until the stream reaches the copy, which made submit() synchronous for those batches (measured:
// 12 ms of a 17 ms submit for 8 192 uplink allocations, and no overlap between two batches in flight).  From this arena
// the copies are asynchronous.
// The block / rate-matching lists copied at the START of submit() deliberately stay pageable: that copy returns when the
// copy engine has finished the PREVIOUS batch's inputs, i.e. it paces a caller that keeps two batches in flight.  Making
// it asynchronous as well was measured and is worse (e feed, two in flight, same box: 13.8 -> 9.5 Gbit/s int16, 19.0 ->
// 16.7 int8: with everything of batch i+1 enqueued while batch i is still copying, the two batches' parts interleave on
// the GPU and both finish late), also with an explicit event wait in its place (13.4 / 16.8).
struct PinnedArena {
  char* p = nullptr;
  size_t cap = 0, used = 0;
  int reset(size_t need) {                   // call while no copy of this batch object is in flight
    used = 0;
    if (need > cap) {
      if (p) cudaFreeHost(p);
      p = nullptr; cap = 0;
      const size_t want = need + (need >> 2) + 4096;
      CU(cudaMallocHost(&p, want));
      cap = want;
    }
    return 0;
  }
  const void* put(const void* src, size_t n) {
    used = (used + 15) & ~(size_t)15;
    if (used + n > cap || getenv("OAI_TURBO_NO_ARENA")) return src;          // (cannot happen with the bound computed in submit(); pageable fallback)
    void* dst = p + used;
    memcpy(dst, src, n);
    used += n;
    return dst;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = used = 0; }
};

// ---- host-buffer batches: pinned staging + one stream per batch object ---------------
struct HostBatch {
  PinnedArena arena;
  Batch b;
  Batch8 b8;                       // 8-bit decoder blocks of the same submit (placed after the 16-bit ones)
  int cap8_blocks = 0, cap8_K = 0, n16 = 0;
  cudaStream_t st = nullptr;
  cudaStream_t st_copy = nullptr, st_out = nullptr;      // input / output copies of the pipelined form
  cudaStream_t st_part[PART_STREAMS] = {};               // compute streams of its parts
  cudaEvent_t ev_part[MAX_PARTS] = {}, ev_done[MAX_PARTS] = {}, ev_out = nullptr, ev_setup = nullptr;
  int8_t* d_in8 = nullptr; int8_t* h_in8 = nullptr; size_t cap_in8 = 0;   // narrow input feed: packed soft bits (pinned stage + device)
  int ensure_in8(size_t bytes) {
    if (bytes > cap_in8) {
      if (d_in8) cudaFree(d_in8);
      if (h_in8) cudaFreeHost(h_in8);
      d_in8 = nullptr; h_in8 = nullptr; cap_in8 = 0;
      CU(cudaMalloc(&d_in8, bytes));
      CU(cudaMallocHost(&h_in8, bytes));
      cap_in8 = bytes;
    }
    return 0;
  }
  bool direct_out = false;
  int dev = -1;
  int cap_blocks = 0, cap_K = 0;
  size_t cap_in = 0, cap_out = 0, cap_h_in = 0, cap_h_out = 0;   // device buffers / pinned staging (staging only on demand)
  int16_t* h_in = nullptr;  int16_t* d_in = nullptr;
  uint8_t* h_out = nullptr; uint8_t* d_out = nullptr;
  uint8_t* h_status = nullptr; uint8_t* d_status = nullptr; int cap_status = 0;
  // fused front end (dematch + deinterleave): soft-bit pool, HARQ w pool, per-block parameters
  size_t cap_e = 0, cap_w = 0; int cap_rm = 0;
  int16_t* h_e = nullptr; int16_t* d_e = nullptr;
  int16_t* h_w = nullptr; int16_t* d_w = nullptr;
  RmBlock* d_rm = nullptr;
  GoldSeq* d_gseq = nullptr; uint32_t* d_gold = nullptr; int cap_gseq = 0; size_t cap_gold = 0;   // scrambling sequences
  std::vector<GoldSeq> gseq;
  std::vector<RmBlock> rm;
  std::vector<int> rm_desc;        // descriptor index of each RmBlock
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // OAI_TURBO_TRACE: submit, H2D done, kernels done, D2H done
  // transport-block outputs (oai_turbo_submit_tbs): reassembled b + return value per transport block, built on the device
  std::vector<oai_tb_desc_t> tbs;
  std::vector<TbMeta> tb_meta;
  std::vector<TbBlk> tb_blk;
  TbMeta* d_tbmeta = nullptr; TbBlk* d_tbblk = nullptr; TbResult* d_tbres = nullptr; TbResult* h_tbres = nullptr;
  uint8_t* d_tbpool = nullptr; uint8_t* h_tbpool = nullptr;
  int cap_tb = 0, cap_tbblk = 0; size_t cap_tbpool = 0;
  bool want_cb_out = true;         // some descriptor has decoded_bytes != NULL
  // uplink front end (oai_ul_front_t): soft bits of the allocations -> e[] on the device
  std::vector<UlFrontDev> ulf;
  std::vector<int> ulf_tb;         // transport-block index of each entry
  std::vector<oai_ul_front_t> ulf_desc;
  UlFrontDev* d_ulf = nullptr; UlFrontOut* d_ulfout = nullptr; UlFrontOut* h_ulfout = nullptr;
  uint8_t* d_llr = nullptr; uint8_t* h_llr = nullptr; int8_t* d_cqi = nullptr; int8_t* h_cqi = nullptr;
  int cap_ulf = 0; size_t cap_llr = 0, cap_h_llr = 0, cap_cqi = 0, cqi_bytes = 0;
  int ensure_ulf(int n, size_t llr_bytes, size_t cqi_b) {
    if (n > cap_ulf) {
      if (d_ulf) cudaFree(d_ulf);
      if (d_ulfout) cudaFree(d_ulfout);
      if (h_ulfout) cudaFreeHost(h_ulfout);
      d_ulf = nullptr; d_ulfout = nullptr; h_ulfout = nullptr; cap_ulf = 0;
      CU(cudaMalloc(&d_ulf, sizeof(UlFrontDev) * n));
      CU(cudaMalloc(&d_ulfout, sizeof(UlFrontOut) * n));
      CU(cudaMallocHost(&h_ulfout, sizeof(UlFrontOut) * n));
      cap_ulf = n;
    }
    if (llr_bytes > cap_llr) {
      if (d_llr) cudaFree(d_llr);
      d_llr = nullptr; cap_llr = 0;
      CU(cudaMalloc(&d_llr, llr_bytes));
      cap_llr = llr_bytes;
    }
    if (cqi_b > cap_cqi) {
      if (d_cqi) cudaFree(d_cqi);
      if (h_cqi) cudaFreeHost(h_cqi);
      d_cqi = nullptr; h_cqi = nullptr; cap_cqi = 0;
      CU(cudaMalloc(&d_cqi, cqi_b));
      CU(cudaMallocHost(&h_cqi, cqi_b));
      cap_cqi = cqi_b;
    }
    return 0;
  }
  int ensure_stage_llr() {
    if (cap_h_llr < cap_llr) {
      if (h_llr) cudaFreeHost(h_llr);
      h_llr = nullptr; cap_h_llr = 0;
      CU(cudaMallocHost(&h_llr, cap_llr));
      cap_h_llr = cap_llr;
    }
    return 0;
  }
  int ensure_tb(int ntb, int nblk, size_t pool_bytes) {
    if (ntb > cap_tb) {
      if (d_tbmeta) cudaFree(d_tbmeta);
      if (d_tbres) cudaFree(d_tbres);
      if (h_tbres) cudaFreeHost(h_tbres);
      d_tbmeta = nullptr; d_tbres = nullptr; h_tbres = nullptr; cap_tb = 0;
      CU(cudaMalloc(&d_tbmeta, sizeof(TbMeta) * ntb));
      CU(cudaMalloc(&d_tbres, sizeof(TbResult) * ntb));
      CU(cudaMallocHost(&h_tbres, sizeof(TbResult) * ntb));
      cap_tb = ntb;
    }
    if (nblk > cap_tbblk) {
      if (d_tbblk) cudaFree(d_tbblk);
      d_tbblk = nullptr; cap_tbblk = 0;
      CU(cudaMalloc(&d_tbblk, sizeof(TbBlk) * nblk));
      cap_tbblk = nblk;
    }
    if (pool_bytes > cap_tbpool) {
      if (d_tbpool) cudaFree(d_tbpool);
      if (h_tbpool) cudaFreeHost(h_tbpool);
      d_tbpool = nullptr; h_tbpool = nullptr; cap_tbpool = 0;
      CU(cudaMalloc(&d_tbpool, pool_bytes));
      CU(cudaMallocHost(&h_tbpool, pool_bytes));
      cap_tbpool = pool_bytes;
    }
    return 0;
  }
  // bookkeeping of the submitted batch
  std::vector<oai_cb_desc_t> descs;
  std::vector<int> order;          // GPU block i <-> descriptor order[i]
  std::vector<uint32_t> out_off;
  unsigned flags = 0;

  // The caller holds a DevGuard on the batch's device.  Capacities are raised only after the matching allocation
  // succeeded; a failed allocation releases the whole object (nothing half-allocated survives into the next call).
  int ensure(DevCtx* c, int nblk, int Kmax, size_t in_hw, size_t out_bytes, int nblk8 = 0, int Kmax8 = 0) {
    if (dev != c->dev) { release(); dev = c->dev; }
    int rc = ensure_inner(c, nblk, Kmax, in_hw, out_bytes, nblk8, Kmax8);
    if (rc) { release(); cudaGetLastError(); }
    return rc;
  }
  int ensure_inner(DevCtx* c, int nblk, int Kmax, size_t in_hw, size_t out_bytes, int nblk8, int Kmax8) {
    int rc;
    if (!st) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    if (nblk > cap_blocks || Kmax > cap_K) {
      b.release();
      const int nb = std::max(nblk, cap_blocks), nk = std::max(Kmax, cap_K);
      cap_blocks = cap_K = 0;
      rc = b.alloc(c, nb, nk);
      if (rc) return rc;
      cap_blocks = nb; cap_K = nk;
    }
    if (nblk8 > cap8_blocks || (nblk8 > 0 && Kmax8 > cap8_K)) {
      b8.release();
      const int nb = std::max(nblk8, cap8_blocks), nk = std::max(Kmax8, cap8_K);
      cap8_blocks = cap8_K = 0;
      rc = b8.alloc(c, nb, nk);
      if (rc) return rc;
      cap8_blocks = nb; cap8_K = nk;
    }
    if (nblk + nblk8 > cap_status) {
      if (h_status) cudaFreeHost(h_status);
      if (d_status) cudaFree(d_status);
      h_status = nullptr; d_status = nullptr; cap_status = 0;
      CU(cudaMallocHost(&h_status, nblk + nblk8));
      CU(cudaMalloc(&d_status, nblk + nblk8));
      cap_status = nblk + nblk8;
    }
    b.ctx = c; b8.ctx = c;
    if (in_hw > cap_in) {
      if (d_in) cudaFree(d_in);
      d_in = nullptr; cap_in = 0;
      CU(cudaMalloc(&d_in, in_hw * sizeof(int16_t)));
      cap_in = in_hw;
    }
    if (out_bytes > cap_out) {
      if (d_out) cudaFree(d_out);
      d_out = nullptr; cap_out = 0;
      CU(cudaMalloc(&d_out, out_bytes));
      cap_out = out_bytes;
    }
    return 0;
  }
  // pinned staging for pageable caller memory; page-locked caller memory never needs it
  int ensure_stage_in() {
    if (cap_h_in < cap_in) {
      if (h_in) cudaFreeHost(h_in);
      h_in = nullptr; cap_h_in = 0;
      CU(cudaMallocHost(&h_in, cap_in * sizeof(int16_t)));
      cap_h_in = cap_in;
    }
    return 0;
  }
  int ensure_stage_out() {
    if (cap_h_out < cap_out) {
      if (h_out) cudaFreeHost(h_out);
      h_out = nullptr; cap_h_out = 0;
      CU(cudaMallocHost(&h_out, cap_out));
      cap_h_out = cap_out;
    }
    return 0;
  }
  int ensure_rm(size_t e_hw, size_t w_hw, int nrm) {
    if (e_hw > cap_e) {
      if (h_e) cudaFreeHost(h_e);
      if (d_e) cudaFree(d_e);
      h_e = nullptr; d_e = nullptr; cap_e = 0;
      CU(cudaMallocHost(&h_e, e_hw * sizeof(int16_t)));
      CU(cudaMalloc(&d_e, e_hw * sizeof(int16_t)));
      cap_e = e_hw;
    }
    if (w_hw > cap_w) {
      if (h_w) cudaFreeHost(h_w);
      if (d_w) cudaFree(d_w);
      h_w = nullptr; d_w = nullptr; cap_w = 0;
      CU(cudaMallocHost(&h_w, w_hw * sizeof(int16_t)));
      CU(cudaMalloc(&d_w, w_hw * sizeof(int16_t)));
      cap_w = w_hw;
    }
    if (nrm > cap_rm) {
      if (d_rm) cudaFree(d_rm);
      d_rm = nullptr; cap_rm = 0;
      CU(cudaMalloc(&d_rm, sizeof(RmBlock) * nrm));
      cap_rm = nrm;
    }
    return 0;
  }
  void release() {
    arena.release();
    b.release();
    b8.release(); cap8_blocks = cap8_K = 0; cap_status = 0;
    if (h_e) cudaFreeHost(h_e);
    if (d_e) cudaFree(d_e);
    if (h_w) cudaFreeHost(h_w);
    if (d_w) cudaFree(d_w);
    h_e = nullptr; d_e = nullptr; h_w = nullptr; d_w = nullptr;
    if (d_rm) { cudaFree(d_rm); d_rm = nullptr; }
    if (d_gseq) { cudaFree(d_gseq); d_gseq = nullptr; }
    if (d_gold) { cudaFree(d_gold); d_gold = nullptr; }
    cap_gseq = 0; cap_gold = 0;
    if (d_tbmeta) cudaFree(d_tbmeta);
    if (d_tbres) cudaFree(d_tbres);
    if (h_tbres) cudaFreeHost(h_tbres);
    if (d_tbblk) cudaFree(d_tbblk);
    if (d_tbpool) cudaFree(d_tbpool);
    if (h_tbpool) cudaFreeHost(h_tbpool);
    d_tbmeta = nullptr; d_tbres = nullptr; h_tbres = nullptr; d_tbblk = nullptr; d_tbpool = nullptr; h_tbpool = nullptr;
    cap_tb = cap_tbblk = 0; cap_tbpool = 0;
    if (d_ulf) cudaFree(d_ulf);
    if (d_ulfout) cudaFree(d_ulfout);
    if (h_ulfout) cudaFreeHost(h_ulfout);
    if (d_llr) cudaFree(d_llr);
    if (h_llr) cudaFreeHost(h_llr);
    if (d_cqi) cudaFree(d_cqi);
    if (h_cqi) cudaFreeHost(h_cqi);
    d_ulf = nullptr; d_ulfout = nullptr; h_ulfout = nullptr; d_llr = nullptr; h_llr = nullptr; d_cqi = nullptr; h_cqi = nullptr;
    cap_ulf = 0; cap_llr = cap_h_llr = cap_cqi = 0;
    cap_e = cap_w = 0; cap_rm = 0;
    if (h_in) { cudaFreeHost(h_in); h_in = nullptr; }
    if (d_in) { cudaFree(d_in); d_in = nullptr; }
    if (h_out) { cudaFreeHost(h_out); h_out = nullptr; }
    if (d_out) { cudaFree(d_out); d_out = nullptr; }
    cap_h_in = cap_h_out = 0;
    if (h_status) cudaFreeHost(h_status);
    if (d_status) cudaFree(d_status);
    h_status = nullptr; d_status = nullptr;
    for (auto& e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
    if (st) { cudaStreamDestroy(st); st = nullptr; }
    if (st_copy) {
      cudaStreamDestroy(st_copy); st_copy = nullptr; cudaStreamDestroy(st_out); st_out = nullptr;
      for (auto& s2 : st_part) { if (s2) cudaStreamDestroy(s2); s2 = nullptr; }
      for (auto& e : ev_part) { cudaEventDestroy(e); e = nullptr; }
      for (auto& e : ev_done) { cudaEventDestroy(e); e = nullptr; }
      cudaEventDestroy(ev_out); ev_out = nullptr;
      if (ev_setup) { cudaEventDestroy(ev_setup); ev_setup = nullptr; }
    }
    if (d_in8) cudaFree(d_in8);
    if (h_in8) cudaFreeHost(h_in8);
    d_in8 = nullptr; h_in8 = nullptr; cap_in8 = 0;
    cap_blocks = cap_K = 0; cap_in = cap_out = 0;
  }

  int submit(const oai_cb_desc_t* cbs, int ncb, unsigned fl, int gpu, const oai_tb_desc_t* tb_in = nullptr, int ntb = 0) {
    descs.assign(cbs, cbs + ncb);
    flags = fl;
    tbs.clear(); tb_meta.clear(); tb_blk.clear();
    for (int i = 0; i < ntb; ++i) {
      const oai_tb_desc_t& t = tb_in[i];
      if (t.C == 0 || t.C > (uint32_t)TB_MAX_C || (unsigned long long)t.first_cb + t.C > (unsigned long long)ncb)
        return fail(-4, "transport block %d: code blocks [%u, %u) are not inside the descriptor array / C > 16", i, t.first_cb, t.first_cb + t.C);
    }
    if (ntb > 0) tbs.assign(tb_in, tb_in + ntb);
    // uplink front ends: descriptor index -> (entry, running soft-bit offset inside the allocation's e[])
    ulf.clear(); ulf_tb.clear(); ulf_desc.clear();
    std::vector<int> ul_of(ncb, -1);
    std::vector<uint32_t> ul_roff(ncb, 0);
    size_t ul_e_hw = 0, ul_llr_b = 0, ul_cqi_b = 0;
    for (int i = 0; i < ntb; ++i) {
      const oai_tb_desc_t& t = tbs[i];
      if (!t.ul_front) continue;
      const oai_ul_front_t& f = *t.ul_front;
      if (!f.llr || (f.Qm != 2 && f.Qm != 4 && f.Qm != 6) || f.Cmux == 0 || f.O_ACK > 2 || f.O_RI > 1 || f.llr_fmt > 1 ||
          f.Hprime < f.Qprime_CQI || (f.Hprime + f.Qprime_RI) % f.Cmux)
        return fail(-4, "transport block %d: invalid uplink front-end parameters (ulsch_decoding returns -1 for O_ACK > 2 / O_RI > 1)", i);
      UlFrontDev u;
      memset(&u, 0, sizeof(u));
      const uint32_t Hpp = f.Hprime + f.Qprime_RI;
      u.llr_off_lo = (uint32_t)(ul_llr_b & 0xffffffffu); u.llr_off_hi = (uint32_t)((unsigned long long)ul_llr_b >> 32);
      u.llr_fmt = f.llr_fmt;
      u.e_off_lo = (uint32_t)(ul_e_hw & 0xffffffffu); u.e_off_hi = (uint32_t)((unsigned long long)ul_e_hw >> 32);
      u.Qm = f.Qm; u.Cmux = f.Cmux; u.Rp = Hpp / f.Cmux;
      u.Qprime_RI = f.Qprime_RI; u.Qprime_ACK = f.Qprime_ACK; u.Qprime_CQI = f.Qprime_CQI; u.Hprime = f.Hprime;
      u.Ncp = f.Ncp; u.O_ACK = f.O_ACK; u.O_RI = f.O_RI; u.bundling = f.bundling; u.Nbundled = f.Nbundled;
      u.cqi_off = (uint32_t)ul_cqi_b;
      const uint32_t G = (f.Hprime - f.Qprime_CQI) * f.Qm;          // data soft bits: the G of lte_rate_matching_turbo_rx
      uint32_t roff = 0;
      for (uint32_t r = 0; r < t.C; ++r) {
        const oai_cb_desc_t& d = descs[t.first_cb + r];
        RmParams q;
        if (!d.dematch_enable || d.G != G || rm_params(d.K, d.G, d.C, d.Nsoft, d.Mdlharq, d.Kmimo, d.rvidx, d.Qm, d.Nl, d.r, 0, &q))
          return fail(-4, "transport block %d, block %u: with ul_front every block needs dematch_enable and G = (Hprime - Qprime_CQI) * Qm = %u", i, r, G);
        ul_of[t.first_cb + r] = (int)ulf.size();
        ul_roff[t.first_cb + r] = roff;
        roff += q.E;
      }
      if (roff > G) return fail(-4, "transport block %d: its blocks consume %u soft bits, the allocation carries %u", i, roff, G);
      ul_e_hw += ((size_t)G + 7) & ~(size_t)7;
      ul_llr_b += (((size_t)Hpp * f.Qm * (f.llr_fmt ? 1 : 2)) + 15) & ~(size_t)15;
      ul_cqi_b += (((size_t)f.Qprime_CQI * f.Qm) + 15) & ~(size_t)15;
      ulf.push_back(u); ulf_tb.push_back(i); ulf_desc.push_back(f);
    }
    cqi_bytes = ul_cqi_b;
    static const bool trace = getenv("OAI_TURBO_TRACE") != nullptr;
    const auto t_sub0 = std::chrono::steady_clock::now();
    auto phase = [&](const char* what) {                        // OAI_TURBO_TRACE: host time spent in submit() so far
      if (trace) fprintf(stderr, "[trace %p] submit +%.2f ms: %s\n", (void*)this,
                         1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_sub0).count(), what);
    };
    // handles are recycled: nothing of the previous batch may survive an early return below (wait() walks these)
    order.clear(); rm.clear(); rm_desc.clear(); gseq.clear(); direct_out = false;
    int Kmax = 40;
    for (int i = 0; i < ncb; ++i) {
      const oai_cb_desc_t& d = descs[i];
      if (d.crc_type > 3 || qpp_index(d.K) < 0) { if (d.status) *d.status = 255; continue; }   // TD16:1003-1018
      if (d.llr8 && !td8_domain(d.K)) { if (d.status) *d.status = 255; continue; }    // outside TD8's parity domain
      // without the front end a block that is not decoded needs no GPU work at all; with it the
      // HARQ buffer is still combined (dlsch_decoding.c:333-385 runs before the err_flag test)
      if (!d.decode_enable && !d.dematch_enable) { if (d.status) *d.status = 0xFE; continue; }
      if (d.in_fmt > 1 || (d.in_fmt && !d.dematch_enable)) return fail(-4, "block %d: in_fmt %d is only defined for front-end blocks (int8 soft bits e)", i, (int)d.in_fmt);
      order.push_back(i);
      Kmax = std::max<int>(Kmax, d.K);
    }
    // 16-bit blocks first, then 8-bit ones; equal-K blocks next to each other (a warp of the MAP
    // kernel carries 8 blocks)
    auto before = [&](int a, int c) {
      if ((descs[a].llr8 != 0) != (descs[c].llr8 != 0)) return descs[a].llr8 == 0;
      return descs[a].K < descs[c].K;
    };
    if (!std::is_sorted(order.begin(), order.end(), before)) std::stable_sort(order.begin(), order.end(), before);
    const int n = (int)order.size();
    if (n == 0) {                                            // no block reaches the GPU: every transport block has failed
      for (auto& t : tbs) { if (t.ret) *t.ret = (uint8_t)(1 + descs[t.first_cb].max_iterations); if (t.valid_bytes) *t.valid_bytes = 0; }
      tbs.clear(); ulf.clear(); ulf_desc.clear();
      return 0;
    }
    n16 = 0;
    int Kmax8 = 0;
    for (int i = 0; i < n; ++i) {
      if (!descs[order[i]].llr8) ++n16;
      else Kmax8 = std::max<int>(Kmax8, descs[order[i]].K);
    }
    size_t in_hw = 0, out_b = 0;
    std::vector<size_t> in_off(n);
    out_off.resize(n);
    for (int i = 0; i < n; ++i) {
      const oai_cb_desc_t& d = descs[order[i]];
      in_off[i] = in_hw;  in_hw += (size_t)3 * d.K + 12;
      out_off[i] = (uint32_t)out_b; out_b += ((size_t)(d.K >> 3) + 15) & ~(size_t)15;
    }
    phase("descriptors checked and ordered");
    DevCtx* dctx;
    int rc = ctx_get(gpu, &dctx);
    if (rc) return rc;
    DevGuard guard;                                          // the caller's current device is restored on every return path
    if (guard.enter(dctx->dev)) return fail(-100, "cannot select CUDA device %d", dctx->dev);
    rc = ensure(dctx, std::max(n16, 1), Kmax, in_hw, out_b, n - n16, Kmax8);
    if (rc) return rc;
    // upper bound of everything that goes through the metadata arena (see PinnedArena)
    rc = arena.reset((size_t)n * (sizeof(CbMeta) + sizeof(RmBlock) + sizeof(GoldSeq) + 48) + (size_t)ncb * sizeof(TbBlk) +
                     (size_t)ntb * (sizeof(TbMeta) + sizeof(UlFrontDev) + sizeof(GoldSeq) + 64) + 4096);
    if (rc) return rc;
    if (trace) { for (auto& e : ev) if (!e) cudaEventCreate(&e); cudaEventRecord(ev[0], st); }
    // host->device: runs of blocks that are contiguous in the caller's memory go with one copy;
    // page-locked caller memory is copied from directly, pageable memory through the pinned stage
    auto enqueue_inputs = [&](int lo, int hi, cudaStream_t cs) -> int {
      for (int i = lo; i < hi;) {
        if (descs[order[i]].dematch_enable) { ++i; continue; }     // y is produced on the device by k_deint
        const int16_t* base = descs[order[i]].in;
        size_t len = (size_t)3 * descs[order[i]].K + 12;
        int j = i + 1;
        while (j < hi && !descs[order[j]].dematch_enable && descs[order[j]].in == base + len) { len += (size_t)3 * descs[order[j]].K + 12; ++j; }
        cudaPointerAttributes at;
        bool pinned = (cudaPointerGetAttributes(&at, base) == cudaSuccess) && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        const int16_t* src = base;
        if (!pinned) {
          int r2 = ensure_stage_in();
          if (r2) return r2;
          memcpy(h_in + in_off[i], base, len * sizeof(int16_t)); src = h_in + in_off[i];
        }
        CU(cudaMemcpyAsync(d_in + in_off[i], src, len * sizeof(int16_t), cudaMemcpyHostToDevice, cs));
        i = j;
      }
      return 0;
    };
    // Pipelined form (plain 16-bit batches that are large enough): the batch is cut into `parts` ranges of blocks; the
    // input copy of part i+1 (copy stream) overlaps the decode of part i (compute stream).  The last part is the
    // smallest one, because its decode is the only one that is not hidden behind a copy.
    int parts = 1;
    bool fe_parts = false;                                   // pipelined parts run the front end too
    {
      // eligible: only 16-bit blocks, and either no block uses the front end or all do with their HARQ buffers in a
      // device pool (a host-authoritative w would have to be staged in and out around every part)
      bool ok = (n == n16) && !getenv("OAI_TURBO_NO_PIPELINE");
      int n_fe = 0, n_pool = 0;
      for (int i = 0; i < n; ++i) { n_fe += descs[order[i]].dematch_enable ? 1 : 0; n_pool += (descs[order[i]].dematch_enable && descs[order[i]].harq_pool) ? 1 : 0; }
      ok = ok && (n_fe == 0 || (n_fe == n && n_pool == n));
      if (ok) parts = std::max(1, std::min(MAX_PARTS, n16 / (n_fe ? 2 * MIN_PART_BLOCKS : MIN_PART_BLOCKS)));
      fe_parts = parts > 1 && n_fe == n;
    }
    auto part_lo = [&](int part) -> int {                    // parts-1 equal ranges, then one of MIN_PART_BLOCKS
      if (part >= parts) return n;
      if (parts == 1) return 0;
      return (int)((long)(n - MIN_PART_BLOCKS) * part / (parts - 1));
    };
    if (parts > 1) {
      if (!st_copy) {
        CU(cudaStreamCreateWithFlags(&st_copy, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&st_out, cudaStreamNonBlocking));
        for (auto& s2 : st_part) CU(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ev_setup, cudaEventDisableTiming));
        for (auto& e : ev_part) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : ev_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming));
      }
    }
    std::vector<CbMeta> meta(n);
    size_t e_hw = ul_e_hw, w_hw = 0, e_bytes_run = 0;       // the allocations' e[] regions come first in the pool
    std::vector<char> rm_nocopy;                              // per rm block: soft bits come from k_ul_front
    oai_turbo_harq_pool* pool = nullptr;
    const char* prev_e_end = nullptr;
    std::unordered_map<uint32_t, int> seq_of;                // c_init -> index in gseq
    std::vector<int> rm_seq;                                 // per rm block: its sequence or -1
    for (int i = 0; i < n; ++i) {
      const oai_cb_desc_t& d = descs[order[i]];
      make_meta(b.ctx, d.K, d.max_iterations, d.crc_type, d.F, d.decode_enable ? 1 : 0, (long)in_off[i], (long)out_off[i], &meta[i],
                d.llr8 != 0);
      if (d.dematch_enable) {
        RmParams q;
        if (rm_params(d.K, d.G, d.C, d.Nsoft, d.Mdlharq, d.Kmimo, d.rvidx, d.Qm, d.Nl, d.r, 0, &q))
          return fail(-4, "invalid rate-matching parameters for block %d (lte_rate_matching_turbo_rx returns -1)", order[i]);
        RmBlock rb;
        memset(&rb, 0, sizeof(rb));
        rb.K = d.K; rb.F = d.F; rb.RTC = q.RTC; rb.Kpi = q.Kpi; rb.ND = q.ND; rb.Ncb = q.Ncb; rb.k0 = q.k0; rb.E = q.E;
        rb.clear = d.clear;
        if (d.harq_pool) {
          if (pool && pool != d.harq_pool) return fail(-4, "block %d: all pool-backed blocks of a submit must use the same HARQ pool", order[i]);
          pool = d.harq_pool;
          if (pool->dev != dev) return fail(-4, "block %d: HARQ pool lives on GPU %d, the batch runs on GPU %d", order[i], pool->dev, dev);
          if (d.harq_slot >= pool->n_slots || 3 * q.Kpi > pool->slot_hw)
            return fail(-4, "block %d: HARQ slot %u out of range or too small for K=%d", order[i], d.harq_slot, (int)d.K);
          rb.w_sel = 1; rb.w_off = d.harq_slot * pool->slot_hw;
        } else {
          rb.w_sel = 0; rb.w_off = (uint32_t)w_hw;
        }
        // soft bits that follow the previous block's in the caller's memory (r_offset slices of one e buffer,
        // ulsch_decoding.c:1259) keep that spacing on the device, so that a run goes over with one copy
        // (int8 soft bits, in_fmt 1: a run must start on an int16 boundary of the pool, so only even-length predecessors
        // continue a run -- E is a multiple of Qm*Nl, odd only for odd Nl*Qm, which LTE does not have)
        rb.e_fmt = d.in_fmt ? 1u : 0u;
        const bool from_ul = ul_of[order[i]] >= 0;          // e[] is produced on the device by k_ul_front: nothing to copy
        const size_t e_bytes = from_ul ? 0 : (d.in_fmt ? (size_t)q.E : 2 * (size_t)q.E);
        if (from_ul) {
          const UlFrontDev& u = ulf[ul_of[order[i]]];
          const size_t eo = ((((size_t)u.e_off_hi) << 32) | u.e_off_lo) + ul_roff[order[i]];
          rb.e_fmt = 0;
          rb.e_off_lo = (uint32_t)(eo & 0xffffffffu); rb.e_off_hi = (uint32_t)((unsigned long long)eo >> 32);
          prev_e_end = nullptr;
        } else {
        if (!d.in) return fail(-4, "block %d: null input pointer", order[i]);
        const bool cont = !rm.empty() && !rm_nocopy.back() && (const char*)d.in == prev_e_end && rm.back().e_fmt == rb.e_fmt && !(e_bytes_run & 1);
        if (!cont) { e_hw = (e_hw + 7) & ~(size_t)7; e_bytes_run = 0; }
        prev_e_end = (const char*)d.in + e_bytes;
        e_bytes_run += e_bytes;
        rb.e_off_lo = (uint32_t)(e_hw & 0xffffffffu); rb.e_off_hi = (uint32_t)((unsigned long long)e_hw >> 32);
        }
        rm_nocopy.push_back(from_ul);
        rb.dummy_off = 0xffffffffu;                      // NULL map derived from (K,F): prefix-count table, or in the kernel
        rb.cnt_off = rm_table(b.ctx, d.K, d.F);
        rb.y_off_lo = (uint32_t)(in_off[i] & 0xffffffffu); rb.y_off_hi = (uint32_t)((unsigned long long)in_off[i] >> 32);
        e_hw += (e_bytes + 1) >> 1;
        if (!d.harq_pool) w_hw += (size_t)3 * q.Kpi;
        rb.gold_off = 0xffffffffu; rb.scr_off = d.scr_offset;
        int sq = -1;
        if (d.scr_enable) {
          auto it = seq_of.find(d.scr_c_init);
          if (it == seq_of.end()) { sq = (int)gseq.size(); seq_of[d.scr_c_init] = sq; gseq.push_back(GoldSeq{d.scr_c_init, 0u, 0u}); }
          else sq = it->second;
          gseq[sq].nwords = std::max(gseq[sq].nwords, (d.scr_offset + q.E + 31) / 32);
        }
        rm_seq.push_back(sq);
        rm.push_back(rb); rm_desc.push_back(order[i]);
      }
    }
    phase("block metadata and rate-matching parameters");
    // soft bits of rm blocks [jlo, jhi): one copy per run that is contiguous in the caller's memory, staged only if pageable
    auto copy_e_runs = [&](size_t jlo, size_t jhi, cudaStream_t cs) -> int {
      for (size_t j = jlo; j < jhi;) {
        if (rm_nocopy[j]) { ++j; continue; }
        const oai_cb_desc_t& d0 = descs[rm_desc[j]];
        const size_t eo = ((size_t)rm[j].e_off_hi << 32) | rm[j].e_off_lo;
        // lengths in bytes; a block continues the run when it follows in the caller's memory AND in the device pool
        auto nbytes = [&](size_t x) -> size_t { return rm[x].e_fmt ? (size_t)rm[x].E : 2 * (size_t)rm[x].E; };
        auto eoff = [&](size_t x) -> size_t { return 2 * (((size_t)rm[x].e_off_hi << 32) | rm[x].e_off_lo); };
        size_t len = nbytes(j), k = j + 1;
        while (k < jhi && !rm_nocopy[k] && (const char*)descs[rm_desc[k]].in == (const char*)d0.in + len && eoff(k) == 2 * eo + len) { len += nbytes(k); ++k; }
        cudaPointerAttributes at;
        const bool pinned = (cudaPointerGetAttributes(&at, d0.in) == cudaSuccess) && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        const void* src = d0.in;
        if (!pinned) { memcpy(h_e + eo, d0.in, len); src = h_e + eo; }
        CU(cudaMemcpyAsync(d_e + eo, src, len, cudaMemcpyHostToDevice, cs));
        j = k;
      }
      return 0;
    };
    int16_t* const hp = pool ? pool->d : nullptr;
    const size_t deint_smem = 3 * (32 * ((Kmax + 4 + 31) / 32)) * sizeof(int16_t);
    // when every block of the batch goes through the front end (and is a 16-bit block), sub-block deinterleaving is
    // fused into k_demux16: the decoder input y is never materialised
    const bool sw = (flags & OAI_BATCH_SLIDING_WINDOW) != 0;
    b.sw_mode = sw ? 1 : 0;
    const bool fuse_deint = (rm.size() == (size_t)n) && (n == n16) && !sw;
    auto front_end = [&](size_t jlo, size_t jhi, cudaStream_t fs) {   // dematch (+ deinterleave) of rm blocks [jlo, jhi) on fs
      const int cnt = (int)(jhi - jlo);
      k_rm_rx<<<cnt, RM_THREADS, 0, fs>>>(d_rm + jlo, cnt, d_w, d_e, nullptr, hp, gseq.empty() ? nullptr : d_gold, b.ctx->rm_tab);
      ++g_launches;
      if (!fuse_deint) { k_deint<<<cnt, RM_THREADS, deint_smem, fs>>>(d_rm + jlo, cnt, d_w, d_in, 0, hp); ++g_launches; }
    };
    // block metadata first (pageable source: see PinnedArena for why), before any bulk input copy of this batch
    CU(cudaMemsetAsync(d_out, 0, out_b, st));
    if (n16 > 0) {
      rc = b.set_meta(std::vector<CbMeta>(meta.begin(), meta.begin() + n16), st);
      if (rc) return rc;
    }
    if (!rm.empty()) {
      rc = ensure_rm(e_hw, w_hw, (int)rm.size());
      if (rc) return rc;
      std::vector<int> ulf_seq(ulf.size());
      for (size_t u = 0; u < ulf.size(); ++u) {              // one sequence per uplink allocation (whole words, + 1 for a tail)
        ulf_seq[u] = (int)gseq.size();
        gseq.push_back(GoldSeq{ulf_desc[u].c_init, 0u, ulf[u].Rp * ulf[u].Cmux * ulf[u].Qm / 32 + 1});
      }
      if (!gseq.empty()) {                                   // scrambling sequences of the codewords of this batch
        uint32_t words = 0;
        for (auto& g : gseq) { g.off = words; words += g.nwords; }
        for (size_t j = 0; j < rm.size(); ++j) if (rm_seq[j] >= 0) rm[j].gold_off = gseq[rm_seq[j]].off;
        for (size_t u = 0; u < ulf.size(); ++u) ulf[u].gold_off = gseq[ulf_seq[u]].off;
        if ((int)gseq.size() > cap_gseq) { if (d_gseq) cudaFree(d_gseq); cap_gseq = (int)gseq.size(); CU(cudaMalloc(&d_gseq, sizeof(GoldSeq) * cap_gseq)); }
        if (words > cap_gold) { if (d_gold) cudaFree(d_gold); cap_gold = words; CU(cudaMalloc(&d_gold, sizeof(uint32_t) * cap_gold)); }
        CU(cudaMemcpyAsync(d_gseq, arena.put(gseq.data(), sizeof(GoldSeq) * gseq.size()), sizeof(GoldSeq) * gseq.size(), cudaMemcpyHostToDevice, st));
        k_gold<<<((int)gseq.size() + 63) / 64, 64, 0, st>>>(d_gseq, (int)gseq.size(), d_gold);
        ++g_launches;
      }
      // (after the Gold offsets are known, before the uplink front end's bulk copy)
      CU(cudaMemcpyAsync(d_rm, rm.data(), sizeof(RmBlock) * rm.size(), cudaMemcpyHostToDevice, st));
      if (!ulf.empty()) {
        // uplink front ends: the allocations' soft bits go over, k_ul_front leaves e[] in the batch's soft-bit pool
        rc = ensure_ulf((int)ulf.size(), ul_llr_b, std::max<size_t>(ul_cqi_b, 16));
        if (rc) return rc;
        // allocations that follow each other in the caller's memory (and in the pool) go over with one copy
        auto llr_off = [&](size_t u) -> size_t { return ((size_t)ulf[u].llr_off_hi << 32) | ulf[u].llr_off_lo; };
        auto llr_bytes = [&](size_t u) -> size_t { return (size_t)ulf[u].Rp * ulf[u].Cmux * ulf[u].Qm * (ulf[u].llr_fmt ? 1 : 2); };
        for (size_t u = 0; u < ulf.size();) {
          const size_t off = llr_off(u);
          size_t nb = llr_bytes(u), v = u + 1;
          while (v < ulf.size() && (const char*)ulf_desc[v].llr == (const char*)ulf_desc[u].llr + nb && llr_off(v) == off + nb) { nb += llr_bytes(v); ++v; }
          cudaPointerAttributes at;
          const bool pinned = (cudaPointerGetAttributes(&at, ulf_desc[u].llr) == cudaSuccess) && at.type == cudaMemoryTypeHost;
          cudaGetLastError();
          const void* src = ulf_desc[u].llr;
          if (!pinned) {
            rc = ensure_stage_llr();
            if (rc) return rc;
            memcpy(h_llr + off, ulf_desc[u].llr, nb); src = h_llr + off;
          }
          CU(cudaMemcpyAsync(d_llr + off, src, nb, cudaMemcpyHostToDevice, st));
          u = v;
        }
        CU(cudaMemcpyAsync(d_ulf, arena.put(ulf.data(), sizeof(UlFrontDev) * ulf.size()), sizeof(UlFrontDev) * ulf.size(), cudaMemcpyHostToDevice, st));
        k_ul_front<<<(int)ulf.size(), ULF_THREADS, 0, st>>>(d_ulf, (int)ulf.size(), d_llr, d_e, d_gold, d_cqi, d_ulfout);
        ++g_launches;
        CU(cudaMemcpyAsync(h_ulfout, d_ulfout, sizeof(UlFrontOut) * ulf.size(), cudaMemcpyDeviceToHost, st));
        if (ul_cqi_b) CU(cudaMemcpyAsync(h_cqi, d_cqi, ul_cqi_b, cudaMemcpyDeviceToHost, st));
      }
      if (!fe_parts) {
        rc = copy_e_runs(0, rm.size(), st);
        if (rc) return rc;
        for (size_t j = 0; j < rm.size(); ++j) {
          const oai_cb_desc_t& d = descs[rm_desc[j]];
          if (rm[j].w_sel) continue;                         // soft buffer stays in HBM
          if (d.w) memcpy(h_w + rm[j].w_off, d.w, sizeof(int16_t) * 3 * rm[j].Kpi);
          else memset(h_w + rm[j].w_off, 0, sizeof(int16_t) * 3 * rm[j].Kpi);
        }
        if (w_hw) CU(cudaMemcpyAsync(d_w, h_w, w_hw * sizeof(int16_t), cudaMemcpyHostToDevice, st));
        front_end(0, rm.size(), st);
        if (w_hw) CU(cudaMemcpyAsync(h_w, d_w, w_hw * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
      }
    }
    phase("front-end tables, soft-bit copies enqueued");
    // device->host: when the callers' decoded_bytes are laid out like the device output (back to back,
    // every block decoded) in page-locked memory, the result is copied straight into them
    direct_out = false;
    {
      uint8_t* base = descs[order[0]].decoded_bytes;
      bool ok = base != nullptr;
      for (int i = 0; ok && i < n; ++i) {
        const oai_cb_desc_t& d = descs[order[i]];
        ok = d.decode_enable && d.max_iterations >= 2 && d.decoded_bytes == base + out_off[i];
      }
      if (ok) {
        cudaPointerAttributes at;
        ok = (cudaPointerGetAttributes(&at, base) == cudaSuccess) && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
      }
      direct_out = ok;
      want_cb_out = false;                                   // no descriptor wants its block's bytes (transport-block outputs only)
      for (int i = 0; i < n && !want_cb_out; ++i) want_cb_out = descs[order[i]].decoded_bytes != nullptr;
      if (!direct_out && want_cb_out) { rc = ensure_stage_out(); if (rc) return rc; }
    }
    if (parts == 1) {
      rc = enqueue_inputs(0, n, st);
      if (rc) return rc;
    }
    // Pipelined form.  Everything the parts depend on (metadata, tables of the front end, the zeroed output) is on st
    // by now; the part streams wait for it once.  Per part: input copy on the copy stream (packed to int8 on the host
    // first when the narrow feed is on and the values allow it), decode on one of the PART_STREAMS compute streams,
    // decoded bytes back on the output stream.  The small metadata copies above were issued BEFORE the first input copy:
    // copies of one direction execute in issue order, and a metadata copy queued behind the inputs would serialise
    // everything (measured: 48 instead of 32 ms per 42624 blocks).
    if (parts > 1) {
      CU(cudaEventRecord(ev_setup, st));
      for (auto& s2 : st_part) CU(cudaStreamWaitEvent(s2, ev_setup, 0));
      bool narrow = !fe_parts && !sw && host_pack_threads() > 0 && !getenv("OAI_TURBO_NO_NARROW_FEED");
      if (narrow && g_pack_pause.load() > 0) { g_pack_pause.fetch_sub(1); narrow = false; }
      if (narrow) { rc = ensure_in8(in_hw); if (rc) return rc; }
      double pack_s = 0, pack_bytes = 0;
      for (int part = 0; part < parts; ++part) {
        const int lo = part_lo(part), hi = part_lo(part + 1);
        int in8 = 0;
        if (narrow) {
          // runs of blocks that are contiguous in the caller's memory are packed in one pass
          const auto t0 = std::chrono::steady_clock::now();
          int bad = 0;
          for (int i = lo; i < hi && !bad;) {
            const int16_t* base = descs[order[i]].in;
            size_t len = (size_t)3 * descs[order[i]].K + 12;
            int j = i + 1;
            while (j < hi && descs[order[j]].in == base + len) { len += (size_t)3 * descs[order[j]].K + 12; ++j; }
            bad = host_pack_i16_to_i8(base, h_in8 + in_off[i], len);
            pack_bytes += 2.0 * (double)len;
            i = j;
          }
          pack_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
          if (!bad) {
            const size_t o0 = in_off[lo], o1 = in_off[hi - 1] + (size_t)3 * descs[order[hi - 1]].K + 12;
            CU(cudaMemcpyAsync(d_in8 + o0, h_in8 + o0, o1 - o0, cudaMemcpyHostToDevice, st_copy));
            in8 = 1;
          } else if (bad < 0) narrow = false;
          // (bad > 0: this part holds values beyond int8 and goes over as int16 below; later parts are tried again)
        }
        if (!in8) {
          rc = fe_parts ? copy_e_runs(lo, hi, st_copy) : enqueue_inputs(lo, hi, st_copy);
          if (rc) return rc;
        }
        CU(cudaEventRecord(ev_part[part], st_copy));
        cudaStream_t sp = st_part[part % PART_STREAMS];
        CU(cudaStreamWaitEvent(sp, ev_part[part], 0));
        if (fe_parts) front_end(lo, hi, sp);
        rc = fuse_deint ? b.decode16(d_in, d_out, d_status, sp, lo, hi - lo, part, d_rm, d_w, hp)
                        : b.decode16(in8 ? reinterpret_cast<const int16_t*>(d_in8) : d_in, d_out, d_status, sp, lo, hi - lo, part,
                                     nullptr, nullptr, nullptr, in8);
        if (rc < 0) return rc;
        // this part's decoded bytes go back while the next parts are still being copied in / decoded
        CU(cudaEventRecord(ev_done[part], sp));
        CU(cudaStreamWaitEvent(st_out, ev_done[part], 0));
        const size_t o0 = out_off[lo], o1 = (size_t)out_off[hi - 1] + (descs[order[hi - 1]].K >> 3);
        if (want_cb_out) CU(cudaMemcpyAsync((direct_out ? descs[order[0]].decoded_bytes : h_out) + o0, d_out + o0, o1 - o0, cudaMemcpyDeviceToHost, st_out));
      }
      for (int part = 0; part < parts; ++part) CU(cudaStreamWaitEvent(st, ev_done[part], 0));
      // a pack pass slower than ~1.3x the link gains nothing: pause the narrow feed for the next batches
      if (pack_s > 0 && pack_bytes / pack_s / 1e9 < pack_min_gbs()) g_pack_pause.store(16);
      if (trace && pack_s > 0) fprintf(stderr, "[trace %p] narrow feed: packed %.0f MB at %.1f GB/s\n", (void*)this, pack_bytes / 1e6, pack_bytes / pack_s / 1e9);
    }
    phase("parts enqueued");
    if (trace) cudaEventRecord(ev[1], st);
    if (n16 > 0 && parts == 1) {
      rc = fuse_deint ? b.decode16(d_in, d_out, d_status, st, 0, -1, 0, d_rm, d_w, hp) : b.decode16(d_in, d_out, d_status, st);
      if (rc < 0) return rc;
    }
    if (n > n16) {
      rc = b8.set_meta(std::vector<CbMeta>(meta.begin() + n16, meta.end()), st);
      if (rc) return rc;
      rc = b8.decode8(d_in, d_out, d_status + n16, st);
      if (rc < 0) return rc;
    }
    if (!tbs.empty()) {
      // transport-block reassembly + return values on the device (tb_kernels.cuh); runs after every decode of the batch
      std::vector<int> gpu_of(descs.size(), -1);
      for (int i = 0; i < n; ++i) gpu_of[order[i]] = i;
      size_t pool_b = 0;
      for (const auto& t : tbs) {
        TbMeta m;
        memset(&m, 0, sizeof(m));
        m.first = (uint32_t)tb_blk.size(); m.C = t.C; m.F = descs[t.first_cb].F; m.b_off = (uint32_t)pool_b;
        m.uplink = t.uplink ? 1 : 0; m.max_iter = descs[t.first_cb].max_iterations;
        m.stop_after_failure = (flags & OAI_BATCH_DL_STOP_AFTER_FAILURE) ? 1 : 0;
        size_t tb_bytes = 0;
        for (uint32_t r = 0; r < t.C; ++r) {
          const int di = (int)(t.first_cb + r), gi = gpu_of[di];
          // blocks that cannot deliver bytes count as failed: not on the GPU, not decoded, or fewer than 2 iterations
          // (no hard decision, TD16:1267)
          const bool usable = gi >= 0 && descs[di].decode_enable && descs[di].max_iterations >= 2;
          tb_blk.push_back(TbBlk{usable ? gi : -1, usable ? out_off[gi] : 0u, (uint32_t)(descs[di].K >> 3)});
          tb_bytes += descs[di].K >> 3;
        }
        tb_meta.push_back(m);
        pool_b += (tb_bytes + 15) & ~(size_t)15;
      }
      rc = ensure_tb((int)tbs.size(), (int)tb_blk.size(), pool_b);
      if (rc) return rc;
      CU(cudaMemcpyAsync(d_tbmeta, arena.put(tb_meta.data(), sizeof(TbMeta) * tb_meta.size()), sizeof(TbMeta) * tb_meta.size(), cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(d_tbblk, arena.put(tb_blk.data(), sizeof(TbBlk) * tb_blk.size()), sizeof(TbBlk) * tb_blk.size(), cudaMemcpyHostToDevice, st));
      k_tb_assemble<<<(int)tbs.size(), TB_THREADS, 0, st>>>(d_tbmeta, (int)tbs.size(), d_tbblk, d_out, d_status, d_tbpool, d_tbres);
      ++g_launches;
      CU(cudaMemcpyAsync(h_tbpool, d_tbpool, pool_b, cudaMemcpyDeviceToHost, st));
      CU(cudaMemcpyAsync(h_tbres, d_tbres, sizeof(TbResult) * tbs.size(), cudaMemcpyDeviceToHost, st));
    }
    if (trace) cudaEventRecord(ev[2], st);
    const size_t out_used = (size_t)out_off[n - 1] + (descs[order[n - 1]].K >> 3);
    if (parts > 1) {
      CU(cudaEventRecord(ev_out, st_out));
      CU(cudaStreamWaitEvent(st, ev_out, 0));
    } else if (want_cb_out) {
      CU(cudaMemcpyAsync(direct_out ? descs[order[0]].decoded_bytes : h_out, d_out, direct_out ? out_used : out_b, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaMemcpyAsync(h_status, d_status, n, cudaMemcpyDeviceToHost, st));
    if (trace) cudaEventRecord(ev[3], st);
    phase("done");
    return 0;
  }

  int wait() {
    const int n = (int)order.size();
    if (n) CU(cudaStreamSynchronize(st));
    if (n && ev[0] && getenv("OAI_TURBO_TRACE")) {               // times relative to this batch's own submit event
      float t[4] = {0, 0, 0, 0};
      for (int i = 1; i < 4; ++i) cudaEventElapsedTime(&t[i], ev[0], ev[i]);
      fprintf(stderr, "[trace %p gpu %d] h2d_done %.2f  kernels_done %.2f  d2h_done %.2f ms after submit\n", (void*)this, dev, t[1], t[2], t[3]);
    }
    for (int i = 0; i < n; ++i) {
      const oai_cb_desc_t& d = descs[order[i]];
      if (!d.decode_enable) { if (d.status) *d.status = 0xFE; continue; }
      // the reference leaves decoded_bytes untouched when max_iterations < 2 (no hard decision, TD16:1267)
      if (!direct_out && d.decoded_bytes && d.max_iterations >= 2) memcpy(d.decoded_bytes, h_out + out_off[i], d.K >> 3);
      if (d.status) *d.status = h_status[i];
    }
    for (size_t j = 0; j < rm.size(); ++j) {                     // HARQ buffers back to their owners
      const oai_cb_desc_t& d = descs[rm_desc[j]];
      if (d.w && !rm[j].w_sel) memcpy(d.w, h_w + rm[j].w_off, sizeof(int16_t) * rm[j].Ncb);
    }
    for (size_t u = 0; u < ulf.size(); ++u) {                   // control information of the uplink allocations
      const oai_ul_front_t& f = ulf_desc[u];
      const UlFrontOut& o = h_ulfout[u];
      if (f.q_ACK) memcpy(f.q_ACK, o.q_ACK, sizeof(o.q_ACK));
      if (f.q_RI) memcpy(f.q_RI, o.q_RI, sizeof(o.q_RI));
      if (f.o_ACK) memcpy(f.o_ACK, o.o_ACK, 2);
      if (f.o_RI) f.o_RI[0] = o.o_RI;
      if (f.q_cqi && f.Qprime_CQI) memcpy(f.q_cqi, h_cqi + ulf[u].cqi_off, (size_t)f.Qprime_CQI * f.Qm);
    }
    for (size_t i = 0; i < tbs.size(); ++i) {                   // transport blocks assembled on the device
      const oai_tb_desc_t& t = tbs[i];
      const uint32_t nb = h_tbres[i].valid_bytes;
      if (t.ret) *t.ret = h_tbres[i].ret;
      if (t.valid_bytes) *t.valid_bytes = nb;
      // downlink NACK: valid_bytes is 0 and b stays untouched like in the reference (dlsch_decoding.c:455-469)
      if (t.b && nb) memcpy(t.b, h_tbpool + tb_meta[i].b_off, std::min<uint32_t>(nb, t.b_capacity));
    }
    if (flags & OAI_BATCH_DL_STOP_AFTER_FAILURE) {
      // dlsch_decoding.c:400,417,448-451: after the first failing block of a transport block the
      // remaining ones are not decoded and their c[r] stays zeroed
      std::unordered_map<uint32_t, bool> failed;             // transport block -> a block of it has failed already
      for (size_t i = 0; i < descs.size(); ++i) {
        const oai_cb_desc_t& d = descs[i];
        if (!d.status) continue;
        bool& f = failed[d.tb_id];
        if (f) { *d.status = 0xFE; if (d.decoded_bytes) memset(d.decoded_bytes, 0, d.K >> 3); }
        else if (*d.status != 0xFE && *d.status >= 1 + d.max_iterations) f = true;
      }
    }
    return 0;
  }
};

}  // namespace oai

using namespace oai;

// ======================================================================================
// C ABI
// ======================================================================================
struct oai_turbo_dev_plan {
  Batch b;
  Batch8 b8;
  int ncb; uint16_t K; uint8_t max_it, crc, llr8;
  long y_stride = -1, out_stride = -1;
};

extern "C" {

const char* oai_turbo_b200_version(void) { return "oai_turbo_b200 0.1 (sm_100a)"; }
const char* oai_turbo_b200_last_error(void) { return g_err; }
unsigned long long oai_turbo_b200_launch_count(void) { return g_launches.load(); }

int oai_turbo_dev_plan_create(int ncb, uint16_t K, uint8_t max_iterations, uint8_t crc_type, uint8_t llr8,
                              oai_turbo_dev_plan_t** plan) {
  if (!plan || ncb <= 0) return fail(-1, "bad arguments");
  if (qpp_index(K) < 0 || crc_type > 3) return fail(-1, "illegal K=%d or crc_type=%d", (int)K, (int)crc_type);
  if (llr8 && !td8_domain(K)) return fail(-1, "K=%d is outside the 8-bit decoder's domain (K >= 256, K %% 16 == 0)", (int)K);
  DevCtx* c;
  int rc = ctx_get(-1, &c);
  if (rc) return rc;
  oai_turbo_dev_plan* p = new oai_turbo_dev_plan();
  p->ncb = ncb; p->K = K; p->max_it = max_iterations; p->crc = crc_type; p->llr8 = llr8;
  rc = llr8 ? p->b8.alloc(c, ncb, K) : p->b.alloc(c, ncb, K);
  if (rc) { p->b.release(); p->b8.release(); delete p; return rc; }
  *plan = p;
  return 0;
}

int oai_turbo_dev_decode(oai_turbo_dev_plan_t* p, const int16_t* y_dev, long y_stride, uint8_t* out_dev,
                         long out_stride, uint8_t* status_dev, void* stream) {
  if (!p) return fail(-1, "null plan");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->y_stride != y_stride || p->out_stride != out_stride) {
    std::vector<CbMeta> m(p->ncb);
    DevCtx* c = p->llr8 ? p->b8.ctx : p->b.ctx;
    for (int i = 0; i < p->ncb; ++i)
      make_meta(c, p->K, p->max_it, p->crc, 0, 1, (long)i * y_stride, (long)i * out_stride, &m[i], p->llr8 != 0);
    int rc = p->llr8 ? p->b8.set_meta(m, st) : p->b.set_meta(m, st);
    if (rc) return rc;
    CU(cudaStreamSynchronize(st));      // h_meta is pageable; done once per plan
    p->y_stride = y_stride; p->out_stride = out_stride;
  }
  return p->llr8 ? p->b8.decode8(y_dev, out_dev, status_dev, st) : p->b.decode16(y_dev, out_dev, status_dev, st);
}

int oai_turbo_dev_plan_set_mode(oai_turbo_dev_plan_t* p, unsigned flags) {
  if (!p) return fail(-1, "null plan");
  if (flags & ~OAI_BATCH_SLIDING_WINDOW) return fail(-1, "oai_turbo_dev_plan_set_mode: unknown flags 0x%x", flags);
  if (p->llr8 && flags) return fail(-1, "the sliding-window mode exists for the 16-bit decoder only");
  p->b.sw_mode = (flags & OAI_BATCH_SLIDING_WINDOW) ? 1 : 0;
  return 0;
}


struct oai_turbo_batch { HostBatch hb; };

// finished batch objects keep their device workspace, pinned staging and stream and are reused
static std::mutex g_pool_mu;
static std::vector<oai_turbo_batch*> g_pool;

int oai_turbo_submit_batch(const oai_cb_desc_t* cbs, int ncb, unsigned flags, int gpu, oai_turbo_batch_t** handle) {
  if (!cbs || ncb <= 0 || !handle) return fail(-1, "bad arguments");
  oai_turbo_batch* h = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!g_pool.empty()) { h = g_pool.back(); g_pool.pop_back(); }
  }
  if (!h) h = new oai_turbo_batch();
  int rc = h->hb.submit(cbs, ncb, flags, gpu);
  if (rc) { h->hb.release(); delete h; return rc; }
  *handle = h;
  return 0;
}

int oai_turbo_submit_tbs(const oai_cb_desc_t* cbs, int ncb, const oai_tb_desc_t* tbs, int ntb, unsigned flags, int gpu,
                         oai_turbo_batch_t** handle) {
  if (!cbs || ncb <= 0 || !handle || ntb < 0 || (ntb > 0 && !tbs)) return fail(-1, "bad arguments");
  oai_turbo_batch* h = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!g_pool.empty()) { h = g_pool.back(); g_pool.pop_back(); }
  }
  if (!h) h = new oai_turbo_batch();
  int rc = h->hb.submit(cbs, ncb, flags, gpu, tbs, ntb);
  if (rc) { h->hb.release(); delete h; return rc; }
  *handle = h;
  return 0;
}

int oai_turbo_wait(oai_turbo_batch_t* h) {
  if (!h) return fail(-1, "null handle");
  int rc = h->hb.wait();
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (rc == 0 && g_pool.size() < 16) g_pool.push_back(h);
  else { h->hb.release(); delete h; }
  return rc;
}

int oai_turbo_dev_plan_profile(oai_turbo_dev_plan_t* p, int enable, double* ms4, long* count4);

int oai_turbo_harq_pool_create(int gpu, uint32_t n_slots, uint16_t max_K, oai_turbo_harq_pool_t** pool) {
  if (!pool || n_slots == 0 || qpp_index(max_K) < 0) return fail(-1, "bad arguments");
  DevCtx* c;
  int rc = ctx_get(gpu, &c);
  if (rc) return rc;
  DevGuard guard;
  if (guard.enter(c->dev)) return fail(-100, "cannot select CUDA device %d", c->dev);
  oai_turbo_harq_pool* p = new oai_turbo_harq_pool();
  p->dev = c->dev; p->n_slots = n_slots;
  p->slot_hw = 3u * 32u * (((uint32_t)max_K + 4 + 31) / 32);
  if ((unsigned long long)p->slot_hw * n_slots > 0xffffffffull) { delete p; return fail(-1, "pool too large for 32-bit offsets"); }
  const size_t bytes = sizeof(int16_t) * (size_t)p->slot_hw * n_slots;
  if (cudaMalloc(&p->d, bytes) != cudaSuccess || cudaMemset(p->d, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    cudaGetLastError(); delete p;
    return fail(-100, "HARQ pool: cudaMalloc of %zu bytes failed", bytes);
  }
  *pool = p;
  return 0;
}
int oai_turbo_harq_pool_read(oai_turbo_harq_pool_t* p, uint32_t slot, int16_t* w_host, uint32_t n) {
  if (!p || !w_host || slot >= p->n_slots || n > p->slot_hw) return fail(-1, "bad arguments");
  DevGuard guard;
  if (guard.enter(p->dev)) return fail(-100, "cannot select CUDA device %d", p->dev);
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(w_host, p->d + (size_t)slot * p->slot_hw, sizeof(int16_t) * n, cudaMemcpyDeviceToHost));
  return 0;
}
void oai_turbo_harq_pool_destroy(oai_turbo_harq_pool_t* p) {
  if (!p) return;
  cudaFree(p->d);
  delete p;
}

// lte_segmentation.c:52-134, parameter part
int oai_lte_segmentation_params(uint32_t B, uint32_t* C, uint32_t* Cplus, uint32_t* Cminus, uint32_t* Kplus,
                                uint32_t* Kminus, uint32_t* F) {
  if (!C || !Cplus || !Cminus || !Kplus || !Kminus || !F) return fail(-1, "null output pointer");
  uint32_t Bp = B;
  *C = 1;
  if (B > 6144) {                                   // :58-70: L = 24 per block once the block is split
    *C = (B + 6119) / 6120;
    Bp = B + 24 * (*C);
  }
  if (*C > 16) return -1;                           // MAX_NUM_DLSCH_SEGMENTS, :72-76
  const uint32_t q = Bp / (*C);
  uint32_t step;
  if (q <= 40) { *Kplus = 40; *Kminus = 0; step = 0; }
  else if (q <= 512) { *Kplus = q & ~7u; *Kminus = q - 8; step = 0; }      // :78-82 (sic: K- from q, not from K+)
  else if (q <= 1024) step = 16;
  else if (q <= 2048) step = 32;
  else if (q <= 6144) step = 64;
  else return -1;                                   // :110-113
  if (step) {
    *Kplus = q & ~(step - 1);
    if (*Kplus < q) *Kplus += step;
    *Kminus = *Kplus - step;
  }
  if (*C == 1) { *Cplus = 1; *Kminus = 0; *Cminus = 0; }
  else {
    *Cminus = ((*C) * (*Kplus) - Bp) / (*Kplus - *Kminus);
    *Cplus = *C - *Cminus;
  }
  *F = (*Cplus) * (*Kplus) + (*Cminus) * (*Kminus) - Bp;
  return 0;
}

// ulsch_decoding.c:381-468
int oai_ulsch_control_sizes(uint32_t O_RI, uint32_t O_ACK, uint32_t Or1, uint32_t Msc_initial, uint32_t Nsymb_initial,
                            uint32_t beta_ri_x8, uint32_t beta_ack_x8, uint32_t beta_cqi_x8, uint32_t sumKr, uint32_t nb_rb,
                            uint32_t Qm, uint32_t Nsymb_pusch, uint32_t* Qprime_RI, uint32_t* Qprime_ACK, uint32_t* Qprime_CQI,
                            uint32_t* G, uint32_t* Hprime, uint32_t* Hpp) {
  if (!Qprime_RI || !Qprime_ACK || !Qprime_CQI || !G || !Hprime || !Hpp || sumKr == 0 || Qm == 0) return fail(-1, "bad arguments");
  auto coded_symbols = [&](uint32_t payload_bits, uint32_t beta_x8, bool capped) -> uint32_t {       // :381-431
    uint32_t q = payload_bits * Msc_initial * Nsymb_initial * beta_x8;
    if (q == 0) return 0;
    q = (q % (8 * sumKr)) ? 1 + q / (8 * sumKr) : q / (8 * sumKr);
    return (capped && q > 4 * nb_rb * 12) ? 4 * nb_rb * 12 : q;
  };
  const uint32_t Gtot = nb_rb * (12 * Qm) * Nsymb_pusch;
  *Qprime_RI = coded_symbols(O_RI, beta_ri_x8, true);
  *Qprime_ACK = coded_symbols(O_ACK, beta_ack_x8, true);
  uint32_t qc = Or1 ? coded_symbols(Or1 + (Or1 < 12 ? 0 : 8), beta_cqi_x8, false) : 0;
  if (qc > Gtot - O_RI) qc = Gtot - O_RI;                                   // :438-439, as written there
  *Qprime_CQI = qc;
  const uint32_t g = Gtot - Qm * (*Qprime_RI) - Qm * qc;                      // :447
  if ((int32_t)g < 0) return -1;                                              // :449-452
  *G = g;
  *Hprime = (g + Qm * qc) / Qm;
  *Hpp = *Hprime + *Qprime_RI;
  return 0;
}

void* oai_turbo_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { fail(-100, "cudaMallocHost(%zu) failed", bytes); cudaGetLastError(); return nullptr; }
  return p;
}
void oai_turbo_host_free(void* p) { if (p) cudaFreeHost(p); }

// one cached single-block batch per calling thread: the call is re-entrant like the reference
// (all scratch is per call there, TD16:967-979) and pays no allocation after the first use
struct SingleHolder {
  HostBatch* hb = nullptr;
  HostBatch* get() { if (!hb) hb = new HostBatch(); return hb; }
  void drop() { if (hb) { hb->release(); delete hb; hb = nullptr; cudaGetLastError(); } }
  ~SingleHolder() { drop(); }                        // thread exit: stream, device workspace and pinned memory go back
};
static thread_local SingleHolder t_single;

static unsigned char decode_one(short* y, unsigned char* decoded_bytes, unsigned short n, unsigned char max_iterations,
                                unsigned char crc_type, unsigned char F, int llr8, const char* who) {
  HostBatch* hb = t_single.get();
  oai_cb_desc_t d;
  memset(&d, 0, sizeof(d));
  uint8_t status = 255;
  d.in = y; d.decoded_bytes = decoded_bytes; d.status = &status; d.K = n; d.max_iterations = max_iterations;
  d.crc_type = crc_type; d.F = F; d.decode_enable = 1; d.llr8 = (uint8_t)llr8;
  if (hb->submit(&d, 1, 0, -1) || hb->wait()) {
    fprintf(stderr, "[oai_turbo_b200] %s: GPU path failed (%s); there is no CPU fallback\n", who, g_err);
    t_single.drop();                                  // a half-built workspace must not serve the next call
    return 255;
  }
  return status;
}

void init_td16(void) {
  DevCtx* c;
  if (ctx_get(-1, &c)) fprintf(stderr, "[oai_turbo_b200] init_td16: no usable CUDA device -- decoder calls will fail\n");
}
// 3gpplte_turbo_decoder_sse_16bit.c:886 / _8bit.c:834 free the interleaver tables; here: the recycled batch objects
// (streams, device workspace, pinned staging) and the per-device tables.  Like the reference's, not to be called while
// other threads decode; the next init/decode call rebuilds what it needs.
void free_td16(void) {
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (oai_turbo_batch* h : g_pool) { h->hb.release(); delete h; }
    g_pool.clear();
  }
  t_single.drop();
  ctx_release_all();
}
void init_td8(void) { init_td16(); }
void free_td8(void) { free_td16(); }

unsigned char phy_threegpplte_turbo_decoder16(short* y, unsigned char* decoded_bytes, unsigned short n,
    unsigned short f1, unsigned short f2, unsigned char max_iterations, unsigned char crc_type, unsigned char F,
    oai_time_stats_t*, oai_time_stats_t*, oai_time_stats_t*, oai_time_stats_t*, oai_time_stats_t*,
    oai_time_stats_t*, oai_time_stats_t*) {
  (void)f1; (void)f2;
  if (crc_type > 3) { fprintf(stderr, "Illegal crc length!\n"); return 255; }          // TD16:1003-1006
  if (qpp_index(n) < 0) { fprintf(stderr, "Illegal frame length!\n"); return 255; }    // TD16:1015-1018
  return decode_one(y, decoded_bytes, n, max_iterations, crc_type, F, 0, "phy_threegpplte_turbo_decoder16");
}

unsigned char phy_threegpplte_turbo_decoder8(short* y, unsigned char* decoded_bytes, unsigned short n, unsigned short f1,
    unsigned short f2, unsigned char max_iterations, unsigned char crc_type, unsigned char F, oai_time_stats_t*,
    oai_time_stats_t*, oai_time_stats_t*, oai_time_stats_t*, oai_time_stats_t*, oai_time_stats_t*, oai_time_stats_t*) {
  (void)f1; (void)f2;
  if (crc_type > 3) { fprintf(stderr, "Illegal crc length!\n"); return 255; }          // TD8:954-957
  if (qpp_index(n) < 0) { fprintf(stderr, "Illegal frame length!\n"); return 255; }    // TD8:969-972
  if (!td8_domain(n)) {
    fprintf(stderr, "[oai_turbo_b200] phy_threegpplte_turbo_decoder8: n=%d is outside the supported domain "
                    "(n >= 256, n %% 16 == 0; the reference overruns its buffers there)\n", (int)n);
    return 255;
  }
  return decode_one(y, decoded_bytes, n, max_iterations, crc_type, F, 1, "phy_threegpplte_turbo_decoder8");
}

uint32_t generate_dummy_w(uint32_t D, uint8_t* w, uint8_t F) {
  const uint32_t RTC = (D >> 5) + ((D & 31) ? 1 : 0), Kpi = RTC << 5, ND = Kpi - D;
  Scratch& sc = scratch_here();
  if (sc.ensure(3 * (size_t)Kpi)) { fprintf(stderr, "[oai_turbo_b200] generate_dummy_w: GPU path failed (%s)\n", g_err); return RTC; }
  memcpy(sc.h, w, 3 * (size_t)Kpi);
  cudaMemcpyAsync(sc.d, sc.h, 3 * (size_t)Kpi, cudaMemcpyHostToDevice, sc.st);
  k_dummy_w<<<(3 * Kpi + 255) / 256, 256, 0, sc.st>>>((uint8_t*)sc.d, RTC, Kpi, ND, F);
  ++g_launches;
  cudaMemcpyAsync(sc.h, sc.d, 3 * (size_t)Kpi, cudaMemcpyDeviceToHost, sc.st);
  if (cudaStreamSynchronize(sc.st) != cudaSuccess) { fail(-100, "generate_dummy_w: CUDA failure"); return RTC; }
  memcpy(w, sc.h, 3 * (size_t)Kpi);
  return RTC;
}

int lte_rate_matching_turbo_rx(uint32_t RTC, uint32_t G, int16_t* w, uint8_t* dummy_w, int16_t* soft_input, uint8_t C,
                               uint32_t Nsoft, uint8_t Mdlharq, uint8_t Kmimo, uint8_t rvidx, uint8_t clear, uint8_t Qm,
                               uint8_t Nl, uint8_t r, uint32_t* E_out) {
  RmParams q;
  if (Kmimo == 0 || Mdlharq == 0 || C == 0 || Qm == 0 || Nl == 0) {
    printf("lte_rate_matching.c: invalid parameters (Kmimo %d, Mdlharq %d, C %d, Qm %d, Nl %d\n", Kmimo, Mdlharq, C, Qm, Nl);
    return -1;
  }
  // K is not an argument of the reference call; everything it needs follows from RTC
  rm_params(32 * RTC - 4, G, C, Nsoft, Mdlharq, Kmimo, rvidx, Qm, Nl, r, RTC, &q);
  Scratch& sc = scratch_here();
  // layout of the scratch buffer: [RmBlock][w: Ncb int16][dummy: Ncb bytes][e: E int16]
  const size_t o_w = 256, o_dm = o_w + (((size_t)q.Ncb * 2 + 255) & ~(size_t)255), o_e = o_dm + (((size_t)q.Ncb + 255) & ~(size_t)255);
  const size_t total = o_e + (size_t)q.E * 2 + 256;
  if (sc.ensure(total)) return fail(-100, "lte_rate_matching_turbo_rx: GPU path failed (%s)", g_err);
  RmBlock b;
  memset(&b, 0, sizeof(b));
  b.K = 32 * RTC - 4; b.F = 0; b.RTC = RTC; b.Kpi = q.Kpi; b.ND = 0; b.Ncb = q.Ncb; b.k0 = q.k0; b.E = q.E; b.clear = clear;
  b.w_off = 0; b.e_off_lo = 0; b.e_off_hi = 0; b.dummy_off = 0; b.gold_off = 0xffffffffu; b.cnt_off = 0xffffffffu;
  char* h = (char*)sc.h; char* d = (char*)sc.d;
  memcpy(h, &b, sizeof(b));
  if (clear != 1) memcpy(h + o_w, w, (size_t)q.Ncb * 2);
  memcpy(h + o_dm, dummy_w, q.Ncb);
  memcpy(h + o_e, soft_input, (size_t)q.E * 2);
  CU(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, sc.st));
  k_rm_rx<<<1, RM_THREADS, 0, sc.st>>>((const RmBlock*)d, 1, (int16_t*)(d + o_w), (const int16_t*)(d + o_e), (const uint8_t*)(d + o_dm));
  ++g_launches;
  CU(cudaMemcpyAsync(h + o_w, d + o_w, (size_t)q.Ncb * 2, cudaMemcpyDeviceToHost, sc.st));
  CU(cudaStreamSynchronize(sc.st));
  memcpy(w, h + o_w, (size_t)q.Ncb * 2);
  *E_out = q.E;
  return 0;
}

void sub_block_deinterleaving_turbo(uint32_t D, int16_t* dd, int16_t* w) {
  const uint32_t RTC = (D >> 5) + ((D & 31) ? 1 : 0), Kpi = RTC << 5, ND = Kpi - D;
  Scratch& sc = scratch_here();
  const size_t o_w = 256, o_y = o_w + (((size_t)3 * Kpi * 2 + 255) & ~(size_t)255), total = o_y + ((size_t)3 * Kpi + 3) * 2 + 256;
  if (sc.ensure(total)) { fprintf(stderr, "[oai_turbo_b200] sub_block_deinterleaving_turbo: GPU path failed (%s)\n", g_err); return; }
  RmBlock b;
  memset(&b, 0, sizeof(b));
  b.K = D - 4; b.RTC = RTC; b.Kpi = Kpi; b.ND = ND; b.w_off = 0; b.y_off_lo = 0; b.y_off_hi = 0;
  char* h = (char*)sc.h; char* d = (char*)sc.d;
  memcpy(h, &b, sizeof(b));
  memcpy(h + o_w, w, (size_t)3 * Kpi * 2);
  cudaMemcpyAsync(d, h, o_y, cudaMemcpyHostToDevice, sc.st);
  k_deint<<<1, RM_THREADS, 3 * Kpi * sizeof(int16_t), sc.st>>>((const RmBlock*)d, 1, (const int16_t*)(d + o_w), (int16_t*)(d + o_y), 1);
  ++g_launches;
  cudaMemcpyAsync(h + o_y, d + o_y, ((size_t)3 * Kpi + 3) * 2, cudaMemcpyDeviceToHost, sc.st);
  if (cudaStreamSynchronize(sc.st) != cudaSuccess) { fail(-100, "sub_block_deinterleaving_turbo: CUDA failure"); return; }
  // the reference writes d1[q], d1 = d - 3*ND, for q in [0, 3*Kpi+3) except q = 2, 3*Kpi, 3*Kpi+1 (:216-231)
  const int16_t* y = (const int16_t*)(h + o_y);
  int16_t* d1 = dd - 3 * (long)ND;
  d1[0] = y[0]; d1[1] = y[1];
  memcpy(d1 + 3, y + 3, ((size_t)3 * Kpi - 3) * 2);
  d1[3 * Kpi + 2] = y[3 * Kpi + 2];
}

// ---- transmit-side mirror -------------------------------------------------------------------
static size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

void threegpplte_turbo_encoder(uint8_t* input, uint16_t input_length_bytes, uint8_t* output, uint8_t F,
                               uint16_t interleaver_f1, uint16_t interleaver_f2) {
  (void)F; (void)interleaver_f1; (void)interleaver_f2;
  const int K = (int)input_length_bytes * 8, idx = qpp_index(K);
  if (idx < 0) { printf("Illegal frame length!\n"); return; }                            // 3gpplte_sse.c:399-402
  Scratch& sc = scratch_here();
  const size_t o_c = 256, o_d = o_c + up256(input_length_bytes), total = o_d + up256(3 * (size_t)K + 12);
  DevCtx* c;
  if (sc.ensure(total) || ctx_get(-1, &c)) { fprintf(stderr, "[oai_turbo_b200] threegpplte_turbo_encoder: GPU path failed (%s)\n", g_err); return; }
  TxBlock b;
  memset(&b, 0, sizeof(b));
  b.K = K; b.f1 = kQpp[idx][0]; b.f2 = kQpp[idx][1]; b.c_off_lo = (uint32_t)o_c; b.d_off_lo = (uint32_t)o_d;
  char* h = (char*)sc.h; char* d = (char*)sc.d;
  memcpy(h, &b, sizeof(b));
  memcpy(h + o_c, input, input_length_bytes);
  cudaMemcpyAsync(d, h, o_d, cudaMemcpyHostToDevice, sc.st);
  k_turbo_enc<<<1, ENC_WARPS * 32, 0, sc.st>>>((const TxBlock*)d, 1, (const uint8_t*)d, (uint8_t*)d);
  ++g_launches;
  cudaMemcpyAsync(h + o_d, d + o_d, 3 * (size_t)K + 12, cudaMemcpyDeviceToHost, sc.st);
  if (cudaStreamSynchronize(sc.st) != cudaSuccess) { fail(-100, "threegpplte_turbo_encoder: CUDA failure"); return; }
  memcpy(output, h + o_d, 3 * (size_t)K + 12);
}

uint32_t sub_block_interleaving_turbo(uint32_t D, uint8_t* dd, uint8_t* w) {
  const uint32_t RTC = (D >> 5) + ((D & 31) ? 1 : 0), Kpi = RTC << 5, ND = Kpi - D;
  dd[3 * D + 2] = dd[2];                                                                  // lte_rate_matching.c:76
  Scratch& sc = scratch_here();
  const size_t nin = 3 * (size_t)Kpi + 3, o_w = up256(nin), total = o_w + up256(3 * (size_t)Kpi);
  if (sc.ensure(total)) { fprintf(stderr, "[oai_turbo_b200] sub_block_interleaving_turbo: GPU path failed (%s)\n", g_err); return RTC; }
  char* h = (char*)sc.h; char* d = (char*)sc.d;
  memcpy(h, dd - 3 * (long)ND, nin);
  cudaMemcpyAsync(d, h, nin, cudaMemcpyHostToDevice, sc.st);
  k_sbi_tx<<<(Kpi + 255) / 256, 256, 0, sc.st>>>((const uint8_t*)d, (uint8_t*)(d + o_w), RTC, Kpi, ND);
  ++g_launches;
  cudaMemcpyAsync(h + o_w, d + o_w, 3 * (size_t)Kpi, cudaMemcpyDeviceToHost, sc.st);
  if (cudaStreamSynchronize(sc.st) != cudaSuccess) { fail(-100, "sub_block_interleaving_turbo: CUDA failure"); return RTC; }
  memcpy(w, h + o_w, 3 * (size_t)Kpi);
  return RTC;
}

uint32_t lte_rate_matching_turbo(uint32_t RTC, uint32_t G, uint8_t* w, uint8_t* e, uint8_t C, uint32_t Nsoft,
                                 uint8_t Mdlharq, uint8_t Kmimo, uint8_t rvidx, uint8_t Qm, uint8_t Nl, uint8_t r,
                                 uint8_t nb_rb, uint8_t m) {
  (void)nb_rb; (void)m;
  RmParams q;
  if (rm_params(32 * RTC - 4, G, C, Nsoft, Mdlharq, Kmimo, rvidx, Qm, Nl, r, RTC, &q)) {
    fprintf(stderr, "[oai_turbo_b200] lte_rate_matching_turbo: invalid parameters (Kmimo %d, Mdlharq %d, C %d, Qm %d, Nl %d)\n",
            Kmimo, Mdlharq, C, Qm, Nl);                                                   // the reference divides by zero here
    return 0;
  }
  if (q.Ncb < 3 * q.Kpi) {                                                                // :508-511
    printf("Exiting, RM condition (Nir %d, Nsoft %d, Kw %d\n", (int)(Nsoft / Kmimo / (Mdlharq < 8 ? Mdlharq : 8)), (int)Nsoft, (int)(3 * q.Kpi));
    return 0;
  }
  Scratch& sc = scratch_here();
  const size_t o_w = 256, o_e = o_w + up256(q.Ncb), total = o_e + up256(q.E);
  if (sc.ensure(total)) { fprintf(stderr, "[oai_turbo_b200] lte_rate_matching_turbo: GPU path failed (%s)\n", g_err); return 0; }
  TxBlock b;
  memset(&b, 0, sizeof(b));
  b.K = 32 * RTC - 4; b.RTC = RTC; b.Kpi = q.Kpi; b.Ncb = q.Ncb; b.k0 = q.k0; b.E = q.E;
  b.w_from_d = 0; b.w_off_lo = (uint32_t)o_w; b.e_off_lo = (uint32_t)o_e;
  char* h = (char*)sc.h; char* d = (char*)sc.d;
  memcpy(h, &b, sizeof(b));
  memcpy(h + o_w, w, q.Ncb);
  cudaMemcpyAsync(d, h, o_e, cudaMemcpyHostToDevice, sc.st);
  k_rm_tx<<<1, RM_THREADS, q.Ncb + 16, sc.st>>>((const TxBlock*)d, 1, nullptr, (const uint8_t*)d, (uint8_t*)d);
  ++g_launches;
  cudaMemcpyAsync(h + o_e, d + o_e, q.E, cudaMemcpyDeviceToHost, sc.st);
  if (cudaStreamSynchronize(sc.st) != cudaSuccess) { fail(-100, "lte_rate_matching_turbo: CUDA failure"); return 0; }
  memcpy(e, h + o_e, q.E);
  return q.E;
}

int oai_turbo_tx_batch(oai_tx_desc_t* blocks, int n, unsigned flags, int gpu) {
  g_err[0] = 0;
  if (n <= 0) return 0;
  if (!blocks) return fail(-1, "oai_turbo_tx_batch: null descriptor array");
  const bool devp = (flags & OAI_TX_DEVICE_POINTERS) != 0;
  DevCtx* c;
  int rc = ctx_get(gpu, &c);
  if (rc) return rc;
  DevGuard guard;
  if (guard.enter(c->dev)) return fail(-100, "cannot select CUDA device %d", c->dev);
  // scratch layout: [TxBlock x n][c bytes][e bytes][d bytes (device only)]
  std::vector<TxBlock> tb(n);
  const size_t o_c = up256(sizeof(TxBlock) * (size_t)n);
  size_t cur_c = o_c, tot_e = 0, tot_d = 0;
  uint32_t max_ncb = 0;
  for (int i = 0; i < n; ++i) {
    oai_tx_desc_t& t = blocks[i];
    const int idx = qpp_index(t.K);
    if (idx < 0) return fail(-3, "oai_turbo_tx_batch: block %d: illegal block size %d", i, (int)t.K);
    if (!t.c || !t.e) return fail(-1, "oai_turbo_tx_batch: block %d: null pointer", i);
    RmParams q;
    if (rm_params(t.K, t.G, t.C, t.Nsoft, t.Mdlharq, t.Kmimo, t.rvidx, t.Qm, t.Nl, t.r, 0, &q))
      return fail(-3, "oai_turbo_tx_batch: block %d: invalid rate-matching parameters", i);
    TxBlock& b = tb[i];
    memset(&b, 0, sizeof(b));
    b.K = t.K; b.F = t.filler_null ? t.F : 0; b.RTC = q.RTC; b.Kpi = q.Kpi; b.ND = q.ND; b.Ncb = q.Ncb; b.k0 = q.k0;
    b.E = (q.Ncb < 3 * q.Kpi) ? 0 : q.E;
    b.f1 = kQpp[idx][0]; b.f2 = kQpp[idx][1]; b.w_from_d = 1;
    t.E = b.E;
    max_ncb = std::max(max_ncb, q.Ncb);
    if (devp) {
      const unsigned long long pc = (unsigned long long)(uintptr_t)t.c, pe = (unsigned long long)(uintptr_t)t.e;
      b.c_off_lo = (uint32_t)pc; b.c_off_hi = (uint32_t)(pc >> 32); b.e_off_lo = (uint32_t)pe; b.e_off_hi = (uint32_t)(pe >> 32);
    } else {
      b.c_off_lo = (uint32_t)cur_c; b.c_off_hi = (uint32_t)((unsigned long long)cur_c >> 32);
      cur_c += ((size_t)t.K / 8 + 3) & ~(size_t)3;
      b.e_off_lo = (uint32_t)tot_e; b.e_off_hi = (uint32_t)((unsigned long long)tot_e >> 32);     // relative, fixed below
      tot_e += b.E;
    }
    b.d_off_lo = (uint32_t)tot_d; b.d_off_hi = (uint32_t)((unsigned long long)tot_d >> 32);
    tot_d += (3 * (size_t)t.K + 12 + 3) & ~(size_t)3;
  }
  const size_t o_e = up256(cur_c), o_d = o_e + up256(tot_e), total = o_d + up256(tot_d);
  Scratch& sc = scratch_here();
  if (sc.ensure(total, o_d)) return fail(-100, "oai_turbo_tx_batch: GPU path failed (%s)", g_err);   // d is device-only
  char* h = (char*)sc.h; char* d = (char*)sc.d;
  for (int i = 0; i < n; ++i) {
    TxBlock& b = tb[i];
    const unsigned long long od = (unsigned long long)o_d + (((unsigned long long)b.d_off_hi << 32) | b.d_off_lo);
    b.d_off_lo = (uint32_t)od; b.d_off_hi = (uint32_t)(od >> 32);
    if (!devp) {
      const unsigned long long oe = (unsigned long long)o_e + (((unsigned long long)b.e_off_hi << 32) | b.e_off_lo);
      b.e_off_lo = (uint32_t)oe; b.e_off_hi = (uint32_t)(oe >> 32);
      memcpy(h + (((unsigned long long)b.c_off_hi << 32) | b.c_off_lo), blocks[i].c, blocks[i].K / 8);
    }
  }
  memcpy(h, tb.data(), sizeof(TxBlock) * (size_t)n);
  CU(cudaMemcpyAsync(d, h, devp ? o_c : o_e, cudaMemcpyHostToDevice, sc.st));
  const uint8_t* cbase = devp ? nullptr : (const uint8_t*)d;
  uint8_t* ebase = devp ? nullptr : (uint8_t*)d;
  k_turbo_enc<<<(n + ENC_WARPS - 1) / ENC_WARPS, ENC_WARPS * 32, 0, sc.st>>>((const TxBlock*)d, n, cbase, (uint8_t*)d);
  k_rm_tx<<<n, RM_THREADS, max_ncb + 16, sc.st>>>((const TxBlock*)d, n, (const uint8_t*)d, nullptr, ebase);
  g_launches += 2;
  if (!devp && tot_e) CU(cudaMemcpyAsync(h + o_e, d + o_e, tot_e, cudaMemcpyDeviceToHost, sc.st));
  CU(cudaStreamSynchronize(sc.st));
  CU(cudaGetLastError());
  if (!devp) {
    size_t off = o_e;
    for (int i = 0; i < n; ++i) { memcpy(blocks[i].e, h + off, tb[i].E); off += tb[i].E; }
  }
  return 0;
}

// Kernel-level test hook: one MAP pass (demux + k_map16) on a single block; the systematic
// input is the channel systematic stream, the parity stream is p1 (term=0) or p2 (term=1).
// policy: 0 = guard decides, 1 = force the non-saturating fast path, 2 = force the exact path.
// ext_out: K values in the reference's lane layout [step*8 + lane].
int oai_turbo_debug_map16(const int16_t* y, uint16_t K, int term, int policy, int16_t* ext_out) {
  if (qpp_index(K) < 0) return fail(-1, "illegal K");
  HostBatch hb;
  size_t in_hw = ((size_t)3 * K + 12 + 7) & ~(size_t)7;
  DevCtx* dctx;
  int rc = ctx_get(-1, &dctx);
  if (rc) return rc;
  rc = hb.ensure(dctx, 1, K, in_hw, 1024);
  if (rc) return rc;
  rc = hb.ensure_stage_in();
  if (rc) return rc;
  memcpy(hb.h_in, y, sizeof(int16_t) * (3 * (size_t)K + 12));
  std::vector<CbMeta> meta(1);
  make_meta(hb.b.ctx, K, 1, 1, 0, 1, 0, 0, &meta[0]);
  CU(cudaMemcpyAsync(hb.d_in, hb.h_in, in_hw * sizeof(int16_t), cudaMemcpyHostToDevice, hb.st));
  rc = hb.b.set_meta(meta, hb.st);
  if (rc) return rc;
  Batch& b = hb.b;
  XchgArgs x{};
  x.meta = b.d_meta; x.state = b.d_state; x.ws = b.d_ws; x.slot_hw = b.slot_hw; x.A = b.A; x.nblk = 1;
  x.pi_pool = b.ctx->pi_pool; x.t_pool = b.ctx->t_pool; x.crc_xp = b.ctx->crc_xp; x.in_base = hb.d_in; x.out_base = hb.d_out;
  x.status_out = nullptr; x.iter = 0; x.guard_b = GUARD_B; x.batch_max = b.d_batch_max; x.active = nullptr; x.nactive = nullptr; x.nactive_next = nullptr; x.rm = nullptr; x.w_pool = nullptr; x.harq_pool = nullptr;
  cudaMemsetAsync(b.d_batch_max, 0, sizeof(int), hb.st);
  k_demux16<<<1, XCHG_THREADS, 3 * b.A * sizeof(int16_t), hb.st>>>(x);
  MapArgs mp{};
  mp.meta = b.d_meta; mp.state = b.d_state; mp.ws = b.d_ws; mp.slot_hw = b.slot_hw; mp.A = b.A;
  mp.ckpt = b.d_ckpt; mp.ckpt_words = b.ckpt_words; mp.nblk = 1; mp.batch_max = (policy == 3) ? b.d_batch_max : nullptr;
  mp.guard_b = policy == 1 ? 0x7fffffff : (policy == 2 ? -1 : GUARD_B);
  mp.sys_arr = ARR_S0; mp.par_arr = term ? ARR_P2 : ARR_P1; mp.out_arr = ARR_EXT; mp.term = term; mp.iter = 1; mp.upd = 0;
  mp.active = nullptr; mp.nactive = nullptr;
  mp.track = (policy == 0 || policy == 4) ? g_track : 0;
  mp.force = (policy == 4) ? 3 : 0;                          // 4: force the tracked fast pass (falls back through the retry launch)
  k_map16<MAP_SEG><<<1, MAP_THREADS, MAP_SMEM_BYTES, hb.st>>>(mp);
  if (mp.track || policy == 4) { mp.retry = 1; k_map16<MAP_SEG><<<1, MAP_THREADS, MAP_SMEM_BYTES, hb.st>>>(mp); ++g_launches; }
  g_launches += 2;
  std::vector<int16_t> tmp(b.A);
  CU(cudaMemcpyAsync(tmp.data(), b.d_ws + (long)ARR_EXT * b.A, sizeof(int16_t) * b.A, cudaMemcpyDeviceToHost, hb.st));
  CU(cudaStreamSynchronize(hb.st));
  const int W = K / 8;
  for (int k = 0; k < W; ++k)
    for (int l = 0; l < 8; ++l) ext_out[k * 8 + l] = tmp[c4_hw(k, l)];
  hb.release();
  return 0;
}

// enable/disable per-launch event timing; when ms4/count4 are given, first returns and resets the
// accumulated totals per kernel class {demux, map, x1, x2} (the stream must be idle)
int oai_turbo_dev_plan_profile(oai_turbo_dev_plan_t* p, int enable, double* ms4, long* count4) {
  if (!p) return fail(-1, "null plan");
  Profiler& pr = p->llr8 ? p->b8.prof : p->b.prof;
  pr.collect();
  if (ms4 && count4) {
    for (int i = 0; i < 4; ++i) { ms4[i] = pr.ms[i]; count4[i] = pr.count[i]; }
    pr.reset();
  }
  pr.on = enable != 0;
  return 0;
}

void oai_turbo_dev_plan_destroy(oai_turbo_dev_plan_t* p) {
  if (!p) return;
  p->b.release();
  p->b8.release();
  delete p;
}

}  // extern "C"
