// Per-block exchange kernels of the 16-bit decoder: one CTA per code block, arrays staged
// in shared memory so that every HBM access is a coalesced 128-bit load/store and the
// QPP permutation happens on chip.
//
//   k_demux16 : y (s,p1,p2 triples + 12 tail LLRs) -> S0/P1/P2 in C4 lane layout, tail
//               metrics for the beta start of lane 7, max|y|
//               (reference: 3gpplte_turbo_decoder_sse_16bit.c:1055-1189, 474-520)
//   k_x1_16   : s2 = ext o pi                                          (:1209-1231)
//   k_x2_16   : s1 = (ext2 o pi^-1 (-) ext) (+) s0 ; hard decision, CRC, early exit
//               (:1241-1351)
// (+)/(-) are int16 saturating, always computed exactly here (SatArith) -- only the MAP
// recursions have a guarded non-saturating fast path.
#pragma once
#include "td_common.cuh"
#include "rm_kernels.cuh"

namespace oai {

constexpr int XCHG_THREADS = 256;      // upper bound; short blocks are launched with fewer threads (xchg_threads_for)
// One CTA per code block: a K=512 block has 64 uint4 per array, so 256 threads would leave three quarters of the CTA idle
// and -- at 8 resident CTAs of 256 threads per SM -- the SM mostly waiting on CTA start-up and barriers (65 536 blocks of
// K=512: k_x2_16 228 us per launch with 256 threads).  The CRC stage needs one thread per 32 decoded bits.
inline int xchg_threads_for(int max_K) { return max_K <= 1024 ? 64 : (max_K <= 2048 ? 128 : 256); }
constexpr int CRC_NM = 193;          // 32-bit words of the longest block (+1)

struct XchgArgs {
  const CbMeta* meta;
  CbState* state;
  int16_t* ws;
  long slot_hw;
  int A;                        // halfwords per array (multiple of 32)
  int nblk;
  const uint16_t* pi_pool;      // per K: H[j] (K entries) = C4 halfword index of natural position j
  const uint16_t* t_pool;       // per K: T[h] (A entries, layout order) = H[pi(pos(h))]; padding maps to itself
  const u32* crc_xp;            // [4][32][CRC_NM]: x^(w + r + 32 m) mod P for the four CRCs (block_crc_check)
  const int16_t* in_base;       // batch input (device)
  uint8_t* out_base;            // batch output (device)
  uint8_t* status_out;          // batch status bytes (device), written when a block finishes
  int iter;                     // iteration_cnt of the reference loop (1..max)
  int guard_b;                  // MAP fast-path guard (same value as MapArgs::guard_b)
  int* batch_max;               // max |y| over the whole batch (atomicMax by k_demux16)
  const int* active;            // packed list of the blocks still being decoded, or nullptr = all nblk blocks
  const int* nactive;           // two counters: fast-policy blocks (front of the list), exact-policy blocks (back)
  int* nactive_next;            // k_x1_16 zeroes the counters of the list that k_compact builds after k_x2_16
  // fused front end (k_demux16<true>): block i of the batch is rm block i; its decoder input is read straight out of
  // the rate-dematched circular buffer w (sub-block deinterleaving on the fly) instead of a materialised y
  const RmBlock* rm;
  const int16_t* w_pool;
  const int16_t* harq_pool;
  int in8;                      // k_demux16: in_base points at int8 soft bits (same element offsets), the narrow host feed
};

// Active-block compaction: the kernels address blocks through a packed list of the blocks that are still being
// decoded, so that the MAP kernel's warps (8 blocks each) stay full when blocks of a batch leave at different
// iterations (early termination).  k_compact rebuilds the list after k_demux16 and after every k_x2_16: each CTA
// packs its 1024 blocks (ascending) and reserves its range with one atomicAdd, so neighbouring list entries stay
// neighbours in memory; k_x1_16 zeroes the counter of the list that is built next.
constexpr int COMPACT_THREADS = 1024;
// Three classes are kept apart so that a warp of the MAP kernel (8 blocks, one arithmetic policy per warp) does not mix
// them: blocks whose guard allows the untracked fast pass are packed from the FRONT of the list (count[0]), blocks that get
// the tracked fast pass right behind them (count[1]; the MAP kernel pads each class to whole warps), blocks on the exact
// saturating policy from the BACK (list[nblk-1-j], count[2]).  The class is taken from the maxima known when the list is
// built (k_map16 re-derives the policy from the current ones, so this is grouping only, never correctness).
// class of a running block: 0 untracked fast, 1 tracked fast, 2 exact (mirrors map_policy in td16_map.cuh)
__device__ __forceinline__ int policy_class(const CbState& st, int guard_b, int track) {
  const int B = max(st.max_sys, st.max_in) + st.max_in;
  if (B <= guard_b && ((32491 / (B + 1) - 11) >> 1) >= 1) return 0;
  return (track && !(st.retry & 2) && B + 1 <= 16000 && max(st.cert[0], st.cert[1]) <= 24000) ? 1 : 2;
}
__global__ void __launch_bounds__(COMPACT_THREADS) k_compact(const CbState* state, int nblk, int* list, int* count, int guard_b, int track) {
  __shared__ int wsum[3][COMPACT_THREADS / 32];
  __shared__ int base[3];
  const int i = blockIdx.x * COMPACT_THREADS + threadIdx.x;
  const bool on = (i < nblk) && (state[i].status == 0);
  const int cls = on ? policy_class(state[i], guard_b, track) : -1;
  const unsigned bal0 = __ballot_sync(0xffffffffu, cls == 0), bal1 = __ballot_sync(0xffffffffu, cls == 1), bal2 = __ballot_sync(0xffffffffu, cls == 2);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { wsum[0][wid] = __popc(bal0); wsum[1][wid] = __popc(bal1); wsum[2][wid] = __popc(bal2); }
  __syncthreads();
  if (threadIdx.x < 3) {
    int t = 0;
    for (int w = 0; w < COMPACT_THREADS / 32; ++w) { const int c = wsum[threadIdx.x][w]; wsum[threadIdx.x][w] = t; t += c; }
    base[threadIdx.x] = t ? atomicAdd(count + threadIdx.x, t) : 0;
  }
  __syncthreads();
  // the fast and tracked classes share the front of the list: fast blocks at [0, nf), tracked ones at [nf, nf + nt).  nf is
  // only known once every CTA has added its count, so the tracked class is written from the back of the FRONT HALF:
  // entries nblk/2.. are not usable for that; instead it gets its own region: list2 = list + nblk (the list has 2 nblk slots
  // per direction, see the allocation) -- fast [0..), tracked list[nblk + j], exact list[nblk - 1 - j].
  const unsigned lt = (1u << lane) - 1u;
  if (cls == 0) list[base[0] + wsum[0][wid] + __popc(bal0 & lt)] = i;
  if (cls == 1) list[nblk + base[1] + wsum[1][wid] + __popc(bal1 & lt)] = i;
  if (cls == 2) list[nblk - 1 - (base[2] + wsum[2][wid] + __popc(bal2 & lt))] = i;
}

// entry `gi` of the list as the MAP kernel walks it: fast class [0, nf) padded to a multiple of 8 (one warp of k_map16),
// then the tracked class (padded likewise), then the exact class; -1: no block
__device__ __forceinline__ int active_block(const int* list, const int* count, int nblk, int gi) {
  const int nf = count[0], nt = count[1], nx = count[2], nfp = (nf + 7) & ~7, ntp = (nt + 7) & ~7;
  if (gi < nf) return list[gi];
  if (gi >= nfp && gi - nfp < nt) return list[nblk + (gi - nfp)];
  if (gi >= nfp + ntp && gi - nfp - ntp < nx) return list[nblk - 1 - (gi - nfp - ntp)];
  return -1;
}

// block handled by this CTA of an exchange kernel (-1: none)
#ifndef XCHG_LIST_MODE
#define XCHG_LIST_MODE 0      // bit 0: k_x1_16 uses the packed list, bit 1: k_x2_16.  Measured: the two dependent loads at
                              // CTA start cost k_x2_16 21 % and k_x1_16 4 %, while a CTA of a finished block exits at once
                              // anyway -- so only k_map16 (whose warps carry 8 blocks) goes through the list
#endif
__device__ __forceinline__ int xchg_block(const XchgArgs& p, bool use_list = true) {
  const int bi = blockIdx.x;
  if (!use_list) return (bi < p.nblk) ? bi : -1;
  if (p.active) return active_block(p.active, p.nactive, p.nblk, bi);
  return (bi < p.nblk) ? bi : -1;
}

__device__ __forceinline__ int blk_max_reduce(int v, int* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    int x = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) x = max(x, __shfl_xor_sync(0xffffffffu, x, o));
    if (threadIdx.x == 0) red[0] = x;
  }
  __syncthreads();
  return red[0];
}

__device__ __forceinline__ void warp_max_to(int v, int* dst) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) atomicMax(dst, v);
}

__device__ __forceinline__ int absmax2(u32 x) { return max(abs(lo16(x)), abs(hi16(x))); }

// running per-halfword max and min (2 instructions per word); |.|max is taken once at the end
struct MinMax2 {
  u32 mx = 0x80008000u, mn = 0x7fff7fffu;
  __device__ __forceinline__ void add(u32 w) { mx = __vmaxs2(mx, w); mn = __vmins2(mn, w); }
  __device__ __forceinline__ void add(const uint4& v) { add(v.x); add(v.y); add(v.z); add(v.w); }
  __device__ __forceinline__ int absmax() const {
    return max(max(lo16(mx), hi16(mx)), max(-lo16(mn), -hi16(mn)));      // >= 0 once a value was added
  }
};

// out word q = in word (q + r) & 3.  The QPP interleaver maps the 8 lanes of step k to the 8 lanes of step
// pi(k) mod W (contention-free property), and pi is a bijection mod 8 when 8 | W; the shared-memory bank of
// an element in the C4 layout is fixed by (step mod 8, lane pair).  A warp holds 8 consecutive chunks; when
// chunk c starts at step (c>>1)&3 its 8 chunks touch 8 different steps mod 8 per instruction, so the 32
// scattered 16-bit accesses fall into 32 different banks.
__device__ __forceinline__ uint4 rot4(uint4 v, int r) {
  if (r & 1) v = make_uint4(v.y, v.z, v.w, v.x);
  if (r & 2) v = make_uint4(v.z, v.w, v.x, v.y);
  return v;
}

// carry-less 32 x 32 -> 64 bit multiplication from 16 integer multiplications: the operands are split into
// 4 classes of every 4th bit, so that at most 8 partial products meet in one bit position and the 3 hole
// bits above it absorb the carries
__device__ __forceinline__ unsigned long long clmul32(u32 x, u32 y) {
  typedef unsigned long long u64;
  const u32 x0 = x & 0x11111111u, x1 = x & 0x22222222u, x2 = x & 0x44444444u, x3 = x & 0x88888888u;
  const u32 y0 = y & 0x11111111u, y1 = y & 0x22222222u, y2 = y & 0x44444444u, y3 = y & 0x88888888u;
  u64 z0 = ((u64)x0 * y0) ^ ((u64)x1 * y3) ^ ((u64)x2 * y2) ^ ((u64)x3 * y1);
  u64 z1 = ((u64)x0 * y1) ^ ((u64)x1 * y0) ^ ((u64)x2 * y3) ^ ((u64)x3 * y2);
  u64 z2 = ((u64)x0 * y2) ^ ((u64)x1 * y1) ^ ((u64)x2 * y0) ^ ((u64)x3 * y3);
  u64 z3 = ((u64)x0 * y3) ^ ((u64)x1 * y2) ^ ((u64)x2 * y1) ^ ((u64)x3 * y0);
  return (z0 & 0x1111111111111111ull) | (z1 & 0x2222222222222222ull) | (z2 & 0x4444444444444444ull) | (z3 & 0x8888888888888888ull);
}

// ------------------------------------------------------------------------------------
// FE = false: y comes from the batch input.  FE = true: sub_block_deinterleaving_turbo (lte_rate_matching.c:193-243) is
// applied on the fly to the block's circular buffer w, staged in shared memory behind the three demux arrays:
// d[3j] = w[k(j)], d[3j+1] = w[Kpi+2k(j)], d[3j+5] = w[Kpi+2k(j)+1], k(j) = bitrev5(j&31)*RTC + (j>>5), y = d + 3*ND.
#ifndef DEMUX_MIN_CTAS
#define DEMUX_MIN_CTAS 6
#endif
template <bool FE, int TH = XCHG_THREADS>
__global__ void __launch_bounds__(TH, DEMUX_MIN_CTAS * (XCHG_THREADS / TH) > 32 ? 32 : DEMUX_MIN_CTAS * (XCHG_THREADS / TH)) k_demux16_t(XchgArgs p) {
  extern __shared__ int16_t sm[];
  __shared__ int red[XCHG_THREADS / 32];
  const int blk = blockIdx.x;
  if (blk >= p.nblk) return;
  const CbMeta m = p.meta[blk];
  CbState* st = &p.state[blk];
  if (!(m.flags & 1)) { if (threadIdx.x == 0) st->status = 0xFE; return; }
  const int K = m.K, A = p.A;
  const int16_t* y = p.in_base + (((long)m.in_off_hi << 32) | m.in_off_lo);
  const int8_t* y8 = reinterpret_cast<const int8_t*>(p.in_base) + (((long)m.in_off_hi << 32) | m.in_off_lo);
  const bool in8 = !FE && p.in8;
  int16_t* s0 = sm, *p1 = sm + A, *p2 = sm + 2 * A;
  const int16_t* sw = sm + 3 * A;
  uint32_t RTC = 0, Kpi = 0, ND = 0;
  for (int i = threadIdx.x; i < 3 * A / 2; i += TH) reinterpret_cast<u32*>(sm)[i] = 0;
  if (FE) {
    const RmBlock b = p.rm[blk];
    RTC = b.RTC; Kpi = b.Kpi; ND = b.ND;
    const u32* w = reinterpret_cast<const u32*>((b.w_sel ? p.harq_pool : p.w_pool) + b.w_off);
    u32* dst = reinterpret_cast<u32*>(sm + 3 * A);
    for (uint32_t i = threadIdx.x; i < 3 * Kpi / 2; i += TH) dst[i] = w[i];
  }
  __syncthreads();
  auto kof = [&](uint32_t j) -> uint32_t { return brev5(j & 31) * RTC + (j >> 5); };
  // decoder input element idx (0 .. 3K+11): the reference's d[3*ND + idx]
  auto ysrc = [&](int idx) -> int {
    if (!FE) return in8 ? (int)y8[idx] : (int)y[idx];
    const uint32_t q = 3 * ND + idx, r = q % 3;
    const uint32_t k = kof(r == 2 ? (q - 5) / 3 : q / 3);
    return r == 0 ? sw[k] : (r == 1 ? sw[Kpi + 2 * k] : sw[Kpi + 2 * k + 1]);
  };
  int mx = 0;
  const uint16_t* H = p.pi_pool + m.pi_off;
  for (int pos = threadIdx.x; pos < K; pos += TH) {
    const int h = H[pos];
    int v0, v1, v2;
    if (FE) {
      const uint32_t k0 = kof(ND + pos), k2 = kof(ND + pos - 1);      // ND >= 1 for every legal K (K+4 is never a multiple of 32)
      v0 = sw[k0]; v1 = sw[Kpi + 2 * k0]; v2 = sw[Kpi + 2 * k2 + 1];
    } else if (in8) {
      v0 = y8[3 * pos]; v1 = y8[3 * pos + 1]; v2 = y8[3 * pos + 2];
    } else {
      v0 = y[3 * pos]; v1 = y[3 * pos + 1]; v2 = y[3 * pos + 2];
    }
    s0[h] = (int16_t)v0; p1[h] = (int16_t)v1; p2[h] = (int16_t)v2;
    mx = max(max(mx, abs(v0)), max(abs(v1), abs(v2)));
  }
  if (threadIdx.x < 12) mx = max(mx, abs(ysrc(3 * K + threadIdx.x)));
  __syncthreads();
  int16_t* slot = p.ws + (long)blk * p.slot_hw;
  for (int i = threadIdx.x; i < A / 8; i += TH) {
    reinterpret_cast<uint4*>(slot + (long)ARR_S0 * A)[i] = reinterpret_cast<uint4*>(s0)[i];
    reinterpret_cast<uint4*>(slot + (long)ARR_P1 * A)[i] = reinterpret_cast<uint4*>(p1)[i];
    reinterpret_cast<uint4*>(slot + (long)ARR_P2 * A)[i] = reinterpret_cast<uint4*>(p2)[i];
    // int8 copies (low bytes; meaningful only when |y| <= 127, which the batch flag certifies)
    const uint4 a = reinterpret_cast<uint4*>(p1)[i], b = reinterpret_cast<uint4*>(p2)[i], c = reinterpret_cast<uint4*>(s0)[i];
    int8_t* b8a = reinterpret_cast<int8_t*>(slot + (long)ARR_B8A * A);
    int8_t* b8b = reinterpret_cast<int8_t*>(slot + (long)ARR_B8B * A);
    reinterpret_cast<uint2*>(b8a)[i] = make_uint2(__byte_perm(a.x, a.y, 0x6420), __byte_perm(a.z, a.w, 0x6420));
    reinterpret_cast<uint2*>(b8a + A)[i] = make_uint2(__byte_perm(b.x, b.y, 0x6420), __byte_perm(b.z, b.w, 0x6420));
    reinterpret_cast<uint2*>(b8b)[i] = make_uint2(__byte_perm(c.x, c.y, 0x6420), __byte_perm(c.z, c.w, 0x6420));
  }
  mx = blk_max_reduce(mx, red);
  if (threadIdx.x == 0 && p.batch_max) atomicMax(p.batch_max, mx);
  if (threadIdx.x < 2) {
    // tail-bit beta start metrics in WRAPPING int16 (reference :474-520); the gamma of
    // the tail uses the saturating add/sub + >>1 of compute_gamma16 (:160-161)
    int tl[6];                                             // (x,z) x 3 of encoder 1 / 2
    for (int i = 0; i < 6; ++i) tl[i] = ysrc(3 * K + 6 * threadIdx.x + i);
    // all values are kept as 32-bit ints and wrapped to int16 explicitly (w16): nvcc 12.9 turned
    // an int16_t max chain here into VIMNMX3.S16x2 on half-extended registers and picked a
    // wrong maximum for some inputs
    auto w16 = [](int v) { return (v << 16) >> 16; };
    int m11[3], m10[3];
    for (int i = 0; i < 3; ++i) {
      m11[i] = sat16i((int)tl[2 * i] + (int)tl[2 * i + 1]) >> 1;
      m10[i] = sat16i((int)tl[2 * i] - (int)tl[2 * i + 1]) >> 1;
    }
    int b0 = w16(-m11[2]), b1 = m11[2];
    int b0_2 = w16(b0 - m11[1]), b1_2 = w16(b0 + m11[1]);
    int b2_2 = w16(b1 + m10[1]), b3_2 = w16(b1 - m10[1]);
    int t[8];
    t[0] = w16(b0_2 - m11[0]); t[1] = w16(b0_2 + m11[0]);
    t[2] = w16(b1_2 + m10[0]); t[3] = w16(b1_2 - m10[0]);
    t[4] = w16(b2_2 - m10[0]); t[5] = w16(b2_2 + m10[0]);
    t[6] = w16(b3_2 + m11[0]); t[7] = w16(b3_2 - m11[0]);
    int bm = t[0];
    for (int i = 1; i < 8; ++i) bm = max(bm, t[i]);
    for (int i = 0; i < 8; ++i) st->T[threadIdx.x][i] = (int16_t)w16(t[i] - bm);
  }
  if (threadIdx.x == 0) {
    st->max_in = mx;
    st->max_sys = mx;
    st->retry = 0; st->cert[0] = 0; st->cert[1] = 0; st->max_ext2 = -1;
    // `while (iteration_cnt++ < max_iterations)` with max 0 returns 1 (reference :1201,1384)
    st->status = (m.max_iter == 0) ? 1 : 0;
    if (m.max_iter == 0 && p.status_out) p.status_out[blk] = 1;
  }
}

#define k_demux16 k_demux16_t<false>

// ------------------------------------------------------------------------------------
template <int TH>
__global__ void __launch_bounds__(TH) k_x1_16(XchgArgs p) {
  extern __shared__ int16_t sm[];
  __shared__ int smax;
  if (blockIdx.x == 0 && threadIdx.x < 3 && p.nactive_next) p.nactive_next[threadIdx.x] = 0;   // the list k_compact fills next
  const int blk = xchg_block(p, (XCHG_LIST_MODE & 1) != 0);
  if (blk < 0) return;
  const CbMeta m = p.meta[blk];
  CbState* st = &p.state[blk];
  if (st->status != 0 || p.iter > m.max_iter) return;
  if (threadIdx.x == 0) smax = 0;
  const int A = p.A;
  int16_t* slot = p.ws + (long)blk * p.slot_hw;
  const uint4* gext = reinterpret_cast<const uint4*>(slot + (long)ARR_EXT * A);
  uint4* gsys = reinterpret_cast<uint4*>(slot + (long)ARR_SYS * A);
  const int n8 = c4_words(m.W) >> 2;           // uint4 groups actually used by this K
  int16_t* in = sm;
  // (the feedback ext = (ext (-) s1) (+) s0 of reference :1354-1375 is applied by the MAP kernel
  // that produced ext, see MapArgs::upd)
  MinMax2 mm;
  for (int i = threadIdx.x; i < n8; i += TH) {
    const uint4 e = gext[i];
    mm.add(e);
    reinterpret_cast<uint4*>(in)[i] = e;
  }
  __syncthreads();
  const uint4* T4 = reinterpret_cast<const uint4*>(p.t_pool + m.t_off);
  for (int i = threadIdx.x; i < n8; i += TH) {      // s2[st(i)] = ext[st(pi(i))], :1209-1231
    // thread (chunk c, lane pair t) walks its 4 steps starting at step (c>>1)&3: the 8 chunks of a warp then
    // address 8 different rows mod 8 in every instruction -> conflict-free 16-bit accesses (see rot4)
    const int r = (i >> 3) & 3;
    const uint4 tt = rot4(__ldg(T4 + i), r);
    const u32 tw[4] = {tt.x, tt.y, tt.z, tt.w};
    u32 o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const u32 lo = (uint16_t)in[tw[q] & 0xffffu], hi = (uint16_t)in[tw[q] >> 16];
      o[q] = lo | (hi << 16);
    }
    gsys[i] = rot4(make_uint4(o[0], o[1], o[2], o[3]), (4 - r) & 3);
  }
  // s2 is a permutation of ext: same maximum
  warp_max_to(mm.absmax(), &smax);
  __syncthreads();
  if (threadIdx.x == 0) { st->max_sys = smax; st->max_ext = smax; }
}

// ------------------------------------------------------------------------------------
// Hard decision of 32 consecutive natural positions j0..j0+31 (clipped to K) from the natural-order
// C4 array in shared memory, generic form (any W): bit = (sat8(x) > 0) = (x > 0), first position in
// bit 31 (reference TD16:1267-1302: MSB-first packing).
__device__ __forceinline__ u32 hd_word32(const int16_t* nat, int j0, int K, int W) {
  int lane = j0 / W, k = j0 - lane * W;
  u32 bits = 0;
  for (int q = 0; q < 32; ++q) {
    const bool bit = (j0 + q < K) && (nat[c4_hw(k, lane)] > 0);
    bits = (bits << 1) | (bit ? 1u : 0u);
    if (++k >= W) { k = 0; ++lane; }
  }
  return bits;
}

// the two 4-bit hard-decision groups of one C4 uint4 (4 steps x lanes 2t, 2t+1): returns the nibble of lane
// 2t in bits 0..3 and of lane 2t+1 in bits 16..19, first step in the nibble's bit 3
__device__ __forceinline__ u32 hd_nibbles(const uint4& d) {
  return (__vcmpgts2(d.x, 0u) & 0x00080008u) | (__vcmpgts2(d.y, 0u) & 0x00040004u) |
         (__vcmpgts2(d.z, 0u) & 0x00020002u) | (__vcmpgts2(d.w, 0u) & 0x00010001u);
}

// 8 nibbles stored one per byte at index n^7 (n = position/4) -> the 32 decisions, first position in bit 31
__device__ __forceinline__ u32 nibbles_to_word(uint2 v) {
  u32 lo = v.x & 0x0F0F0F0Fu, hi = v.y & 0x0F0F0F0Fu;
  lo = (lo | (lo >> 4)) & 0x00FF00FFu; lo = (lo | (lo >> 8)) & 0xFFFFu;
  hi = (hi | (hi >> 4)) & 0x00FF00FFu; hi = (hi | (hi >> 8)) & 0xFFFFu;
  return (hi << 16) | lo;
}

// CRC early-termination test of one block.  Thread w < ceil(K/32) owns the decisions of positions
// 32w..32w+31 (`bits`, first position in bit 31); all threads of the CTA call this; the result is valid
// in thread 0.  Also writes the decoded bytes to the output.  The reference walks n-w(-F) bits starting
// at byte F>>3 (CRC24A skips the filler bits, TD16:1312-1313) and compares with the trailing crc bytes.
// A CRC with zero start value and no final xor is the polynomial remainder M(x) x^w mod P, linear over
// GF(2): thread w multiplies its 32 message bits by x^(distance to the message end + w) mod P (one
// coalesced table read + a carry-less multiplication), the products are XORed over the CTA and thread 0
// reduces the 32+w-bit sum once.  This also covers message lengths that are not a multiple of 8
// (crc_byte.c:130-131).
__device__ __forceinline__ bool block_crc_check(u32 bits, uint8_t* sbytes, uint8_t* outp, const CbMeta& m,
                                                const u32* crc_tab, u32* xred) {
  typedef unsigned long long u64;
  const int K = m.K, nb = K >> 3, nw = (nb + 3) >> 2;
  const int ct = m.crc_type;
  const int w = (ct <= 1) ? 24 : (ct == 2 ? 16 : 8);
  const int j_lo = (ct == 0) ? ((m.F >> 3) << 3) : 0;
  const int j_hi = j_lo + K - w - ((ct == 0) ? m.F : 0);       // message = bits [j_lo, j_hi)
  const int Q = j_hi >> 5, r = j_hi & 31;
  u64 acc = 0;
  const int wi = threadIdx.x;
  if (wi < nw) {
    const u32 word = __byte_perm(bits, 0, 0x0123);             // first byte in the low 8 bits
    reinterpret_cast<u32*>(sbytes)[wi] = word;
    const int b0 = wi << 2;
    if (b0 + 3 < nb && ((reinterpret_cast<uintptr_t>(outp) & 3) == 0)) reinterpret_cast<u32*>(outp)[wi] = word;
    else
      for (int q = 0; q < 4 && b0 + q < nb; ++q) outp[b0 + q] = (uint8_t)(word >> (8 * q));
    u32 mb = bits;
    const int lead = j_lo - (wi << 5);                         // positions before the message start
    if (lead > 0) mb = (lead < 32) ? (mb & (0xffffffffu >> lead)) : 0u;
    if (wi < Q) acc = clmul32(mb, __ldg(crc_tab + (ct * 32 + r) * CRC_NM + (Q - wi - 1)));
    else if (wi == Q && r > 0) acc = clmul32(mb >> (32 - r), __ldg(crc_tab + (ct * 32) * CRC_NM));
  }
  u32 alo = (u32)acc, ahi = (u32)(acc >> 32);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { alo ^= __shfl_xor_sync(0xffffffffu, alo, o); ahi ^= __shfl_xor_sync(0xffffffffu, ahi, o); }
  if ((threadIdx.x & 31) == 0) { xred[2 * (threadIdx.x >> 5)] = alo; xred[2 * (threadIdx.x >> 5) + 1] = ahi; }
  __syncthreads();
  bool pass = false;
  if (threadIdx.x == 0) {
    const u32 poly = (ct == 0) ? 0x864cfbu : (ct == 1) ? 0x800063u : (ct == 2) ? 0x1021u : 0x9Bu;
    const u64 pfull = ((u64)1 << w) | poly;
    u64 V = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) V ^= ((u64)xred[2 * i + 1] << 32) | xred[2 * i];
    for (int bit = 31 + w; bit >= w; --bit)
      if ((V >> bit) & 1) V ^= pfull << (bit - w);
    const u32 crc = (u32)V & (u32)(((u64)1 << w) - 1);
    // received CRC as the reference assembles it: big-endian for the 24-bit CRCs (after its
    // byte swap, TD16:1314-1326), a little-endian 16-bit load for CRC16 (:1329-1333)
    u32 rx;
    if (w == 24) rx = ((u32)sbytes[nb - 3] << 16) | ((u32)sbytes[nb - 2] << 8) | sbytes[nb - 1];
    else if (w == 16) rx = ((u32)sbytes[nb - 1] << 8) | sbytes[nb - 2];
    else rx = sbytes[nb - 1];
    pass = (crc == rx && crc != 0);                            // TD16:1348
  }
  return pass;
}

// ------------------------------------------------------------------------------------
#ifndef X2_MIN_CTAS
#define X2_MIN_CTAS 8      // 32 registers -> 8 CTAs/SM (shared-memory limit); measured 6 % faster than 40 registers / 6 CTAs
#endif
template <int TH>
__global__ void __launch_bounds__(TH, X2_MIN_CTAS * (XCHG_THREADS / TH) > 32 ? 32 : X2_MIN_CTAS * (XCHG_THREADS / TH)) k_x2_16(XchgArgs p) {
  extern __shared__ int16_t sm[];
  __shared__ int smax;
  __shared__ u32 xred[2 * XCHG_THREADS / 32];
  __shared__ __align__(16) uint8_t sbytes[768 + 32];
  __shared__ __align__(16) uint8_t snib[1536 + 16];              // hard decisions, one 4-step group per byte
  const int blk = xchg_block(p, (XCHG_LIST_MODE & 2) != 0);
  if (blk < 0) return;
  const CbMeta m = p.meta[blk];
  CbState* st = &p.state[blk];
  if (st->status != 0 || p.iter > m.max_iter) return;
  if (threadIdx.x == 0) smax = 0;
  // max_sys still describes the second decoder's input (written by k_x1_16 of this iteration)
  const int guardB = max(st->max_sys, st->max_in) + st->max_in;
  const int K = m.K, A = p.A;
  int16_t* slot = p.ws + (long)blk * p.slot_hw;
  const uint4* gext2 = reinterpret_cast<const uint4*>(slot + (long)ARR_EXT2 * A);
  const uint4* gext = reinterpret_cast<const uint4*>(slot + (long)ARR_EXT * A);
  const uint4* gs0 = reinterpret_cast<const uint4*>(slot + (long)ARR_S0 * A);
  // (reading s0 from its int8 copy was measured: 14 % SLOWER despite 6 KB less traffic per block)
  uint4* gsys = reinterpret_cast<uint4*>(slot + (long)ARR_SYS * A);
  const int n8 = c4_words(m.W) >> 2;
  int16_t* nat = sm;
  const uint4* T4 = reinterpret_cast<const uint4*>(p.t_pool + m.t_off);
  for (int i = threadIdx.x; i < n8; i += TH) {      // ext2 back to natural order (scatter)
    const int r = (i >> 3) & 3;                                 // rotated step order: conflict-free banks (rot4)
    const uint4 v = rot4(gext2[i], r), tt = rot4(__ldg(T4 + i), r);
    const u32 vw[4] = {v.x, v.y, v.z, v.w}, tw[4] = {tt.x, tt.y, tt.z, tt.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      nat[tw[q] & 0xffffu] = (int16_t)(vw[q] & 0xffffu);
      nat[tw[q] >> 16] = (int16_t)(vw[q] >> 16);
    }
  }
  __syncthreads();
  // When this block satisfies the MAP fast-path guard (B <= guard) its a-posteriori
  // LLRs obey |ext2| <= 12(B+1)+276, so |ext2| + |ext| + |s0| stays inside int16 and the two
  // saturating operations below are plain adds: 4 instead of 14 instructions per word.
  // ... or when the second decoder's pass ran tracked and recorded its largest |ext2| (an exact bound)
  const int mx2 = st->max_ext2;
  const bool nosat = ((guardB <= p.guard_b) && (12 * (guardB + 1) + 276 + st->max_ext + st->max_in <= 32767)) ||
                     (mx2 >= 0 && mx2 + st->max_ext + st->max_in <= 32767);
  MinMax2 mm;
  // hard decisions (iteration_cnt > 1 only, :1267): when 4 | W every uint4 of the natural-order array holds two
  // complete 4-position groups; they are parked one per byte (index n^7, n = position/4) and packed below
  const bool hd = p.iter > 1, hd4 = hd && ((m.W & 3) == 0);
  const int Wq = m.W >> 2;
  auto park = [&](int i, const uint4& d) {
    const u32 nb2 = hd_nibbles(d);
    const int c = i >> 2, l0 = (i & 3) << 1;
    snib[(l0 * Wq + c) ^ 7] = (uint8_t)(nb2 & 0xffu);
    snib[((l0 + 1) * Wq + c) ^ 7] = (uint8_t)(nb2 >> 16);
  };
  if (nosat) {
    for (int i = threadIdx.x; i < n8; i += TH) {     // s1 = (ext2 - ext) + s0, :1241-1265
      const uint4 d = reinterpret_cast<uint4*>(nat)[i], e = gext[i], s0 = gs0[i];
      if (hd4) park(i, d);
      uint4 r;
      r.x = __vadd2(d.x, __vsub2(s0.x, e.x));
      r.y = __vadd2(d.y, __vsub2(s0.y, e.y));
      r.z = __vadd2(d.z, __vsub2(s0.z, e.z));
      r.w = __vadd2(d.w, __vsub2(s0.w, e.w));
      gsys[i] = r;
      mm.add(r);
    }
  } else {
    for (int i = threadIdx.x; i < n8; i += TH) {     // s1 = (ext2 (-) ext) (+) s0
      const uint4 d = reinterpret_cast<uint4*>(nat)[i], e = gext[i], s0 = gs0[i];
      if (hd4) park(i, d);
      uint4 r;
      r.x = __vaddss2(__vsubss2(d.x, e.x), s0.x);
      r.y = __vaddss2(__vsubss2(d.y, e.y), s0.y);
      r.z = __vaddss2(__vsubss2(d.z, e.z), s0.z);
      r.w = __vaddss2(__vsubss2(d.w, e.w), s0.w);
      gsys[i] = r;
      mm.add(r);
    }
  }
  warp_max_to(mm.absmax(), &smax);

  bool pass = false;
  if (hd) {                                                    // :1267-1351
    uint8_t* outp = p.out_base + m.out_off;
    if (hd4) __syncthreads();
    // bit = ext2 > 0 at natural position j, MSB first: thread w packs positions 32w..32w+31
    u32 word = 0;
    if ((int)threadIdx.x < ((K >> 3) + 3) >> 2)
      word = hd4 ? nibbles_to_word(reinterpret_cast<const uint2*>(snib)[threadIdx.x]) : hd_word32(nat, threadIdx.x << 5, K, m.W);
    pass = block_crc_check(word, sbytes, outp, m, p.crc_xp, xred);
  } else {
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    st->max_sys = smax;
    st->max_ext2 = -1;                                       // consumed: valid for the pass that wrote it only
    int s = 0;
    if (pass) s = p.iter;
    else if (p.iter >= m.max_iter) s = m.max_iter + 1;
    if (s) {
      st->status = s;
      if (p.status_out) p.status_out[blk] = (uint8_t)s;
    }
  }
}

}  // namespace oai
