// 8-bit max-log-MAP turbo decoder (reference: openair1/PHY/CODING/3gpplte_turbo_decoder_sse_8bit.c).
// int8 arithmetic saturates routinely (metrics live in [-128,127]), so there is no non-saturating
// fast path as in the 16-bit kernel; every add/sub is clamped like _mm_adds_epi8/_mm_subs_epi8.
//
// k_map8 mapping: 8 threads per code block, thread t = the reference's SIMD lanes 2t, 2t+1 (lane l covers
// trellis positions [l*W,(l+1)*W), W = n/16, TD8:874-881) packed in the halfwords of a register, the 8 states
// in 8 registers; 4 blocks per warp.  Metrics are kept in OFFSET form (value + 128, range 0..255) in int16
// halfwords, so that one saturating int8 operation is one DPX instruction with the relu flag:
//     sat8(x + g) + 128 = max(min(x' + g, 255), 0)                       (__viaddmin_s16x2_relu)
// and, clamping being monotone, an add-compare-select max(sat8(x+g), sat8(y-g)) is
//     clamp(max(x' + g, y' - g), 0, 255) = VIADD + VIADDMNMX + VIMNMX.relu.
// Memory: like the 16-bit kernel, alpha is checkpointed every 8 steps in HBM during the forward sweep and
// recomputed per 8-step segment into shared memory in the backward sweep; beta is never stored.  (The
// reference stores both arrays for every step, TD8:928-929.)
//
// Data layout of the int8 per-position arrays ("C8"): chunk = 8 steps = 128 bytes,
//     byte(k, l) = (k>>3)*128 + (l>>1)*16 + (k&7)*2 + (l&1)
// so thread t reads the 8 steps of its two lanes with one LDG.128.
//
//   k_demux8 : input scaling int16 -> int8 (TD8:1000-1029) and demux (TD8:1062-1077)
//   k_map8   : log_map8 = gamma/alpha/beta/ext (TD8:95-149, 151-827)
//   k_x1_8   : feedback ext = (ext (-) s1) (+) s0 (TD8:1632-1653) and gather s2 = ext o pi (TD8:1341-1379)
//   k_x2_8   : s1 = (ext2 o pi^-1 (-) ext) (+) s0, hard decision (two rules, TD8:1392-1581), CRC, exit
// Parity domain n >= 256, n % 16 == 0; the tail LLRs never influence the
// reference's output and are not read.
#pragma once
#include <type_traits>
#include "td_common.cuh"
#include "td16_map.cuh"
#include "td16_xchg.cuh"

namespace oai {

constexpr int MAP8_THREADS = 64;      // 8 code blocks per CTA
constexpr int MAP8_SMEM_BYTES = 8 * MAP8_THREADS * 32;   // alpha of the 8 steps of a segment, 32 B per thread and step
constexpr int INIT8 = -63;            // -MAX8/2 (TD8:92,234)
constexpr int RERUN8 = 16;            // L (TD8:211)

enum { A8_S0 = 0, A8_P1 = 1, A8_P2 = 2, A8_SYS = 3, A8_EXT = 4, A8_EXT2 = 5, A8_COUNT = 6 };

__host__ __device__ inline int h8(int k, int lane) { return ((k >> 3) << 7) + ((lane >> 1) << 4) + ((k & 7) << 1) + (lane & 1); }
__host__ __device__ inline int c8_bytes(int W) { return ((W + 7) >> 3) << 7; }          // bytes per array actually used
// trellis position -> byte index (lane = pos / W, step = pos % W)
__device__ __forceinline__ int st8(int pos, int W) { const int lane = pos / W; return h8(pos - lane * W, lane); }
// checkpoint pool per block: one 256-byte slot (8 threads x 32 B) per 8-step segment + 4 special slots
__host__ __device__ inline long ckpt8_words(int W) { return (long)(((W + 7) >> 3) + 4) * 64; }

struct Td8Args {
  const CbMeta* meta;
  CbState* state;
  int8_t* ws;            // per block: A8_COUNT arrays of `A` bytes
  long slot_b;
  int A;                 // bytes per array (>= n, multiple of 16)
  u32* ck;               // alpha checkpoint pool
  long ck_words;         // words per block in `ck`
  int nblk;
  const uint16_t* qpp;   // plain QPP tables pi[i]
  const uint16_t* t8;    // per K: T8[h] = C8 byte index of the QPP image of the position stored at byte h (padding: itself)
  const u32* crc_xp;
  const int16_t* in_base;
  uint8_t* out_base;
  uint8_t* status_out;
  int iter;
  int sys_arr, par_arr, out_arr;
};

__device__ __forceinline__ int s8(int v) { return max(-128, min(127, v)); }

// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(XCHG_THREADS) k_demux8(Td8Args p) {
  extern __shared__ int8_t sm8[];
  __shared__ int red[XCHG_THREADS / 32];
  const int blk = blockIdx.x;
  if (blk >= p.nblk) return;
  const CbMeta m = p.meta[blk];
  CbState* st = &p.state[blk];
  if (!(m.flags & 1)) { if (threadIdx.x == 0) st->status = 0xFE; return; }
  const int n = m.K, W = n >> 4, A = p.A;
  const int16_t* y = p.in_base + (((long)m.in_off_hi << 32) | m.in_off_lo);
  // round_avg (TD8:1001-1008): over the first 3*(n>>4)+1 vectors of 8, |y0..y3| + 2|y4| + 2|y5|,
  // with _mm_abs_epi16 leaving -32768 negative, summed in 32-bit lanes (wrapping)
  unsigned sum = 0;
  const int nvec = 3 * (n >> 4) + 1;
  for (int i = threadIdx.x; i < nvec * 8; i += XCHG_THREADS) {
    const int e = i & 7, v = y[i];
    const int a = (v == -32768) ? -32768 : abs(v);
    if (e < 4) sum += (unsigned)a;
    else if (e < 6) sum += 2u * (unsigned)a;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = (int)sum;
  __syncthreads();
  unsigned tot = 0;
  for (int i = 0; i < XCHG_THREADS / 32; ++i) tot += (unsigned)red[i];
  const int round_avg = (int)tot / (n * 3);
  const int bracket = round_avg < 16 ? 0 : (round_avg < 32 ? 1 : (round_avg < 64 ? 2 : (round_avg < 128 ? 3 : 4)));
  int8_t* s0 = sm8, *p1 = sm8 + A, *p2 = sm8 + 2 * A;
  for (int pos = threadIdx.x; pos < n; pos += XCHG_THREADS) {
    const int h = st8(pos, W);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int i = 3 * pos + c;
      const int sh = bracket < 4 ? bracket : (((i >> 3) & 1) ? 4 : 3);     // TD8:1027-1029
      const int v = s8((int)y[i] >> sh);                                    // _mm_srai_epi16 + _mm_packs_epi16
      (c == 0 ? s0 : (c == 1 ? p1 : p2))[h] = (int8_t)v;
    }
  }
  __syncthreads();
  int8_t* slot = p.ws + (long)blk * p.slot_b;
  for (int i = threadIdx.x; i < c8_bytes(W) / 16; i += XCHG_THREADS) {
    reinterpret_cast<uint4*>(slot + (long)A8_S0 * A)[i] = reinterpret_cast<uint4*>(s0)[i];
    reinterpret_cast<uint4*>(slot + (long)A8_P1 * A)[i] = reinterpret_cast<uint4*>(p1)[i];
    reinterpret_cast<uint4*>(slot + (long)A8_P2 * A)[i] = reinterpret_cast<uint4*>(p2)[i];
  }
  if (threadIdx.x == 0) {
    st->status = (m.max_iter == 0) ? 1 : 0;
    if (m.max_iter == 0 && p.status_out) p.status_out[blk] = 1;
  }
}

// ------------------------------------------------------------------------------------
// packed (2 lanes) offset-form arithmetic
constexpr u32 K255 = 0x00ff00ffu, K128 = 0x00800080u;
struct G8 { u32 g1, g0, n1, n0; };      // m11, m10 and their negatives (signed halfwords)
// per-halfword arithmetic >>1 of a value in [-256, 255]: bias to non-negative, logical shift, un-bias (2 ALU-pipe
// instructions instead of 3; the adds go to the other pipe)
__device__ __forceinline__ u32 vsra1_9bit(u32 x) {
  return __vsub2(((__vadd2(x, 0x01000100u)) >> 1) & 0x7fff7fffu, K128);
}
__device__ __forceinline__ G8 gamma8(u32 s, u32 p) {   // TD8:178-185: widen, add/sub, >>1 (exact floor halves, no saturation)
  G8 g;
  g.g1 = vsra1_9bit(__vadd2(s, p));
  g.g0 = vsra1_9bit(__vsub2(s, p));
  g.n1 = __vadd2(~g.g1, 0x00010001u);
  g.n0 = __vadd2(~g.g0, 0x00010001u);
  return g;
}
// max(sat8(x+gx), sat8(y+gy)) in offset form = clamp(max(x' + gx, y' + gy), 0, 255).  Every metric that enters a
// recursion step is max-normalised or a start value (<= 0, offset form <= 128), so a sum can exceed 255 only by one,
// when a branch metric is +128 = -(-128), i.e. when m11 or m10 of the step is -128 (see chunk_pair_hazard).
// SAFE = false (no such step in the segment, checked by the caller): only the lower clamp can bind and the
// add-compare-select is VIADD + VIADDMNMX.RELU; SAFE = true: + VIMNMX.RELU against 255.
template <bool SAFE>
__device__ __forceinline__ u32 acs8(u32 x, u32 gx, u32 y, u32 gy) {
  if (SAFE) return __vimin_s16x2_relu(__viaddmax_s16x2(x, gx, __vadd2(y, gy)), K255);
  return __viaddmax_s16x2_relu(x, gx, __vadd2(y, gy));
}
// m11 = (s+p)>>1 or m10 = (s-p)>>1 equals -128 only for (s,p) in {(-128,-128), (-128,-127), (-127,-128), (-128,127)}.
// Conservative chunk-level test (8 steps x 2 lanes each): some systematic byte is <= -127 AND some parity byte is
// <= -127 or == 127.  (Systematic inputs saturate routinely once the decoder converges; channel parity rarely does.)
__device__ __forceinline__ u32 zero_byte_mask(u32 x) { return (x - 0x01010101u) & ~x & 0x80808080u; }
__device__ __forceinline__ bool chunk_pair_hazard(const uint4& sv, const uint4& pv) {
  auto lo = [](u32 x) -> u32 { return zero_byte_mask((x ^ 0x80808080u) & 0xfefefefeu); };       // byte is 0x80 or 0x81
  auto hi = [](u32 x) -> u32 { return zero_byte_mask(x ^ 0x7f7f7f7fu); };                       // byte is 0x7f
  const u32 hs = lo(sv.x) | lo(sv.y) | lo(sv.z) | lo(sv.w);
  if (hs == 0) return false;
  return (lo(pv.x) | lo(pv.y) | lo(pv.z) | lo(pv.w) | hi(pv.x) | hi(pv.y) | hi(pv.z) | hi(pv.w)) != 0;
}
// out = sat8(n - max_s n) in offset form: max(n' - mx' + 128, 0); n' + c <= 128, so the min with 255 of the
// add-min-relu form never binds and the zero needs no register
__device__ __forceinline__ void norm8(u32 (&v)[8], const u32 (&n)[8]) {
  const u32 mx = __vmaxs2(__vimax3_s16x2(n[0], n[1], n[2]), __vimax3_s16x2(n[3], n[4], __vimax3_s16x2(n[5], n[6], n[7])));
  const u32 c = __vadd2(~mx, 0x00810081u);              // 128 - mx per halfword
#pragma unroll
  for (int s = 0; s < 8; ++s) v[s] = __viaddmin_s16x2_relu(n[s], c, K255);
}
// forward recursion (TD8:251-297)
template <bool SAFE = true>
__device__ __forceinline__ void alpha8_step(u32 (&a)[8], const G8& g) {
  u32 n[8];
  n[0] = acs8<SAFE>(a[1], g.g1, a[0], g.n1);
  n[1] = acs8<SAFE>(a[3], g.n0, a[2], g.g0);
  n[2] = acs8<SAFE>(a[5], g.g0, a[4], g.n0);
  n[3] = acs8<SAFE>(a[7], g.n1, a[6], g.g1);
  n[4] = acs8<SAFE>(a[1], g.n1, a[0], g.g1);
  n[5] = acs8<SAFE>(a[3], g.g0, a[2], g.n0);
  n[6] = acs8<SAFE>(a[5], g.n0, a[4], g.g0);
  n[7] = acs8<SAFE>(a[7], g.g1, a[6], g.n1);
  norm8(a, n);
}
// backward recursion (TD8:579-650)
template <bool SAFE = true>
__device__ __forceinline__ void beta8_step(u32 (&b)[8], const G8& g) {
  u32 n[8];
  n[0] = acs8<SAFE>(b[4], g.g1, b[0], g.n1);
  n[1] = acs8<SAFE>(b[4], g.n1, b[0], g.g1);
  n[2] = acs8<SAFE>(b[5], g.n0, b[1], g.g0);
  n[3] = acs8<SAFE>(b[5], g.g0, b[1], g.n0);
  n[4] = acs8<SAFE>(b[6], g.g0, b[2], g.n0);
  n[5] = acs8<SAFE>(b[6], g.n0, b[2], g.g0);
  n[6] = acs8<SAFE>(b[7], g.n1, b[3], g.g1);
  n[7] = acs8<SAFE>(b[7], g.g1, b[3], g.n1);
  norm8(b, n);
}
// max of four saturated sums a_i (+) b_j, offset form (a signed = offset - 128, b offset)
__device__ __forceinline__ u32 max4sum8(u32 a0, u32 b0, u32 a1, u32 b1, u32 a2, u32 b2, u32 a3, u32 b3) {
  u32 x = __vadd2(a0, b0);
  x = __viaddmax_s16x2(a1, b1, x);
  x = __viaddmax_s16x2(a2, b2, x);
  return __viaddmax_s16x2_relu(a3, b3, x);       // a <= 0, b' <= 128: only the lower clamp can bind
}
// a-posteriori LLR of one step (TD8:715-770); a, b in offset form; returns SIGNED int8-range halfwords
__device__ __forceinline__ u32 ext8_step(const u32 (&ao)[8], const u32 (&b)[8], const G8& g) {
  u32 a[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) a[s] = __vsub2(ao[s], K128);
  u32 m00 = max4sum8(a[0], b[0], a[1], b[4], a[6], b[7], a[7], b[3]);
  u32 m11 = max4sum8(a[0], b[4], a[1], b[0], a[6], b[3], a[7], b[7]);
  u32 m01 = max4sum8(a[2], b[5], a[3], b[1], a[4], b[2], a[5], b[6]);
  u32 m10 = max4sum8(a[2], b[1], a[3], b[5], a[4], b[6], a[5], b[2]);
  m01 = __viaddmin_s16x2_relu(m01, g.n0, K255);
  m00 = __viaddmin_s16x2_relu(m00, g.n1, K255);
  m10 = __viaddmin_s16x2_relu(m10, g.g0, K255);
  m11 = __viaddmin_s16x2_relu(m11, g.g1, K255);
  const u32 d = __vsub2(__vmaxs2(m10, m11), __vmaxs2(m01, m00));        // in [-255, 255]
  return __vsub2(__viaddmin_s16x2_relu(d, K128, K255), K128);           // sat8
}

// step e (0..7) of a chunk register: two int8 -> packed sign-extended int16 pair
__device__ __forceinline__ u32 unp8(const uint4& v, int e) {
  const u32 w = (e >> 1) == 0 ? v.x : ((e >> 1) == 1 ? v.y : ((e >> 1) == 2 ? v.z : v.w));
  return prmt_sx(w, (e & 1) ? 0xB3A2u : 0x9180u);
}

#ifndef MAP8_MAX_REGS
#define MAP8_MAX_REGS 128      // measured: 96 / 112 / 128 / 140 -> 27.3 / 27.6 / 25.3 / 25.9 ms per 3 decodes of 21312 blocks
#endif
__global__ void __maxnreg__(MAP8_MAX_REGS) k_map8(Td8Args p) {
  extern __shared__ uint4 abuf8[];
  const int tid = threadIdx.x;
  const int gt = blockIdx.x * MAP8_THREADS + tid;
  const int blk = gt >> 3, t = gt & 7;
  const unsigned gmask = 0xFFu << ((tid & 31) & ~7);            // the 8 threads of this block
  if (blk >= p.nblk) return;
  const CbMeta m = p.meta[blk];
  const CbState* st = &p.state[blk];
  if (st->status != 0 || !(m.flags & 1) || p.iter > m.max_iter) return;
  const int n = m.K, W = n >> 4, nseg = (W + 7) >> 3;
  int8_t* slot = p.ws + (long)blk * p.slot_b;
  const uint4* sys4 = reinterpret_cast<const uint4*>(slot + (long)p.sys_arr * p.A) + t;      // chunk c at [c*8]
  const uint4* par4 = reinterpret_cast<const uint4*>(slot + (long)p.par_arr * p.A) + t;
  int8_t* ext = slot + (long)p.out_arr * p.A + t * 16;                                       // chunk c at + c*128
  u32* ck = p.ck + (long)blk * p.ck_words + t * 8;                                           // slot i at + i*64
  const int CH0 = nseg, CH8 = nseg + 1, CH16 = nseg + 2, A0 = nseg + 3;
  auto put = [&](int e, const u32 (&a)[8]) {
    abuf8[(2 * e) * MAP8_THREADS + tid] = make_uint4(a[0], a[1], a[2], a[3]);
    abuf8[(2 * e + 1) * MAP8_THREADS + tid] = make_uint4(a[4], a[5], a[6], a[7]);
  };
  auto get = [&](int e, u32 (&a)[8]) {
    const uint4 x = abuf8[(2 * e) * MAP8_THREADS + tid], y = abuf8[(2 * e + 1) * MAP8_THREADS + tid];
    a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w; a[4] = y.x; a[5] = y.y; a[6] = y.z; a[7] = y.w;
  };
  auto ckput = [&](int i, const u32 (&a)[8]) { ckpt_put(ck + i * 64, a); };
  auto ckget = [&](int i, u32 (&a)[8]) { ckpt_get(ck + i * 64, a); };
  const u32 kInit = pack2(INIT8 + 128, INIT8 + 128);
  u32 a[8], b[8];

  // ---- alpha pass 1 (TD8:234-297): W steps from (0,-63..) in lane 0 and -63 elsewhere; checkpoint per segment ----
#pragma unroll
  for (int s = 0; s < 8; ++s) a[s] = kInit;
  if (t == 0) a[0] = pack2(128, INIT8 + 128);
  {
    uint4 S = __ldg(sys4), P = __ldg(par4);
    for (int seg = 0; seg < nseg; ++seg) {
      ckput(seg, a);
      const uint4 Sc = S, Pc = P;
      if (seg + 1 < nseg) { S = __ldg(sys4 + (seg + 1) * 8); P = __ldg(par4 + (seg + 1) * 8); }
      if (seg * 8 + 8 <= W) {
        if (!chunk_pair_hazard(Sc, Pc)) {
#pragma unroll
          for (int e = 0; e < 8; ++e) alpha8_step<false>(a, gamma8(unp8(Sc, e), unp8(Pc, e)));
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) alpha8_step<true>(a, gamma8(unp8(Sc, e), unp8(Pc, e)));
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (seg * 8 + e < W) alpha8_step(a, gamma8(unp8(Sc, e), unp8(Pc, e)));
      }
    }
  }
  // ---- re-seed (TD8:299-316): lane l <- alpha[W] of lane l-1, lane 0 <- (0,-63..); 16-step re-run ----
  auto shift_up = [&](u32 (&dst)[8], const u32 (&src)[8]) {
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      u32 prev = __shfl_sync(gmask, src[s], (t + 7) & 7, 8);
      if (t == 0) prev = pack2(0, (s == 0) ? 128 : INIT8 + 128);     // hi half is what gets used
      dst[s] = __byte_perm(prev, src[s], 0x5432);                     // lo <- prev.hi, hi <- mine.lo
    }
  };
  {
    u32 c[8];
    shift_up(c, a);
    ckput(CH0, c);
    const uint4 S0 = __ldg(sys4), P0 = __ldg(par4), S1 = __ldg(sys4 + 8), P1 = __ldg(par4 + 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) alpha8_step(c, gamma8(unp8(S0, e), unp8(P0, e)));
    ckput(CH8, c);
#pragma unroll
    for (int e = 0; e < 8; ++e) alpha8_step(c, gamma8(unp8(S1, e), unp8(P1, e)));
    ckput(CH16, c);
    if (W == RERUN8) {                          // the re-run reached alpha[W] (K = 256): it is the new final vector
#pragma unroll
      for (int s = 0; s < 8; ++s) a[s] = c[s];
    }
    shift_up(c, a);                             // second re-seed: the final alpha[0]
    ckput(A0, c);
  }

  // alpha[k0 .. k1) of segment `seg` (final values) -> shared memory entries 0..; leaves the chunk inputs in S, P
  uint4 S, P;
  // (S, P must hold the segment's chunk registers; safe = std::false_type when the chunk has no extreme systematic value)
  auto fill_alpha = [&](auto full, auto safe, int seg) {
    constexpr bool FULL = decltype(full)::value;                // all 8 steps of the segment exist
    constexpr bool SAFE = decltype(safe)::value;
    const int k0 = seg * 8;
    u32 x[8];
    ckget(seg == 0 ? CH0 : (seg == 1 ? CH8 : seg), x);          // steps 1..16 come from the re-run chain
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (FULL || k0 + e < W) {
        put(e, x);
        if (e < 7) alpha8_step<SAFE>(x, gamma8(unp8(S, e), unp8(P, e)));
      }
    }
    if (seg == 0) { ckget(A0, x); put(0, x); }                  // alpha[0]: the value of the second re-seed
    if (seg == 2) { ckget(CH16, x); put(0, x); }                // alpha[16]: last value of the chain
  };
  // backward over one segment with beta[k1] in b: ext for k in [elo, ehi], beta steps for k >= blo.
  // full = the segment has 8 steps, all of them in the ext range and all stepping beta (no per-step tests).
  auto back_segment_impl = [&](auto full, auto safe, int seg, int elo, int ehi, int blo) {
    constexpr bool FULL = decltype(full)::value;
    constexpr bool SAFE = decltype(safe)::value;
    const int k0 = seg * 8;
    fill_alpha(full, safe, seg);
    u32 o[8];
#pragma unroll
    for (int e = 7; e >= 0; --e) {
      o[e] = 0;
      const int k = k0 + e;
      if (FULL || k < W) {
        const G8 g = gamma8(unp8(S, e), unp8(P, e));
        if (FULL || (k >= elo && k <= ehi)) {
          u32 x[8];
          get(e, x);
          o[e] = ext8_step(x, b, g);
        }
        if (FULL || k >= blo) beta8_step<SAFE>(b, g);
      }
    }
    int8_t* dst = ext + seg * 128;
    if (FULL && k0 + 7 > ehi) {        // full segment that reaches into the re-run's range: only the LLRs up to ehi are this sweep's
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (k0 + e <= ehi) *reinterpret_cast<uint16_t*>(dst + 2 * e) = (uint16_t)__byte_perm(o[e], 0, 0x4420);
    } else if (FULL || (k0 >= elo && k0 + 7 <= ehi && k0 + 7 < W)) {   // whole segment: one 16-byte store
      *reinterpret_cast<uint4*>(dst) = make_uint4(__byte_perm(o[0], o[1], 0x6420), __byte_perm(o[2], o[3], 0x6420),
                                                  __byte_perm(o[4], o[5], 0x6420), __byte_perm(o[6], o[7], 0x6420));
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = k0 + e;
        if (k < W && k >= elo && k <= ehi) *reinterpret_cast<uint16_t*>(dst + 2 * e) = (uint16_t)__byte_perm(o[e], 0, 0x4420);
      }
    }
  };
  const std::true_type kFull{};
  const std::false_type kPart{};
  auto back_segment = [&](auto full, int seg, int elo, int ehi, int blo) {
    S = __ldg(sys4 + seg * 8); P = __ldg(par4 + seg * 8);
    if (decltype(full)::value && !chunk_pair_hazard(S, P)) back_segment_impl(full, std::false_type{}, seg, elo, ehi, blo);
    else back_segment_impl(full, std::true_type{}, seg, elo, ehi, blo);
  };

  // ---- beta pass 1 (TD8:505-650): from alpha[W], lane 15 <- 0; all W steps; ext where beta[k+1] is final ----
#pragma unroll
  for (int s = 0; s < 8; ++s) b[s] = (t == 7) ? ((a[s] & 0xffffu) | (128u << 16)) : a[s];
  for (int seg = nseg - 1; seg >= 0; --seg) {
    // (a full segment inside the last 18 steps takes the test-free code as well: its beta steps are the same, the LLRs it
    // computes beyond W - 18 are simply not stored -- 13.6 instead of 12.9 Gbit/s at K = 6144)
    if (seg * 8 + 8 <= W) back_segment(kFull, seg, 0, W - RERUN8 - 2, 0);
    else back_segment(kPart, seg, 0, W - RERUN8 - 2, 0);
  }
  // ---- re-seed (TD8:652-666): lane l <- beta[0] of lane l+1, lane 15 <- 0; re-run of the last 16 steps ----
  auto shift_down = [&](u32 (&dst)[8], const u32 (&src)[8]) {
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      u32 next = __shfl_sync(gmask, src[s], (t + 1) & 7, 8);
      if (t == 7) next = 128u;                                        // lo half is what gets used
      dst[s] = __byte_perm(src[s], next, 0x5432);                     // lo <- mine.hi, hi <- next.lo
    }
  };
  u32 bw[8];                                    // beta[W] after the first re-seed
  shift_down(bw, b);
#pragma unroll
  for (int s = 0; s < 8; ++s) b[s] = bw[s];
  const int klo = max(W - RERUN8 - 1, 0);       // ext(W-2 .. W-17) uses the re-run's beta[W-1 .. W-16]
  for (int seg = nseg - 1; seg >= 0 && seg * 8 + 7 >= klo; --seg) back_segment(kPart, seg, klo, W - 2, W - RERUN8);
  // ---- ext(W-1) uses beta[W] of the SECOND re-seed: the same vector unless the re-run reached step 0 (K = 256) ----
  if (W == RERUN8) shift_down(bw, b);
#pragma unroll
  for (int s = 0; s < 8; ++s) b[s] = bw[s];
  back_segment(kPart, nseg - 1, W - 1, W - 1, W);
}

// ------------------------------------------------------------------------------------
// Exchange kernels: one CTA per block, arrays staged in shared memory, 128-bit HBM accesses, the QPP
// permutation as a byte gather / scatter through the T8 table (CbMeta::t_off).
__global__ void __launch_bounds__(XCHG_THREADS) k_x1_8(Td8Args p) {
  extern __shared__ int8_t sm8[];
  const int blk = blockIdx.x;
  if (blk >= p.nblk) return;
  const CbMeta m = p.meta[blk];
  if (p.state[blk].status != 0 || p.iter > m.max_iter) return;
  const int n = m.K, W = n >> 4, A = p.A;
  int8_t* slot = p.ws + (long)blk * p.slot_b;
  uint4* gext = reinterpret_cast<uint4*>(slot + (long)A8_EXT * A);
  const uint4* gsys4 = reinterpret_cast<const uint4*>(slot + (long)A8_SYS * A);
  const uint4* gs0 = reinterpret_cast<const uint4*>(slot + (long)A8_S0 * A);
  const int n16 = c8_bytes(W) >> 4;     // elementwise work covers the padded extent (holes of a partial last chunk are never used)
  for (int i = threadIdx.x; i < n16; i += XCHG_THREADS) {
    uint4 e = gext[i];
    if (p.iter > 1) {                   // ext = (ext (-) s1) (+) s0, TD8:1632-1653
      const uint4 s1 = gsys4[i], z = gs0[i];
      e.x = __vaddss4(__vsubss4(e.x, s1.x), z.x); e.y = __vaddss4(__vsubss4(e.y, s1.y), z.y);
      e.z = __vaddss4(__vsubss4(e.z, s1.z), z.z); e.w = __vaddss4(__vsubss4(e.w, s1.w), z.w);
      gext[i] = e;
    }
    reinterpret_cast<uint4*>(sm8)[i] = e;
  }
  __syncthreads();
  const uint4* T4 = reinterpret_cast<const uint4*>(p.t8 + m.t_off);
  uint2* gsys = reinterpret_cast<uint2*>(slot + (long)A8_SYS * A);
  const uint8_t* in = reinterpret_cast<const uint8_t*>(sm8);
  for (int i = threadIdx.x; i < 2 * n16; i += XCHG_THREADS) {   // s2[st8(i)] = ext[st8(pi(i))], TD8:1341-1379; 8 bytes per thread
    const uint4 tt = __ldg(T4 + i);
    uint2 o;
    o.x = (u32)in[tt.x & 0xffffu] | ((u32)in[tt.x >> 16] << 8) | ((u32)in[tt.y & 0xffffu] << 16) | ((u32)in[tt.y >> 16] << 24);
    o.y = (u32)in[tt.z & 0xffffu] | ((u32)in[tt.z >> 16] << 8) | ((u32)in[tt.w & 0xffffu] << 16) | ((u32)in[tt.w >> 16] << 24);
    gsys[i] = o;
  }
}

// the four 4-position hard-decision groups of one C8 uint4 (8 steps x lanes 2t, 2t+1 of int8): byte > 0;
// returns lane 2t steps 0-3 in bits 0..3, steps 4-7 in bits 4..7, lane 2t+1 in bits 8..15 (first step = bit 3 of its nibble)
__device__ __forceinline__ u32 hd_nibbles8(const uint4& d) {
  const u32 m0 = __vcmpgts4(d.x, 0u) & 0x01010101u, m1 = __vcmpgts4(d.y, 0u) & 0x01010101u;
  const u32 m2 = __vcmpgts4(d.z, 0u) & 0x01010101u, m3 = __vcmpgts4(d.w, 0u) & 0x01010101u;
  // word q holds steps 2q (bytes 0,1 = lanes 2t, 2t+1) and 2q+1 (bytes 2,3)
  auto lane_bits = [](u32 m, int sh) -> u32 { const u32 x = m >> sh; return ((x & 1u) << 1) | ((x >> 16) & 1u); };   // (step 2q, step 2q+1)
  const u32 l0 = (lane_bits(m0, 0) << 2) | lane_bits(m1, 0) | (((lane_bits(m2, 0) << 2) | lane_bits(m3, 0)) << 4);
  const u32 l1 = (lane_bits(m0, 8) << 2) | lane_bits(m1, 8) | (((lane_bits(m2, 8) << 2) | lane_bits(m3, 8)) << 4);
  return l0 | (l1 << 8);
}

__global__ void __launch_bounds__(XCHG_THREADS, 8) k_x2_8(Td8Args p) {
  extern __shared__ int8_t sm8[];
  __shared__ u32 xred[2 * XCHG_THREADS / 32];
  __shared__ __align__(16) uint8_t sbytes[768 + 32];
  __shared__ __align__(16) uint8_t snib[1536 + 16];
  const int blk = blockIdx.x;
  if (blk >= p.nblk) return;
  const CbMeta m = p.meta[blk];
  CbState* st = &p.state[blk];
  if (st->status != 0 || p.iter > m.max_iter) return;
  const int n = m.K, W = n >> 4, A = p.A;
  int8_t* slot = p.ws + (long)blk * p.slot_b;
  const uint2* gext2 = reinterpret_cast<const uint2*>(slot + (long)A8_EXT2 * A);
  const uint2* gsys2 = reinterpret_cast<const uint2*>(slot + (long)A8_SYS * A);
  const uint4* gext = reinterpret_cast<const uint4*>(slot + (long)A8_EXT * A);
  const uint4* gs0 = reinterpret_cast<const uint4*>(slot + (long)A8_S0 * A);
  uint4* gsys = reinterpret_cast<uint4*>(slot + (long)A8_SYS * A);
  int8_t* nat = sm8, *natdec = sm8 + A;       // ext2 and the decision variable, both in natural order (C8 layout)
  const bool mode1 = (n & 0x7f) == 0;         // TD8:1392 / 1488: decide on ext2, else on ext2 (+) sys2 (TD8:1456)
  const bool hd = p.iter > 1;
  const int n16 = c8_bytes(W) >> 4;
  const uint4* T4 = reinterpret_cast<const uint4*>(p.t8 + m.t_off);
  for (int i = threadIdx.x; i < 2 * n16; i += XCHG_THREADS) {   // scatter to natural order, 8 bytes per thread
    const uint2 v = gext2[i];
    const uint4 tt = __ldg(T4 + i);
    const u32 tw[4] = {tt.x, tt.y, tt.z, tt.w};
    const u32 vw[2] = {v.x, v.y};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      nat[tw[q] & 0xffffu] = (int8_t)(vw[q >> 1] >> (16 * (q & 1)));
      nat[tw[q] >> 16] = (int8_t)(vw[q >> 1] >> (16 * (q & 1) + 8));
    }
    if (hd && !mode1) {
      const uint2 s2 = gsys2[i];
      const u32 dw[2] = {__vaddss4(v.x, s2.x), __vaddss4(v.y, s2.y)};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        natdec[tw[q] & 0xffffu] = (int8_t)(dw[q >> 1] >> (16 * (q & 1)));
        natdec[tw[q] >> 16] = (int8_t)(dw[q >> 1] >> (16 * (q & 1) + 8));
      }
    }
  }
  __syncthreads();
  const int8_t* decv = mode1 ? nat : natdec;
  const bool hd4 = hd && ((W & 3) == 0);
  const int Wq = W >> 2;
  for (int i = threadIdx.x; i < n16; i += XCHG_THREADS) {       // s1 = (ext2 o pi^-1 (-) ext) (+) s0, TD8:1392-1459
    const uint4 d = reinterpret_cast<const uint4*>(nat)[i], e = gext[i], z = gs0[i];
    uint4 r;
    r.x = __vaddss4(__vsubss4(d.x, e.x), z.x); r.y = __vaddss4(__vsubss4(d.y, e.y), z.y);
    r.z = __vaddss4(__vsubss4(d.z, e.z), z.z); r.w = __vaddss4(__vsubss4(d.w, e.w), z.w);
    gsys[i] = r;
    if (hd4) {                                                  // park the four 4-position decision groups of this uint4
      const u32 nb4 = hd_nibbles8(mode1 ? d : reinterpret_cast<const uint4*>(natdec)[i]);
      const int c = i >> 3, l0 = (i & 7) << 1;
      const int n0 = l0 * Wq + 2 * c, n1 = (l0 + 1) * Wq + 2 * c;
      snib[n0 ^ 7] = (uint8_t)(nb4 & 15u);
      snib[n1 ^ 7] = (uint8_t)((nb4 >> 8) & 15u);
      if (8 * c + 4 < W) {                                      // steps 4..7 of a partial last chunk do not exist
        snib[(n0 + 1) ^ 7] = (uint8_t)((nb4 >> 4) & 15u);
        snib[(n1 + 1) ^ 7] = (uint8_t)((nb4 >> 12) & 15u);
      }
    }
  }
  bool pass = false;
  if (hd) {
    if (hd4) __syncthreads();
    u32 word = 0;                              // thread w packs natural positions 32w..32w+31, first position in bit 31
    if ((int)threadIdx.x < ((n >> 3) + 3) >> 2) {
      if (hd4) word = nibbles_to_word(reinterpret_cast<const uint2*>(snib)[threadIdx.x]);
      else {
        const int j0 = threadIdx.x << 5;
        int lane = j0 / W, k = j0 - lane * W;
        for (int q = 0; q < 32; ++q) {
          word = (word << 1) | ((j0 + q < n && decv[h8(k, lane)] > 0) ? 1u : 0u);
          if (++k >= W) { k = 0; ++lane; }
        }
      }
    }
    pass = block_crc_check(word, sbytes, p.out_base + m.out_off, m, p.crc_xp, xred);
  }
  if (threadIdx.x == 0) {
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
