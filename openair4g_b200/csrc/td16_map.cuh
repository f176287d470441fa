// 16-bit max-log-MAP pass (one constituent decoder run) for a batch of code blocks.
//
// Replaces log_map16 = compute_gamma16 + compute_alpha16 + compute_beta16 + compute_ext16
// (reference: openair1/PHY/CODING/3gpplte_turbo_decoder_sse_16bit.c:84-119, 121-169,
// 173-439, 442-693, 695-879) with one kernel.
//
// Mapping: 4 threads per code block; thread t owns the reference's SIMD lanes 2t and
// 2t+1 packed in the two halfwords of a register, and all 8 trellis states in 8
// registers, so one trellis step needs no cross-thread traffic.  A warp carries 8
// blocks, a 128-thread CTA 32 blocks.
//
// Memory: the reference stores alpha for every step (16 B x K per block).  Here the
// forward sweep keeps alpha in registers and writes a checkpoint every S steps; the
// backward sweep recomputes alpha for one S-step segment into shared memory, then runs
// beta and the extrinsic computation over that segment.  Recomputation repeats the
// identical operations on identical inputs, so every alpha value is bit-identical to a
// stored one.  beta is never stored.
//
// The reference's boundary heuristic is reproduced exactly:
//   alpha: full pass from fixed start metrics, then a 5-step re-run (L>>3, :171,189)
//          seeded with the previous lane's final metrics (:232-259) that replaces
//          alpha[1..5];
//   beta : lane 7 starts from the tail-bit metrics (computed per block at demux time in
//          wrapping int16, :474-520), lanes 0..6 from their own final alpha (:531-538);
//          full pass; then the last 5 steps are re-run seeded with beta[0] of the next
//          lane (:541-549,585-587).  ext(k) reads alpha[k] and beta[k+1] of the final
//          arrays, i.e. re-run values for k<=5 (alpha) and k>=W-6 (beta).
#pragma once
#include <type_traits>
#include "td_common.cuh"

namespace oai {

constexpr int MAP_THREADS = 64;      // 16 code blocks per CTA
#ifndef MAP_SEG_STEPS
#define MAP_SEG_STEPS 16
#endif
constexpr int MAP_SEG = MAP_SEG_STEPS;   // checkpoint distance S of the fast path (steps)
constexpr int MAP_ABUF_ENTRIES = 8;      // alpha entries per thread in shared memory (fast path: the even steps of a segment;
                                         // exact path: all steps of its 8-step segments)
constexpr int MAP_CKPT_STEPS = 8;        // the checkpoint pool is sized for the smaller of the two distances
// dynamic shared memory per CTA: 32 B of alpha per entry and thread
constexpr int MAP_SMEM_BYTES = MAP_ABUF_ENTRIES * MAP_THREADS * 32;
constexpr int RERUN_STEPS = 5;       // L>>3, reference :171,189
constexpr int NEG_INIT = -128;       // -MAX/2, reference :79,201

struct MapArgs {
  const CbMeta* meta;
  CbState* state;
  int16_t* ws;           // workspace base
  long slot_hw;          // halfwords per block slot (ARR_COUNT * A)
  int A;                 // halfwords per array
  u32* ckpt;             // checkpoint pool
  long ckpt_words;       // words per block in the pool
  int nblk;
  int sys_arr, par_arr, out_arr;
  int term;              // 0: first constituent decoder, 1: second
  int iter;              // blocks with max_iter < iter are finished (skipped)
  int guard_b;           // fast path allowed when max(max_sys,max_in) + max_in <= guard_b
  const int* batch_max;  // max |y| over the batch: <= 127 selects the int8 parity / s0 copies
  int upd;               // 1: write ext = (ext (-) sys) (+) s0 (the feedback step, reference :1354-1375)
  const int* active;     // compacted list of running blocks (k_compact) or nullptr = all nblk blocks
  const int* nactive;    // its three counters (fast / tracked / exact class)
  int retry;             // 1: the retry launch -- only blocks whose tracked pass failed run, on the exact policy
  int* retry_flag;       // a failing tracked pass stores `seq` here; the retry launch returns at once unless it finds it
  int seq;               // number of this MAP pass within the decode (1, 2, ...); the flag is zeroed when the decode starts
  int track;             // 0: no tracked passes (blocks beyond the guard go to the exact policy, the round-1 behaviour)
  int force;             // test hook: 0 = decide, 1 = untracked fast, 2 = exact, 3 = tracked fast
};

__device__ __forceinline__ u32 pick4(const uint4& v, int q) { return q == 0 ? v.x : (q == 1 ? v.y : (q == 2 ? v.z : v.w)); }

template <class AR>
struct Gam { u32 g1, g0, n1, n0; };

template <class AR>
__device__ __forceinline__ Gam<AR> gamma2(u32 s, u32 p) {
  Gam<AR> g;
  g.g1 = vsra1(AR::add(s, p));     // m11, reference :146
  g.g0 = vsra1(AR::sub(s, p));     // m10, reference :147
  g.n1 = __vneg2(g.g1);            // |m| <= 16384, negation is exact
  g.n0 = __vneg2(g.g0);
  return g;
}

// forward add-compare-select + max normalisation, reference :292-330,373-380.
// a - g is computed as a + (-g): exact because |g| <= 16384.
template <class AR>
__device__ __forceinline__ void alpha_step(u32 (&a)[8], const Gam<AR>& g) {
  u32 n0 = vmax(AR::add(a[1], g.g1), AR::add(a[0], g.n1));
  u32 n1 = vmax(AR::add(a[3], g.n0), AR::add(a[2], g.g0));
  u32 n2 = vmax(AR::add(a[5], g.g0), AR::add(a[4], g.n0));
  u32 n3 = vmax(AR::add(a[7], g.n1), AR::add(a[6], g.g1));
  u32 n4 = vmax(AR::add(a[1], g.n1), AR::add(a[0], g.g1));
  u32 n5 = vmax(AR::add(a[3], g.g0), AR::add(a[2], g.n0));
  u32 n6 = vmax(AR::add(a[5], g.n0), AR::add(a[4], g.g0));
  u32 n7 = vmax(AR::add(a[7], g.g1), AR::add(a[6], g.n1));
  u32 mx = vmax(vmax(vmax(n0, n1), vmax(n2, n3)), vmax(vmax(n4, n5), vmax(n6, n7)));
  a[0] = AR::sub(n0, mx); a[1] = AR::sub(n1, mx); a[2] = AR::sub(n2, mx); a[3] = AR::sub(n3, mx);
  a[4] = AR::sub(n4, mx); a[5] = AR::sub(n5, mx); a[6] = AR::sub(n6, mx); a[7] = AR::sub(n7, mx);
}

// backward recursion, reference :592-636
template <class AR>
__device__ __forceinline__ void beta_step(u32 (&b)[8], const Gam<AR>& g) {
  u32 n0 = vmax(AR::add(b[4], g.g1), AR::add(b[0], g.n1));
  u32 n1 = vmax(AR::add(b[4], g.n1), AR::add(b[0], g.g1));
  u32 n2 = vmax(AR::add(b[5], g.n0), AR::add(b[1], g.g0));
  u32 n3 = vmax(AR::add(b[5], g.g0), AR::add(b[1], g.n0));
  u32 n4 = vmax(AR::add(b[6], g.g0), AR::add(b[2], g.n0));
  u32 n5 = vmax(AR::add(b[6], g.n0), AR::add(b[2], g.g0));
  u32 n6 = vmax(AR::add(b[7], g.n1), AR::add(b[3], g.g1));
  u32 n7 = vmax(AR::add(b[7], g.g1), AR::add(b[3], g.n1));
  u32 mx = vmax(vmax(vmax(n0, n1), vmax(n2, n3)), vmax(vmax(n4, n5), vmax(n6, n7)));
  b[0] = AR::sub(n0, mx); b[1] = AR::sub(n1, mx); b[2] = AR::sub(n2, mx); b[3] = AR::sub(n3, mx);
  b[4] = AR::sub(n4, mx); b[5] = AR::sub(n5, mx); b[6] = AR::sub(n6, mx); b[7] = AR::sub(n7, mx);
}

// ---- the same recursions in the inverted representation of InvArith (td_common.cuh) ----
constexpr u32 KINV_CAP = 0xbfffbfffu;      // R(-32768) = 49151
__device__ __forceinline__ void norm_inv(u32 (&v)[8], const u32 (&n)[8]) {
  const u32 mn = __vminu2(__vimin3_u16x2(n[0], n[1], n[2]), __vimin3_u16x2(n[3], n[4], __vimin3_u16x2(n[5], n[6], n[7])));   // R(max)
  const u32 c = __vadd2(~mn, 0x40004000u);                       // 16383 - R(max)
#pragma unroll
  for (int s = 0; s < 8; ++s) v[s] = __viaddmin_u16x2(n[s], c, KINV_CAP);   // R(sat(n - max)) = min(n' - mn' + 16383, 49151)
}
// R(max(sat(x + gx), sat(y + gy))) = min(x' - gx, y' - gy, 49151); callers pass the NEGATED branch metrics
__device__ __forceinline__ u32 acs_inv(u32 x, u32 ngx, u32 y, u32 ngy) {
  return __vimin3_u16x2(__vadd2(x, ngx), __vadd2(y, ngy), KINV_CAP);
}
template <class G>
__device__ __forceinline__ void alpha_step_inv(u32 (&a)[8], const G& g) {
  u32 n[8];
  n[0] = acs_inv(a[1], g.n1, a[0], g.g1);
  n[1] = acs_inv(a[3], g.g0, a[2], g.n0);
  n[2] = acs_inv(a[5], g.n0, a[4], g.g0);
  n[3] = acs_inv(a[7], g.g1, a[6], g.n1);
  n[4] = acs_inv(a[1], g.g1, a[0], g.n1);
  n[5] = acs_inv(a[3], g.n0, a[2], g.g0);
  n[6] = acs_inv(a[5], g.g0, a[4], g.n0);
  n[7] = acs_inv(a[7], g.n1, a[6], g.g1);
  norm_inv(a, n);
}
template <class G>
__device__ __forceinline__ void beta_step_inv(u32 (&b)[8], const G& g) {
  u32 n[8];
  n[0] = acs_inv(b[4], g.n1, b[0], g.g1);
  n[1] = acs_inv(b[4], g.g1, b[0], g.n1);
  n[2] = acs_inv(b[5], g.g0, b[1], g.n0);
  n[3] = acs_inv(b[5], g.n0, b[1], g.g0);
  n[4] = acs_inv(b[6], g.n0, b[2], g.g0);
  n[5] = acs_inv(b[6], g.g0, b[2], g.n0);
  n[6] = acs_inv(b[7], g.g1, b[3], g.n1);
  n[7] = acs_inv(b[7], g.n1, b[3], g.g1);
  norm_inv(b, n);
}

// a-posteriori LLR of one step (reference :757-818) from alpha, beta in the inverted representation; returns the SIGNED
// packed LLR.  The 16 sums a + b of two metrics in [-32768, 0] saturate at -32768; with magnitudes ua = -a, ub = -b in
// [0, 32768] that is min(ua + ub, 32768), and ua + ub fits 16 bits except for 32768 + 32768: with ua~ = min(ua, 32767),
// max(ua~ + ub, ua + ub~) equals ua + ub except in that case, where it is 65535 -- still above the cap.  The cap commutes
// with the max over the four sums of a group.  Adding +-gamma needs 16383 - x >= 0 for the added metric x, i.e. no
// metric of +16384: the caller routes such steps (hazard) to the signed form.
template <class G>
__device__ __forceinline__ u32 ext_step_inv(const u32 (&al)[8], const u32 (&be)[8], const G& g) {
  constexpr u32 K16383 = 0x3fff3fffu, K32767 = 0x7fff7fffu, K32768 = 0x80008000u;
  u32 ua[8], ub[8], uat[8], ubt[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    ua[s] = __vsub2(al[s], K16383); ub[s] = __vsub2(be[s], K16383);
    uat[s] = __vminu2(ua[s], K32767); ubt[s] = __vminu2(ub[s], K32767);
  }
  auto sum = [&](int i, int j) -> u32 { return __vmaxu2(__vadd2(uat[i], ub[j]), __vadd2(ua[i], ubt[j])); };
  auto grp = [&](int i0, int j0, int i1, int j1, int i2, int j2, int i3, int j3) -> u32 {   // magnitude of the saturated max
    return __vimin3_u16x2(__vimin3_u16x2(sum(i0, j0), sum(i1, j1), K32768), sum(i2, j2), sum(i3, j3));
  };
  const u32 u00 = grp(0, 0, 1, 4, 6, 7, 7, 3);
  const u32 u11 = grp(0, 4, 1, 0, 6, 3, 7, 7);
  const u32 u01 = grp(2, 5, 3, 1, 4, 2, 5, 6);
  const u32 u10 = grp(2, 1, 3, 5, 4, 6, 5, 2);
  // R(m + x) = um + (16383 - x), capped at R(-32768)
  const u32 eg1 = __vadd2(~g.g1, 0x40004000u), eg0 = __vadd2(~g.g0, 0x40004000u);          // 16383 - g
  const u32 r01 = __viaddmin_u16x2(u01, __vadd2(g.g0, K16383), KINV_CAP);                   // m01 + n0, n0 = -g0
  const u32 r00 = __viaddmin_u16x2(u00, __vadd2(g.g1, K16383), KINV_CAP);                   // m00 + n1
  const u32 r10 = __viaddmin_u16x2(u10, eg0, KINV_CAP);                                     // m10 + g0
  const u32 r11 = __viaddmin_u16x2(u11, eg1, KINV_CAP);                                     // m11 + g1
  const u32 ru = __vminu2(r10, r11), rv = __vminu2(r01, r00);                               // R of the two maxima
  return __vsubss2(__vadd2(~ru, 0x40004000u), __vadd2(~rv, 0x40004000u));
}

// ---- hazard steps: a branch metric of exactly -16384 in some halfword ----------------------------------------
// With the NEGATED metric +16384 the candidate 0 + 16384 = 16384 lies one above what R can hold.  In the halfwords
// concerned the whole step is shifted by one (R_o(v) = 16383 + o - v, o = 1): the low end of the candidate range moves to
// 0, the high end, -32768 - 16384 -> 65536, would wrap -- but only for the operand -32768 in the add of -16384, whose
// result is capped to -32768 anyway, so that operand is clamped to -32767 first (R <= 49150) and still lands above the
// cap.  The max-normalisation subtracts two values of the same offset, so its result is back in the plain representation.
struct HzOff { u32 o, k1, k0; };     // o: offset per halfword; k1 / k0: clamp of the operands that take -g1 / -g0 negated, i.e. +16384
template <class G>
__device__ __forceinline__ HzOff hz_offsets(const G& g) {
  HzOff h;
  const u32 o1 = __vcmpeq2(g.g1, 0xC000C000u) & 0x00010001u, o0 = __vcmpeq2(g.g0, 0xC000C000u) & 0x00010001u;
  h.o = o1 | o0; h.k1 = KINV_CAP - o1; h.k0 = KINV_CAP - o0;
  return h;
}
// min(min(xc, kc) + mc, y + my, cap): xc is the operand added to the negated metric
__device__ __forceinline__ u32 acs_inv_hz(u32 xc, u32 kc, u32 mc, u32 y, u32 my, u32 cap) {
  return __vimin3_u16x2(__vadd2(__vminu2(xc, kc), mc), __vadd2(y, my), cap);
}
template <class G>
__device__ __forceinline__ void alpha_step_inv_hz(u32 (&a)[8], const G& g) {
  const HzOff h = hz_offsets(g);
  const u32 cap = KINV_CAP + h.o;
  const u32 n1 = __vadd2(g.n1, h.o), g1 = __vadd2(g.g1, h.o), n0 = __vadd2(g.n0, h.o), g0 = __vadd2(g.g0, h.o);
  u32 n[8];
  n[0] = acs_inv_hz(a[1], h.k1, n1, a[0], g1, cap);
  n[1] = acs_inv_hz(a[2], h.k0, n0, a[3], g0, cap);
  n[2] = acs_inv_hz(a[5], h.k0, n0, a[4], g0, cap);
  n[3] = acs_inv_hz(a[6], h.k1, n1, a[7], g1, cap);
  n[4] = acs_inv_hz(a[0], h.k1, n1, a[1], g1, cap);
  n[5] = acs_inv_hz(a[3], h.k0, n0, a[2], g0, cap);
  n[6] = acs_inv_hz(a[4], h.k0, n0, a[5], g0, cap);
  n[7] = acs_inv_hz(a[7], h.k1, n1, a[6], g1, cap);
  norm_inv(a, n);
}
template <class G>
__device__ __forceinline__ void beta_step_inv_hz(u32 (&b)[8], const G& g) {
  const HzOff h = hz_offsets(g);
  const u32 cap = KINV_CAP + h.o;
  const u32 n1 = __vadd2(g.n1, h.o), g1 = __vadd2(g.g1, h.o), n0 = __vadd2(g.n0, h.o), g0 = __vadd2(g.g0, h.o);
  u32 n[8];
  n[0] = acs_inv_hz(b[4], h.k1, n1, b[0], g1, cap);
  n[1] = acs_inv_hz(b[0], h.k1, n1, b[4], g1, cap);
  n[2] = acs_inv_hz(b[1], h.k0, n0, b[5], g0, cap);
  n[3] = acs_inv_hz(b[5], h.k0, n0, b[1], g0, cap);
  n[4] = acs_inv_hz(b[6], h.k0, n0, b[2], g0, cap);
  n[5] = acs_inv_hz(b[2], h.k0, n0, b[6], g0, cap);
  n[6] = acs_inv_hz(b[3], h.k1, n1, b[7], g1, cap);
  n[7] = acs_inv_hz(b[7], h.k1, n1, b[3], g1, cap);
  norm_inv(b, n);
}
// LLR of a hazard step: the four sums m + x get the same offset; m10 + g0 and m11 + g1 add -16384 to a magnitude that
// may be 32768 -- clamped to 32767 like above.  16383 + o - r turns the two minima back into exact signed values.
template <class G>
__device__ __forceinline__ u32 ext_step_inv_hz(const u32 (&al)[8], const u32 (&be)[8], const G& g) {
  constexpr u32 K16383 = 0x3fff3fffu, K32767 = 0x7fff7fffu, K32768 = 0x80008000u;
  u32 ua[8], ub[8], uat[8], ubt[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    ua[s] = __vsub2(al[s], K16383); ub[s] = __vsub2(be[s], K16383);
    uat[s] = __vminu2(ua[s], K32767); ubt[s] = __vminu2(ub[s], K32767);
  }
  auto sum = [&](int i, int j) -> u32 { return __vmaxu2(__vadd2(uat[i], ub[j]), __vadd2(ua[i], ubt[j])); };
  auto grp = [&](int i0, int j0, int i1, int j1, int i2, int j2, int i3, int j3) -> u32 {
    return __vimin3_u16x2(__vimin3_u16x2(sum(i0, j0), sum(i1, j1), K32768), sum(i2, j2), sum(i3, j3));
  };
  const u32 u00 = grp(0, 0, 1, 4, 6, 7, 7, 3);
  const u32 u11 = grp(0, 4, 1, 0, 6, 3, 7, 7);
  const u32 u01 = grp(2, 5, 3, 1, 4, 2, 5, 6);
  const u32 u10 = grp(2, 1, 3, 5, 4, 6, 5, 2);
  const u32 o1 = __vcmpeq2(g.g1, 0xC000C000u) & 0x00010001u, o0 = __vcmpeq2(g.g0, 0xC000C000u) & 0x00010001u, o = o1 | o0;
  const u32 cap = KINV_CAP + o, ko = K16383 + o;
  const u32 r01 = __viaddmin_u16x2(u01, __vadd2(g.g0, ko), cap);                                   // m01 - g0
  const u32 r00 = __viaddmin_u16x2(u00, __vadd2(g.g1, ko), cap);                                   // m00 - g1
  const u32 r10 = __viaddmin_u16x2(__vminu2(u10, K32768 - o0), __vadd2(~g.g0, 0x40004000u + o), cap);   // m10 + g0
  const u32 r11 = __viaddmin_u16x2(__vminu2(u11, K32768 - o1), __vadd2(~g.g1, 0x40004000u + o), cap);   // m11 + g1
  const u32 ru = __vminu2(r10, r11), rv = __vminu2(r01, r00);
  return __vsubss2(__vadd2(~ru, 0x40004000u + o), __vadd2(~rv, 0x40004000u + o));
}

// a-posteriori LLR of one step, reference :757-818
template <class AR>
__device__ __forceinline__ u32 ext_step(const u32 (&a)[8], const u32 (&b)[8], const Gam<AR>& g) {
  u32 m00 = vmax(vmax(AR::add(a[0], b[0]), AR::add(a[1], b[4])), vmax(AR::add(a[6], b[7]), AR::add(a[7], b[3])));
  u32 m11 = vmax(vmax(AR::add(a[0], b[4]), AR::add(a[1], b[0])), vmax(AR::add(a[6], b[3]), AR::add(a[7], b[7])));
  u32 m01 = vmax(vmax(AR::add(a[2], b[5]), AR::add(a[3], b[1])), vmax(AR::add(a[4], b[2]), AR::add(a[5], b[6])));
  u32 m10 = vmax(vmax(AR::add(a[2], b[1]), AR::add(a[3], b[5])), vmax(AR::add(a[4], b[6]), AR::add(a[5], b[2])));
  m01 = AR::add(m01, g.n0);
  m00 = AR::add(m00, g.n1);
  m10 = AR::add(m10, g.g0);
  m11 = AR::add(m11, g.g1);
  return AR::sub(vmax(m10, m11), vmax(m01, m00));
}

// shared-memory alpha segment buffer: entry e of thread tid = two uint4 at
// [(2e+h)*MAP_THREADS + tid]  (conflict-free 128-bit accesses)
__device__ __forceinline__ void abuf_put(uint4* abuf, int e, int tid, const u32 (&a)[8]) {
  abuf[(2 * e) * MAP_THREADS + tid] = make_uint4(a[0], a[1], a[2], a[3]);
  abuf[(2 * e + 1) * MAP_THREADS + tid] = make_uint4(a[4], a[5], a[6], a[7]);
}
__device__ __forceinline__ void abuf_get(const uint4* abuf, int e, int tid, u32 (&a)[8]) {
  uint4 x = abuf[(2 * e) * MAP_THREADS + tid], y = abuf[(2 * e + 1) * MAP_THREADS + tid];
  a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w; a[4] = y.x; a[5] = y.y; a[6] = y.z; a[7] = y.w;
}

__device__ __forceinline__ void ckpt_put(u32* c, const u32 (&a)[8]) {
  reinterpret_cast<uint4*>(c)[0] = make_uint4(a[0], a[1], a[2], a[3]);
  reinterpret_cast<uint4*>(c)[1] = make_uint4(a[4], a[5], a[6], a[7]);
}
__device__ __forceinline__ void ckpt_get(const u32* c, u32 (&a)[8]) {
  uint4 x = reinterpret_cast<const uint4*>(c)[0], y = reinterpret_cast<const uint4*>(c)[1];
  a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w; a[4] = y.x; a[5] = y.y; a[6] = y.z; a[7] = y.w;
}

// One MAP pass for the block owned by this 4-thread group.
//   sys/par/ext: this thread's view of the C4 arrays (uint32 words, already offset by t*4)
//   ck: this thread's checkpoint area: slot i at ck + i*32 words (8 words used per thread,
//       threads interleaved by the caller through the base pointer)
template <class AR, int S>
__device__ void map_pass(const u32* __restrict__ sys, const u32* __restrict__ par, const u32* __restrict__ s0, bool upd,
                         u32* __restrict__ ext, u32* ck, int W, int t, unsigned gmask, const int16_t* Tv, uint4* abuf,
                         int tid) {
  constexpr bool INV = AR::kInv;      // alpha / beta held as R(v) = 16383 - v (InvArith), else packed signed int16
  // feedback step fused into the output: ext = (ext (-) sys) (+) s0 with the reference's saturation
  auto fb = [&](u32 x, int k) -> u32 {
    if (!upd) return x;
    return __vaddss2(__vsubss2(x, __ldg(sys + c4_word(k, 0))), __ldg(s0 + c4_word(k, 0)));
  };
  auto gam = [&](int k) { return gamma2<AR>(__ldg(sys + c4_word(k, 0)), __ldg(par + c4_word(k, 0))); };
  // In the inverted representation a candidate a + g must stay <= 16383; a = 0 with the NEGATED branch metric
  // -(-16384) = +16384 is the one combination that does not (it needs sat(s +- p) = -32768).  Such steps run in signed
  // saturating arithmetic like the first beta step below.
  auto hazard = [&](const Gam<AR>& g) -> bool {
    auto zh = [](u32 x) -> u32 { return (x - 0x00010001u) & ~x & 0x80008000u; };     // some halfword of x is zero
    return (zh(g.g1 ^ 0xC000C000u) | zh(g.g0 ^ 0xC000C000u)) != 0;
  };
  // Every step picks its form from its own branch metrics; both forms stay in the inverted representation.  Only the first
  // beta step of a sweep (b still signed, see below) runs in signed saturating arithmetic.
  auto astep = [&](u32 (&x)[8], const Gam<AR>& g) { if (hazard(g)) alpha_step_inv_hz(x, g); else alpha_step_inv(x, g); };
  auto bstep = [&](u32 (&be)[8], const Gam<AR>& g) { if (hazard(g)) beta_step_inv_hz(be, g); else beta_step_inv(be, g); };
  auto extv = [&](const u32 (&al)[8], const u32 (&be)[8], const Gam<AR>& g) -> u32 {
    return hazard(g) ? ext_step_inv_hz(al, be, g) : ext_step_inv(al, be, g);
  };
  // The beta vector that enters the FIRST step of a sweep contains the tail metrics of lane 7, which are computed in
  // wrapping int16 and may be positive (TD16:474-520): that step (and the LLR that reads this vector) runs in signed
  // saturating arithmetic; its max-normalised result is <= 0 and goes on in the inverted representation.
  auto ext_first = [&](const u32 (&al)[8], const u32 (&bs)[8], const Gam<AR>& g) -> u32 {
    u32 as[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) as[s] = AR::dec(al[s]);
    return ext_step<AR>(as, bs, g);
  };
  auto bstep_first = [&](u32 (&be)[8], const Gam<AR>& g) {
    beta_step<AR>(be, g);
#pragma unroll
    for (int s = 0; s < 8; ++s) be[s] = AR::enc(be[s]);
  };
  const int nseg = (W + S - 1) / S;
  u32 a[8];
  static_assert(S == 8 && INV, "exact policy: 8-step segments (two 4-step chunks per stream) in the inverted representation");
  const uint4* sys4 = reinterpret_cast<const uint4*>(sys);     // chunk c (4 steps) of this thread at [c*4]
  const uint4* par4 = reinterpret_cast<const uint4*>(par);
  const uint4* s04 = reinterpret_cast<const uint4*>(s0);
  const int nchunk = (W + 3) >> 2;
  // Branch metrics of a segment, computed once and shared by the alpha recomputation, the LLR and the beta step.
  // Measured (tools/exact_path_probe.py, uniform noise +-3000 / +-12000, Gbit/s): every step testing for itself with a
  // signed-arithmetic fallback 7.5 / 4.3; the same with the in-representation hazard form 7.5 / 6.3; a per-segment mask of
  // hazard steps 8.0 / 4.3; two bodies, the hazard one testing every step 9.1 / 2.6; the 4-thread group instead of the
  // warp voting on the body 8.4 / 3.3; two branch-free bodies chosen per segment 9.0 / 3.5; hazard body only 6.5 / 6.4;
  // this version (per segment, but per pass once hazards are frequent) 9.0 / 6.3.
  // Returns whether some thread of the warp has a hazard step in the segment (min over all metrics == -16384; they are
  // >= -16384 by construction).  Segments without one run the plain steps, the others the hazard form of every step
  // (which equals the plain form where the offset is 0); both bodies are branch-free, the choice is warp-uniform
  // (threads that disagree would make the warp run both bodies) and changes no result.
  auto seg_gamma = [&](const uint4 (&sv)[2], const uint4 (&pv)[2], int k0, int k1, Gam<AR> (&g8)[S], bool force) -> bool {
    u32 m = 0;
#pragma unroll
    for (int e = 0; e < S; ++e) {
      if (k0 + e < k1) {
        g8[e] = gamma2<AR>(pick4(sv[e >> 2], e & 3), pick4(pv[e >> 2], e & 3));
        m = __vimin3_s16x2(m, g8[e].g1, g8[e].g0);
      }
    }
    const u32 x = m ^ 0xC000C000u;
    return __any_sync(__activemask(), force || ((x - 0x00010001u) & ~x & 0x80008000u) != 0);
  };
  auto load_seg2 = [&](int seg, uint4 (&sv)[2], uint4 (&pv)[2]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = seg * 2 + j;
      if (c < nchunk) { sv[j] = __ldg(sys4 + c * 4); pv[j] = __ldg(par4 + c * 4); }
    }
  };

  // ---- forward sweep (alpha pass 1), checkpoint every S steps ----------------------
  int nhz = 0;                                                  // hazard segments met by this warp
#pragma unroll
  for (int s = 0; s < 8; ++s) a[s] = AR::enc(pack2(NEG_INIT, NEG_INIT));
  if (t == 0) a[0] = AR::enc(pack2(0, NEG_INIT));               // reference :201-208
  {
    uint4 sf[2], pf[2], sfn[2], pfn[2];
    load_seg2(0, sf, pf);
    nhz = 0;
    for (int seg = 0; seg < nseg; ++seg) {
      const int k0 = seg * S, k1 = min(W, k0 + S);
      if (seg + 1 < nseg) load_seg2(seg + 1, sfn, pfn);         // one segment ahead
      Gam<AR> g8[S];
      const bool hz = seg_gamma(sf, pf, k0, k1, g8, false);
      nhz += hz ? 1 : 0;
      ckpt_put(ck + seg * 32, a);
      if (!hz) {
#pragma unroll
        for (int e = 0; e < S; ++e)
          if (k0 + e < k1) alpha_step_inv(a, g8[e]);
      } else {
#pragma unroll
        for (int e = 0; e < S; ++e)
          if (k0 + e < k1) alpha_step_inv_hz(a, g8[e]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) { sf[j] = sfn[j]; pf[j] = pfn[j]; }
    }
  }

  // ---- alpha re-run seed: lane l <- final metrics of lane l-1, lane 0 <- start -----
  u32 seed[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    u32 prev = __shfl_sync(gmask, a[s], (t + 3) & 3, 4);      // thread t-1 (lanes 2t-2, 2t-1)
    if (t == 0) prev = AR::enc(pack2(0, (s == 0) ? 0 : NEG_INIT));   // hi half is what gets used
    seed[s] = __byte_perm(prev, a[s], 0x5432);                 // lo <- prev.hi, hi <- mine.lo
  }
  ckpt_put(ck + nseg * 32, seed);
  if (W <= RERUN_STEPS) {
    // the re-run covers the whole lane, so alpha[W] itself is the re-run value (K=40)
#pragma unroll
    for (int s = 0; s < 8; ++s) a[s] = seed[s];
    for (int k = 0; k < W; ++k) astep(a, gam(k));
  }

  // ---- beta start: lanes 0..6 <- own alpha[W], lane 7 <- tail metrics (signed form, see bstep_first) ----
  u32 b[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    b[s] = AR::dec(a[s]);
    if (t == 3) b[s] = (b[s] & 0xffffu) | ((u32)(uint16_t)Tv[s] << 16);
  }
  bstep_first(b, gam(W - 1));        // step W-1 of the sweep; the segment loop below skips it

  // ---- backward sweep, pass 1 -----------------------------------------------------
  // A segment is two 4-step chunks per stream; its inputs are loaded once into registers (the next segment's are
  // requested at the segment start, a whole segment ahead of their use) and its 8 branch-metric pairs are computed once
  // and shared by the alpha recomputation, the LLR and the beta step.
  // The two unrolled bodies of the backward sweep do not fit the instruction cache together: a pass in which more than
  // a quarter of the segments have a hazard step (uniform +-12000 noise: alternating bodies ran at 3.5 Gbit/s, the hazard
  // body alone at 6.4) takes the hazard body for every segment.
  const bool always_hz = nhz * 4 > nseg;
  uint4 sc[2], pc[2], zc[2], sn[2], pn[2], zn[2];
  auto load_seg = [&](int seg, uint4 (&sv)[2], uint4 (&pv)[2], uint4 (&zv)[2]) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = seg * 2 + j;
      if (c < nchunk) { sv[j] = __ldg(sys4 + c * 4); pv[j] = __ldg(par4 + c * 4); if (upd) zv[j] = __ldg(s04 + c * 4); }
    }
  };
  load_seg(nseg - 1, sc, pc, zc);
  for (int seg = nseg - 1; seg >= 0; --seg) {
    const int k0 = seg * S, k1 = min(W, k0 + S);
    if (seg > 0) load_seg(seg - 1, sn, pn, zn);
    ckpt_get(ck + seg * 32, a);
    Gam<AR> g8[S];
    const bool hz = seg_gamma(sc, pc, k0, k1, g8, always_hz);
    auto seg_body = [&](auto hz_tag) {
      constexpr bool HZ = decltype(hz_tag)::value;
      auto a_st = [&](const Gam<AR>& g) { if constexpr (HZ) alpha_step_inv_hz(a, g); else alpha_step_inv(a, g); };
#pragma unroll
      for (int e = 0; e < S; ++e) {
        if (k0 + e < k1) {
          abuf_put(abuf, e, tid, a);
          if (k0 + e + 1 < k1) a_st(g8[e]);
        }
      }
      if (seg == 0) {          // alpha[0..5] come from the re-run chain
#pragma unroll
        for (int s = 0; s < 8; ++s) a[s] = seed[s];
#pragma unroll
        for (int k = 0; k <= RERUN_STEPS; ++k) {
          if (k < k1) {
            abuf_put(abuf, k, tid, a);
            if (k < RERUN_STEPS) a_st(g8[k]);
          }
        }
      }
#pragma unroll
      for (int j = 1; j >= 0; --j) {
        u32 e4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int q = 3; q >= 0; --q) {
          const int e = j * 4 + q, k = k0 + e;
          if (k < k1) {
            if (k <= W - 7) {       // steps whose beta[k+1] is not replaced by the re-run
              abuf_get(abuf, e, tid, a);
              u32 x;
              if constexpr (HZ) x = ext_step_inv_hz(a, b, g8[e]); else x = ext_step_inv(a, b, g8[e]);
              if (upd) x = __vaddss2(__vsubss2(x, pick4(sc[j], q)), pick4(zc[j], q));     // feedback, reference :1354-1375
              e4[q] = x;
            }
            if (k != W - 1) { if constexpr (HZ) beta_step_inv_hz(b, g8[e]); else beta_step_inv(b, g8[e]); }
          }
        }
        const int kc = k0 + j * 4;
        if (kc + 3 <= W - 7) *reinterpret_cast<uint4*>(ext + ((kc >> 2) << 4)) = make_uint4(e4[0], e4[1], e4[2], e4[3]);
        else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (kc + q < k1 && kc + q <= W - 7) ext[c4_word(kc + q, 0)] = e4[q];
        }
      }
    };
    if (hz) seg_body(std::true_type{}); else seg_body(std::false_type{});
#pragma unroll
    for (int j = 0; j < 2; ++j) { sc[j] = sn[j]; pc[j] = pn[j]; zc[j] = zn[j]; }
  }

  // ---- beta re-run: lane l <- beta[0] of lane l+1, lane 7 <- tail metrics ------------
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const u32 mine = AR::dec(b[s]);                            // back to signed form: lane 7 takes the tail metrics again
    u32 next = __shfl_sync(gmask, mine, (t + 1) & 3, 4);      // thread t+1 (lanes 2t+2, 2t+3)
    if (t == 3) next = (u32)(uint16_t)Tv[s];
    b[s] = __byte_perm(mine, next, 0x5432);                    // lo <- mine.hi, hi <- next.lo
  }
  {
    const int kk0 = max(W - 6, 0);
    const int sa = kk0 / S;
    ckpt_get(ck + sa * 32, a);
    for (int k = sa * S; k < W; ++k) {
      if (k >= kk0) abuf_put(abuf, k - kk0, tid, a);
      if (k + 1 < W) astep(a, gam(k));
    }
    if (kk0 <= RERUN_STEPS) {
#pragma unroll
      for (int s = 0; s < 8; ++s) a[s] = seed[s];
      for (int k = 0; k <= RERUN_STEPS && k < W; ++k) {
        if (k >= kk0) abuf_put(abuf, k - kk0, tid, a);
        if (k < RERUN_STEPS) astep(a, gam(k));
      }
    }
    {                                                          // step W-1: b is still signed
      const Gam<AR> g = gam(W - 1);
      abuf_get(abuf, W - 1 - kk0, tid, a);
      ext[c4_word(W - 1, 0)] = fb(ext_first(a, b, g), W - 1);
      bstep_first(b, g);
    }
    for (int k = W - 2; k >= kk0; --k) {
      const Gam<AR> g = gam(k);
      abuf_get(abuf, k - kk0, tid, a);
      ext[c4_word(k, 0)] = fb(extv(a, b, g), k);
      if (k >= W - RERUN_STEPS) bstep(b, g);                   // loopval=(n-40)>>3, reference :585
    }
  }
}


// =====================================================================================
// Fast path: non-saturating DPX arithmetic (VIADDMNMX.S16x2 / VIADD.16x2), used only
// when the per-pass guard proves that (a) the reference itself never saturates in this
// pass, so its result equals exact integer arithmetic, and (b) nothing below leaves the
// int16 range (DESIGN.md "fast-path guard").  Under (a) the a-posteriori LLR
//     ext = max_{u=1}(alpha+gamma+beta) - max_{u=0}(alpha+gamma+beta)
// is invariant to adding any per-(step,lane) constant to all alpha states, to all beta
// states, or to all four branch metrics of a step.  We use that freedom three ways:
//   * branch metrics are shifted by +m11:  {+m11,-m11,+m10,-m10} -> {X,0,Y,Z} with
//     X = 2*m11, Y = m11+m10, Z = m11-m10, so half of the add-compare-selects are a single
//     VIADDMNMX (max(a+X, b)) and the other half VIADD + VIADDMNMX: 12 instead of 24
//     instructions per recursion step;
//   * with s' = s - ((s^p)&1):  X = s'+p, Y = s', Z = p exactly (m11,m10 are floor halves
//     of s+p and s-p, which have equal parity) -- 3 ALU instructions per step;
//   * metrics are not max-normalised every step: every P steps state 0 is subtracted from
//     all states (P from the guard so that the drift P*M stays inside int16).
// =====================================================================================
struct FC { u32 X, Y, Z; };

__device__ __forceinline__ void l2_prefetch(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// Measured and rejected (profiles/r1g_sweep_*, r1i_sweep_*, r1k_sweep_*): L2 evict_first / evict_last cache hints on the
// streams (neutral to -2 %), a second prefetch per 64-byte chunk (neutral), one cp.async.bulk.prefetch.L2 per segment
// (-9 %), prefetch distances of 2 / 4 / 5 / 6 segments (0 to -8 % against 3), register prefetch two segments ahead in
// the forward sweep (-1 %).

__device__ __forceinline__ u32 vaddmax(u32 a, u32 b, u32 c) { return __viaddmax_s16x2(a, b, c); }   // max(a+b, c)

__device__ __forceinline__ FC fconst(u32 s, u32 p) {
  FC c;
  u32 bit = (s ^ p) & 0x00010001u;
  c.Y = __vadd2(s, bit * 0xffffu);       // s - bit per halfword (IMAD on the FMA pipe + VIADD.16x2)
  c.X = __vadd2(c.Y, p);
  c.Z = p;
  return c;
}

__device__ __forceinline__ void alpha_fast(u32 (&a)[8], const FC& c) {
  u32 t2 = __vadd2(a[2], c.Z), t3 = __vadd2(a[3], c.Z), t4 = __vadd2(a[4], c.Z), t5 = __vadd2(a[5], c.Z);
  u32 n0 = vaddmax(a[1], c.X, a[0]);     // max(a1+m11, a0-m11) + m11
  u32 n4 = vaddmax(a[0], c.X, a[1]);
  u32 n3 = vaddmax(a[6], c.X, a[7]);
  u32 n7 = vaddmax(a[7], c.X, a[6]);
  u32 n1 = vaddmax(a[2], c.Y, t3);       // max(a3-m10, a2+m10) + m11
  u32 n5 = vaddmax(a[3], c.Y, t2);
  u32 n2 = vaddmax(a[5], c.Y, t4);
  u32 n6 = vaddmax(a[4], c.Y, t5);
  a[0] = n0; a[1] = n1; a[2] = n2; a[3] = n3; a[4] = n4; a[5] = n5; a[6] = n6; a[7] = n7;
}

__device__ __forceinline__ void beta_fast(u32 (&b)[8], const FC& c) {
  u32 t1 = __vadd2(b[1], c.Z), t2 = __vadd2(b[2], c.Z), t5 = __vadd2(b[5], c.Z), t6 = __vadd2(b[6], c.Z);
  u32 n0 = vaddmax(b[4], c.X, b[0]);     // max(b4+m11, b0-m11) + m11
  u32 n1 = vaddmax(b[0], c.X, b[4]);
  u32 n6 = vaddmax(b[3], c.X, b[7]);
  u32 n7 = vaddmax(b[7], c.X, b[3]);
  u32 n2 = vaddmax(b[1], c.Y, t5);       // max(b5-m10, b1+m10) + m11
  u32 n3 = vaddmax(b[5], c.Y, t1);
  u32 n4 = vaddmax(b[6], c.Y, t2);
  u32 n5 = vaddmax(b[2], c.Y, t6);
  b[0] = n0; b[1] = n1; b[2] = n2; b[3] = n3; b[4] = n4; b[5] = n5; b[6] = n6; b[7] = n7;
}

// max(m10+m10g, m11+m11g) - max(m01-m10g, m00-m11g) with both maxima shifted by +m11
__device__ __forceinline__ u32 ext_fast(const u32 (&a)[8], const u32 (&b)[8], const FC& c) {
  u32 m00 = vaddmax(a[7], b[3], vaddmax(a[6], b[7], vaddmax(a[1], b[4], __vadd2(a[0], b[0]))));
  u32 m11 = vaddmax(a[7], b[7], vaddmax(a[6], b[3], vaddmax(a[1], b[0], __vadd2(a[0], b[4]))));
  u32 m01 = vaddmax(a[5], b[6], vaddmax(a[4], b[2], vaddmax(a[3], b[1], __vadd2(a[2], b[5]))));
  u32 m10 = vaddmax(a[5], b[2], vaddmax(a[4], b[6], vaddmax(a[3], b[5], __vadd2(a[2], b[1]))));
  u32 u = vaddmax(m10, c.Y, __vadd2(m11, c.X));
  u32 v = vaddmax(m01, c.Z, m00);
  return __vsub2(u, v);
}

// subtract state 0 from all states (any uniform shift is allowed, see above)
__device__ __forceinline__ void renorm(u32 (&a)[8]) {
  u32 n = __vneg2(a[0]);
  a[0] = 0;
#pragma unroll
  for (int s = 1; s < 8; ++s) a[s] = __vadd2(a[s], n);
}

// ---- a-posteriori range check ("tracked" fast pass) -----------------------------------------------------------------
// The guard in front of the fast path is a worst-case bound (spread of a metric vector <= 10 Gmax + 128); measured spreads
// are ~2.5 Gmax (coded signals at A = 256: sum of the two spreads + 2 Gmax = 18 000 in the last pass of a decode, where
// the guard's bound is far beyond int16, and the reference never saturates).  A tracked pass runs the fast path with a
// renormalisation every step and records what actually happened:
//   sp_a / sp_b : the largest spread max_s - min_s of any alpha / beta vector of the pass (spreads are invariant to the
//                 fast path's shifts, so they ARE the reference's spreads as long as nothing has wrapped),
//   ext range   : the largest |LLR| before the feedback step.
// Certificate, evaluated at the end of the pass (k_map16):  sp_a + sp_b + M <= 32767  with M = B + 1 >= 2 Gmax, |X|,|Y|,|Z|.
// It implies (a) the reference saturated nowhere in this pass -- candidates a + g >= -(spread + Gmax), normalised values
// n - max >= -(spread + 2 Gmax), LLR sums >= -(sp_a + sp_b + Gmax), LLR within +-(sp_a + sp_b + 2 Gmax) -- so its result is
// exact integer arithmetic; and (b) the fast path wrapped nowhere: after the per-step renormalisation |a'| <= spread, so
// every candidate, sum and LLR is within sp_a + sp_b + M.  Soundness of reading the spreads off possibly wrapped values:
// let k* be the first step whose computation wraps or saturates; the vector entering it has its TRUE spread recorded, and
// that spread (together with the other sweep's and M) already violates the certificate, because a true spread within the
// bound makes step k* safe.  Spreads grow by at most 2 Gmax <= M per step, so the first unsafe vector is still
// representable (<= 32767 + M - M) and is recorded correctly.  With the feedback step fused in, max|LLR| + Ms + Mi <= 32767
// is checked as well (the reference's two saturating operations there).
// A pass whose certificate fails has written garbage; its inputs are intact, so the block is flagged and the pass is
// repeated on the exact policy by the retry launch that follows every k_map16 launch.
struct Trk { u32 sp_a, sp_b, emx, emn; };
__device__ __forceinline__ u32 vec_spread(const u32 (&a)[8]) {
  const u32 mx = __vmaxs2(__vimax3_s16x2(__vimax3_s16x2(a[0], a[1], a[2]), a[3], a[4]), __vimax3_s16x2(a[5], a[6], a[7]));
  const u32 mn = __vmins2(__vimin3_s16x2(__vimin3_s16x2(a[0], a[1], a[2]), a[3], a[4]), __vimin3_s16x2(a[5], a[6], a[7]));
  return __vsub2(mx, mn);                      // per halfword, as unsigned 16-bit
}
template <bool TRACK>
__device__ __forceinline__ void renorm_ta(u32 (&a)[8], Trk& tk) { if (TRACK) tk.sp_a = __vmaxu2(tk.sp_a, vec_spread(a)); renorm(a); }
template <bool TRACK>
__device__ __forceinline__ void renorm_tb(u32 (&b)[8], Trk& tk) { if (TRACK) tk.sp_b = __vmaxu2(tk.sp_b, vec_spread(b)); renorm(b); }
template <bool TRACK>
__device__ __forceinline__ void track_ext(u32 x, Trk& tk) { if (TRACK) { tk.emx = __vmaxs2(tk.emx, x); tk.emn = __vmins2(tk.emn, x); } }

// shared-memory alpha buffer of the fast path: NE entries of 8 packed states per thread
// (two conflict-free 128-bit rows per entry)
struct FastSmem {
  uint4* a0; uint4* a1;
  __device__ __forceinline__ FastSmem(unsigned char* base) {
    a0 = reinterpret_cast<uint4*>(base);
    a1 = a0 + MAP_ABUF_ENTRIES * MAP_THREADS;
  }
  __device__ __forceinline__ void put(int e, int tid, const u32 (&a)[8]) const {
    a0[e * MAP_THREADS + tid] = make_uint4(a[0], a[1], a[2], a[3]);
    a1[e * MAP_THREADS + tid] = make_uint4(a[4], a[5], a[6], a[7]);
  }
  __device__ __forceinline__ void get(int e, int tid, u32 (&a)[8]) const {
    uint4 x = a0[e * MAP_THREADS + tid], y = a1[e * MAP_THREADS + tid];
    a[0] = x.x; a[1] = x.y; a[2] = x.z; a[3] = x.w; a[4] = y.x; a[5] = y.y; a[6] = y.z; a[7] = y.w;
  }
};


// Fast MAP pass.  PM = P-1 with P the renormalisation period (1, 4 or 16 steps; compile time so
// that the unrolled steady-state code has no data-dependent branches).
//
// The kernel is bound by the SM's shared-memory/L1 data pipe (one 128-byte wavefront per cycle),
// so the steady state keeps shared-memory traffic to the minimum:
//   forward : alpha in registers, checkpoint (HBM) at every segment start (16 steps)
//   backward: per segment, the inputs of its 16 steps stay in registers (4 chunks of 4 steps, one
//             LDG.128 per chunk and stream); alpha is recomputed from the checkpoint and only the
//             EVEN steps are written to shared memory; the beta / ext sweep reloads an even alpha
//             and recomputes the following odd one from it (one extra alpha step per two trellis
//             steps instead of 32 B stored + 32 B loaded); branch constants are recomputed from
//             the input registers instead of going through shared memory.
//   The registers of chunk j are refilled with chunk j of the NEXT segment as soon as the beta sweep
//   has consumed them (chunk 0, which is needed first and freed last, goes through a spare set).
// P8: parity (and s0) are read from the int8 copies: chunk c of this thread = 8 bytes at par8[c*4]
template <int S, int PM, bool UPD, bool P8, bool TRACK = false>
__device__ void map_pass_fast(const u32* __restrict__ sys, const u32* __restrict__ par, const u32* __restrict__ s0,
                              u32* __restrict__ ext, u32* ck, int W, int t, unsigned gmask, const int16_t* Tv,
                              unsigned char* smem, int tid, Trk* trk = nullptr) {
  static_assert(S == 16, "steady-state segment = 4 chunks of 4 steps");
  static_assert(!TRACK || PM == 0, "a tracked pass renormalises (and records) every step");
  Trk tk;
  tk.sp_a = 0x00800080u;                                        // the start vectors (0, -128, ...) have spread 128
  tk.sp_b = 0; tk.emx = 0x80008000u; tk.emn = 0x7fff7fffu;
  constexpr int NCH = S / 4;
  constexpr int HS = MAP_ABUF_ENTRIES;                         // boundary code works on sub-segments of HS steps
#ifndef MAP_PF_SEGS
#define MAP_PF_SEGS 3
#endif
  constexpr int PF = MAP_PF_SEGS;                              // L2 prefetch distance in segments
  const FastSmem sm(smem);
  const int nseg = (W + S - 1) / S, nchunk = (W + 3) >> 2;
  const uint4* sys4 = reinterpret_cast<const uint4*>(sys);     // chunk c of this thread at sys4[c*4]
  const uint4* par4 = reinterpret_cast<const uint4*>(par);
  const uint4* s04 = reinterpret_cast<const uint4*>(s0);
  const uint2* par8 = reinterpret_cast<const uint2*>(par);     // P8: `par`/`s0` point at the int8 copies
  const uint2* s08 = reinterpret_cast<const uint2*>(s0);
  // chunk loaders: in P8 mode only .x/.y of the uint4 are used (8 bytes = 4 steps x 2 lanes)
  auto ldp = [&](int c) -> uint4 { if (P8) { uint2 v = __ldg(par8 + c * 4); return make_uint4(v.x, v.y, 0, 0); } return __ldg(par4 + c * 4); };
  auto ldz = [&](int c) -> uint4 { if (P8) { uint2 v = __ldg(s08 + c * 4); return make_uint4(v.x, v.y, 0, 0); } return __ldg(s04 + c * 4); };
  // value of step q inside a chunk register (sign-extending the two int8 to a packed int16 pair)
  auto pk = [&](const uint4& v, int q) -> u32 {
    if (P8) return prmt_sx(q < 2 ? v.x : v.y, (q & 1) ? 0xB3A2u : 0x9180u);
    return pick4(v, q);
  };
  // single-step loaders for the boundary code
  auto ldp1 = [&](int k) -> u32 {
    if (P8) { u32 w = reinterpret_cast<const uint16_t*>(par)[c4_word(k, 0)]; return prmt_sx(w, 0x9180u); }
    return __ldg(par + c4_word(k, 0));
  };
  auto ldz1 = [&](int k) -> u32 {
    if (P8) { u32 w = reinterpret_cast<const uint16_t*>(s0)[c4_word(k, 0)]; return prmt_sx(w, 0x9180u); }
    return __ldg(s0 + c4_word(k, 0));
  };
  auto c1 = [&](int k) -> FC { return fconst(__ldg(sys + c4_word(k, 0)), ldp1(k)); };
  auto d1 = [&](int k) -> u32 { return UPD ? __vsub2(ldz1(k), __ldg(sys + c4_word(k, 0))) : 0u; };   // s0 - sys
  u32 a[8];
  uint4 sb[NCH], pb[NCH], zb[NCH];

  // ---- forward sweep ------------------------------------------------------------------
#pragma unroll
  for (int s = 0; s < 8; ++s) a[s] = pack2(NEG_INIT, NEG_INIT);
  if (t == 0) a[0] = pack2(0, NEG_INIT);
#pragma unroll
  for (int j = 0; j < NCH; ++j)
    if (j < nchunk) { sb[j] = __ldg(sys4 + j * 4); pb[j] = ldp(j); }
  const int nfull = W / S;                                     // segments with all 16 steps
  for (int seg = 0; seg < nfull; ++seg) {
    if (PF > 0) {                                              // thread t warms L2 with chunk t of segment seg+PF
      const int cp = (seg + PF) * NCH + (t % NCH);
      if (cp < nchunk) { l2_prefetch(sys4 + cp * 4 - t); if (P8) l2_prefetch(par8 + cp * 4 - t); else l2_prefetch(par4 + cp * 4 - t); }
    }
    ckpt_put(ck + seg * 32, a);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (((j * 4 + q) & PM) == 0 && (j | q) != 0) renorm_ta<TRACK>(a, tk);
        alpha_fast(a, fconst(pick4(sb[j], q), pk(pb[j], q)));
      }
      const int cn = (seg + 1) * NCH + j;
      if (cn < nchunk) { sb[j] = __ldg(sys4 + cn * 4); pb[j] = ldp(cn); }
    }
    renorm_ta<TRACK>(a, tk);                                   // checkpoints are stored normalised
  }
  if (nfull < nseg) {                                          // partial last segment
    ckpt_put(ck + nfull * 32, a);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = j * 4 + q;
        if (nfull * S + e < W) {
          if ((e & PM) == 0 && e != 0) renorm_ta<TRACK>(a, tk);
          alpha_fast(a, fconst(pick4(sb[j], q), pk(pb[j], q)));
        }
      }
    }
    if (TRACK) renorm_ta<TRACK>(a, tk);                        // the final vector of the lane is recorded too
  }

  // ---- alpha re-run seed (kept in the checkpoint pool, slot nseg) -----------------------
  {
    u32 seed[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      u32 prev = __shfl_sync(gmask, a[s], (t + 3) & 3, 4);
      if (t == 0) prev = pack2(0, (s == 0) ? 0 : NEG_INIT);
      seed[s] = __byte_perm(prev, a[s], 0x5432);
    }
    ckpt_put(ck + nseg * 32, seed);
    if (W <= RERUN_STEPS) {
#pragma unroll
      for (int s = 0; s < 8; ++s) a[s] = seed[s];
      for (int k = 0; k < W; ++k) {
        if ((k & PM) == 0) renorm_ta<TRACK>(a, tk);
        alpha_fast(a, c1(k));
      }
      if (TRACK) renorm_ta<TRACK>(a, tk);
    }
  }

  // ---- beta start -------------------------------------------------------------------
  u32 b[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    b[s] = a[s];
    if (t == 3) b[s] = (a[s] & 0xffffu) | ((u32)(uint16_t)Tv[s] << 16);
  }
  if (TRACK) tk.sp_b = __vmaxu2(tk.sp_b, vec_spread(b));          // lane 7 starts from the tail metrics

  // ---- backward sweep, pass 1 -----------------------------------------------------------
  // loads the inputs of a whole segment into the chunk registers and its checkpoint into a[]
  auto load_segment = [&](int seg) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int c = seg * NCH + j;
      if (c < nchunk) {
        sb[j] = __ldg(sys4 + c * 4); pb[j] = ldp(c);
        if (UPD) zb[j] = ldz(c);
      }
    }
    ckpt_get(ck + seg * 32, a);
  };
  bool loaded = false;                             // sb/pb/zb/a hold segment `seg`
  for (int seg = nseg - 1; seg >= 0; --seg) {
    const int k0 = seg * S, k1 = min(W, k0 + S);
    // all 16 steps exist.  The last 6 steps of the lane take their LLR from the re-run below (which stores them after
    // this sweep, in program order), so a full last segment may run the steady-state code too: its 6 LLRs from pass-1 beta
    // are overwritten (and only make a tracked pass's recorded range a little more conservative).  Blocks with W a
    // multiple of 16 then have no boundary segment at all (K = 512: a quarter of the steps, at 3-4x the cost per step).
    const bool steady = (k0 + S <= W);
    if (PF > 0 && seg >= PF) {                    // warm L2 for segment seg-PF (inputs + checkpoint)
      const int cp = (seg - PF) * NCH + (t % NCH);
      l2_prefetch(sys4 + cp * 4 - t);
      if (P8) l2_prefetch(par8 + cp * 4 - t); else l2_prefetch(par4 + cp * 4 - t);
      if (UPD) { if (P8) l2_prefetch(s08 + cp * 4 - t); else l2_prefetch(s04 + cp * 4 - t); }
      if (t == 0) l2_prefetch(ck + (seg - PF) * 32);
    }
    if (!steady) {
      // boundary segment (the last one; every segment of a short block): two sub-segments of HS steps,
      // alpha of every step in shared memory, inputs re-read step by step
      for (int half = (S / HS) - 1; half >= 0; --half) {
        const int ka = k0 + half * HS, kb = min(k1, ka + HS);
        if (ka >= kb) continue;
        ckpt_get(ck + seg * 32, a);
        for (int k = k0; k < ka; ++k) {
          if (((k - k0) & PM) == 0 && k != k0) renorm(a);
          alpha_fast(a, c1(k));
        }
        for (int k = ka; k < kb; ++k) {
          if (((k - k0) & PM) == 0 && k != k0) renorm(a);
          sm.put(k - ka, tid, a);
          alpha_fast(a, c1(k));
        }
        if (seg == 0 && half == 0) {               // alpha[0..5] come from the re-run chain
          ckpt_get(ck + nseg * 32, a);
          for (int k = 0; k <= RERUN_STEPS && k < kb; ++k) {
            if ((k & PM) == 0) renorm_ta<TRACK>(a, tk);
            sm.put(k, tid, a);
            if (k < RERUN_STEPS) alpha_fast(a, c1(k));
          }
        }
        for (int k = kb - 1; k >= ka; --k) {
          const FC c = c1(k);
          if (k <= W - 7) {                        // steps whose beta[k+1] is not replaced by the re-run
            sm.get(k - ka, tid, a);
            u32 x = ext_fast(a, b, c);
            track_ext<TRACK>(x, tk);
            if (UPD) x = __vadd2(x, d1(k));
            ext[c4_word(k, 0)] = x;
          }
          beta_fast(b, c);
          if ((k & PM) == 0) renorm_tb<TRACK>(b, tk);
        }
      }
      loaded = false;
      continue;
    }
    if (!loaded) { load_segment(seg); loaded = true; }
    // ---- recompute alpha of the segment; even steps go to shared memory ----
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int e = j * 4 + q;
        if ((e & PM) == 0 && e != 0) renorm(a);
        if ((e & 1) == 0) sm.put(e >> 1, tid, a);
        if (e != S - 1) alpha_fast(a, fconst(pick4(sb[j], q), pk(pb[j], q)));
      }
    }
    if (seg == 0) {          // alpha[0..5] come from the re-run chain: entries 0, 2, 4 (1, 3, 5 follow from them)
      ckpt_get(ck + nseg * 32, a);
#pragma unroll
      for (int k = 0; k < RERUN_STEPS; ++k) {
        if ((k & PM) == 0) renorm_ta<TRACK>(a, tk);
        if ((k & 1) == 0) sm.put(k >> 1, tid, a);
        if (k < RERUN_STEPS - 1) alpha_fast(a, fconst(pick4(sb[k >> 2], k & 3), pk(pb[k >> 2], k & 3)));
      }
    }
    // next segment: checkpoint and chunk 0 (needed first, freed last -> spare registers) are requested now
    uint4 cka = make_uint4(0, 0, 0, 0), ckb = cka, sx = cka, px = cka, zx = cka;
    if (seg > 0) {
      cka = reinterpret_cast<const uint4*>(ck + (seg - 1) * 32)[0];
      ckb = reinterpret_cast<const uint4*>(ck + (seg - 1) * 32)[1];
      const int cn = (seg - 1) * NCH;
      sx = __ldg(sys4 + cn * 4); px = ldp(cn);
      if (UPD) zx = ldz(cn);
    }
    // ---- beta / ext sweep over the segment, two steps (even e, odd e+1) per round ----
    {
      u32 an[8];
      sm.get((S >> 1) - 1, tid, an);
#pragma unroll
      for (int j = NCH - 1; j >= 0; --j) {
        u32 e4[4];
#pragma unroll
        for (int h = 1; h >= 0; --h) {
          const int ee = j * 4 + 2 * h, eo = ee + 1;
          u32 ae[8], ao[8];
#pragma unroll
          for (int s = 0; s < 8; ++s) { ae[s] = an[s]; ao[s] = an[s]; }
          if (ee > 0) sm.get((ee >> 1) - 1, tid, an);          // in flight during this round's arithmetic
          const u32 sve = pick4(sb[j], 2 * h), svo = pick4(sb[j], 2 * h + 1);
          const FC ce = fconst(sve, pk(pb[j], 2 * h)), co = fconst(svo, pk(pb[j], 2 * h + 1));
          alpha_fast(ao, ce);                                   // alpha[eo] from alpha[ee]
          if (TRACK && seg == 0) tk.sp_a = __vmaxu2(tk.sp_a, vec_spread(ao));    // the odd entries of the re-run chain exist only here
          if ((eo & PM) == 0) renorm(ao);
          u32 xo = ext_fast(ao, b, co);
          track_ext<TRACK>(xo, tk);
          if (UPD) xo = __vadd2(xo, __vsub2(pk(zb[j], 2 * h + 1), svo));
          beta_fast(b, co);
          if ((eo & PM) == 0) renorm_tb<TRACK>(b, tk);
          u32 xe = ext_fast(ae, b, ce);
          track_ext<TRACK>(xe, tk);
          if (UPD) xe = __vadd2(xe, __vsub2(pk(zb[j], 2 * h), sve));
          beta_fast(b, ce);
          if ((ee & PM) == 0) renorm_tb<TRACK>(b, tk);
          e4[2 * h + 1] = xo; e4[2 * h] = xe;
        }
        *reinterpret_cast<uint4*>(ext + ((k0 >> 2) + j) * 16) = make_uint4(e4[0], e4[1], e4[2], e4[3]);
        if (seg > 0 && j > 0) {                                  // chunk j is consumed: refill with the next segment's
          const int cn = (seg - 1) * NCH + j;
          sb[j] = __ldg(sys4 + cn * 4); pb[j] = ldp(cn);
          if (UPD) zb[j] = ldz(cn);
        }
      }
    }
    sb[0] = sx; pb[0] = px; zb[0] = zx;
    a[0] = cka.x; a[1] = cka.y; a[2] = cka.z; a[3] = cka.w; a[4] = ckb.x; a[5] = ckb.y; a[6] = ckb.z; a[7] = ckb.w;
  }

  // ---- beta re-run over the last 5 steps, ext for the last 6 ------------------------------
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    u32 next = __shfl_sync(gmask, b[s], (t + 1) & 3, 4);
    if (t == 3) next = (u32)(uint16_t)Tv[s];
    b[s] = __byte_perm(b[s], next, 0x5432);
  }
  {
    const int kk0 = max(W - 6, 0);
    const int sa = kk0 / S;
    ckpt_get(ck + sa * 32, a);
    for (int k = sa * S; k < W; ++k) {
      if ((k & PM) == 0 && k != sa * S) renorm(a);
      if (k >= kk0) sm.put(k - kk0, tid, a);
      if (k + 1 < W) alpha_fast(a, c1(k));
    }
    if (kk0 <= RERUN_STEPS) {
      ckpt_get(ck + nseg * 32, a);
      for (int k = 0; k <= RERUN_STEPS && k < W; ++k) {
        if ((k & PM) == 0) renorm(a);
        if (k >= kk0) sm.put(k - kk0, tid, a);
        if (k < RERUN_STEPS) alpha_fast(a, c1(k));
      }
    }
    for (int k = W - 1; k >= kk0; --k) {
      const FC c = c1(k);
      sm.get(k - kk0, tid, a);
      u32 x = ext_fast(a, b, c);
      track_ext<TRACK>(x, tk);
      if (UPD) x = __vadd2(x, d1(k));
      ext[c4_word(k, 0)] = x;
      if (k >= W - RERUN_STEPS) {
        beta_fast(b, c);
        if ((k & PM) == 0) renorm_tb<TRACK>(b, tk);
      }
    }
  }
  if (TRACK) *trk = tk;
}

#ifndef MAP_MAX_REGS
#define MAP_MAX_REGS 168          // 6 CTAs (12 warps) per SM; measured against 128 / 144 / 184 / 200 / 216 / 243
#endif
// policy of one block for this pass: 16 / 4 / 1 = untracked fast pass with that renormalisation period (the a-priori guard
// holds), -1 = tracked fast pass (beyond the guard, range check a posteriori), 0 = exact saturating policy
constexpr int TRACK_CERT_LIMIT = 24000;   // of 32767: a tracked pass is attempted while the latest certificate (of either decoder) was below this
__device__ __forceinline__ int map_policy(const CbState& st, int guard_b, int track, int term) {
  const int B = max(st.max_sys, st.max_in) + st.max_in, M = B + 1;
  if (B <= guard_b) {
    // no wrap needs (11 + 2P) * M + 276 <= 32767  (DESIGN.md "fast-path guard")
    const int pmax = (32491 / M - 11) >> 1;
    if (pmax >= 1) return pmax >= 16 ? 16 : (pmax >= 4 ? 4 : 1);
  }
  // beyond the guard: try the fast path with the a-posteriori certificate unless the block has failed one before, or its
  // inputs alone (M >= 2 Gmax) leave no room for any spread
  if (track && !(st.retry & 2) && M <= 16000 && max(st.cert[0], st.cert[1]) <= TRACK_CERT_LIMIT) return -1;
  return 0;
}

template <int S>
__global__ void __maxnreg__(MAP_MAX_REGS) k_map16(MapArgs p) {
  extern __shared__ uint4 abuf[];
  if (p.retry && p.retry_flag && *p.retry_flag != p.seq) return;      // nothing failed in the launch before this one
  const int tid = threadIdx.x;
  const int gt = blockIdx.x * MAP_THREADS + tid;
  const int t = gt & 3;
  int blk = gt >> 2;
  const unsigned gmask = 0xFu << ((tid & 31) & ~3);
  if (p.active) {              // three classes, each padded to whole warps: fast from the front, tracked behind it, exact from the back
    const int nf = p.nactive[0], nt = p.nactive[1], nx = p.nactive[2], nfp = (nf + 7) & ~7, ntp = (nt + 7) & ~7;
    if (blk < nf) blk = p.active[blk];
    else if (blk >= nfp && blk - nfp < nt) blk = p.active[p.nblk + (blk - nfp)];          // (k_compact: tracked class at list[nblk + j])
    else if (blk >= nfp + ntp && blk - nfp - ntp < nx) blk = p.active[p.nblk - 1 - (blk - nfp - ntp)];
    else blk = p.nblk;
  }

  bool active = false;
  int W = 0, P = 64;           // P: see map_policy (64 = idle lane: does not constrain the warp)
  if (blk < p.nblk) {
    const CbMeta m = p.meta[blk];
    const CbState* st = &p.state[blk];
    active = (st->status == 0) && (m.flags & 1) && (p.iter <= m.max_iter);
    if (p.retry) active = active && (st->retry & 1);
    W = m.W;
    if (active) {
      P = p.retry ? 0 : map_policy(*st, p.guard_b, p.track, p.term);
      if (p.force == 1) P = 1; else if (p.force == 2) P = 0; else if (p.force == 3 && !p.retry) P = -1;
    }
  }
  // a warp carries 8 blocks and runs ONE policy: the most conservative of its blocks (exact < tracked < shorter period <
  // longer period; each is valid for every block that asked for a later one)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) P = min(P, __shfl_xor_sync(0xffffffffu, P, o));
  if (!active) return;

  int16_t* slot = p.ws + (long)blk * p.slot_hw;
  const u32* sys = reinterpret_cast<const u32*>(slot + (long)p.sys_arr * p.A) + t * 4;
  const u32* par = reinterpret_cast<const u32*>(slot + (long)p.par_arr * p.A) + t * 4;
  u32* ext = reinterpret_cast<u32*>(slot + (long)p.out_arr * p.A) + t * 4;
  u32* ck = p.ckpt + (long)blk * p.ckpt_words + t * 8;
  const int16_t* Tv = p.state[blk].T[p.term];

  unsigned char* smem = reinterpret_cast<unsigned char*>(abuf);
  const u32* s0 = reinterpret_cast<const u32*>(slot + (long)ARR_S0 * p.A) + t * 4;
  if (P == 0) {                                 // exact saturating policy, int16 arrays, 8-step segments
    map_pass<InvArith, MAP_ABUF_ENTRIES>(sys, par, s0, p.upd != 0, ext, ck, W, t, gmask, Tv, abuf, tid);
    if (p.retry && t == 0) p.state[blk].retry = 2;          // done; this block stays on the exact policy
    return;
  }
  if (P < 0) {                                  // tracked fast pass (int16 arrays, renormalisation every step)
    Trk tk;
    if (p.upd) map_pass_fast<S, 0, true, false, true>(sys, par, s0, ext, ck, W, t, gmask, Tv, smem, tid, &tk);
    else       map_pass_fast<S, 0, false, false, true>(sys, par, s0, ext, ck, W, t, gmask, Tv, smem, tid, &tk);
    // certificate of this block: maxima over its 4 threads (2 lanes each)
    int sa = max((int)(tk.sp_a & 0xffffu), (int)(tk.sp_a >> 16)), sb = max((int)(tk.sp_b & 0xffffu), (int)(tk.sp_b >> 16));
    int ex = max(max(lo16(tk.emx), hi16(tk.emx)), max(-lo16(tk.emn), -hi16(tk.emn)));
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      sa = max(sa, __shfl_xor_sync(gmask, sa, o, 4)); sb = max(sb, __shfl_xor_sync(gmask, sb, o, 4)); ex = max(ex, __shfl_xor_sync(gmask, ex, o, 4));
    }
    const CbState* st = &p.state[blk];
    const int Ms = st->max_sys, Mi = st->max_in, M = max(Ms, Mi) + Mi + 1;
    bool ok = (sa + sb + M <= 32767);
    if (p.upd) ok = ok && (ex + Ms + Mi <= 32767);
    if (t == 0) {
      if (p.term == 1) p.state[blk].max_ext2 = ok ? ex : -1;      // lets k_x2_16 skip its saturating forms (exact bound)
      p.state[blk].cert[p.term] = sa + sb + M;
      if (!ok) {                                    // repeat this pass on the exact policy (retry launch), and stay there
        p.state[blk].retry = 3;
        if (p.retry_flag) *p.retry_flag = p.seq;
      }
    }
    return;
  }
  // int8 copies of parity / s0 when the whole batch has |y| <= 127 (warp-uniform: one flag per batch)
  const bool p8 = p.batch_max && (*p.batch_max <= 127);
  if (p8) {
    const int8_t* b8a = reinterpret_cast<const int8_t*>(slot + (long)ARR_B8A * p.A);
    const int8_t* b8b = reinterpret_cast<const int8_t*>(slot + (long)ARR_B8B * p.A);
    par = reinterpret_cast<const u32*>(b8a + (p.par_arr == ARR_P2 ? p.A : 0) + t * 8);
    s0 = reinterpret_cast<const u32*>(b8b + t * 8);
  }
#define MAP_DISPATCH(PMV)                                                                                   \
  do {                                                                                                      \
    if (p.upd) { if (p8) map_pass_fast<S, PMV, true, true>(sys, par, s0, ext, ck, W, t, gmask, Tv, smem, tid);   \
                 else    map_pass_fast<S, PMV, true, false>(sys, par, s0, ext, ck, W, t, gmask, Tv, smem, tid); } \
    else       { if (p8) map_pass_fast<S, PMV, false, true>(sys, par, s0, ext, ck, W, t, gmask, Tv, smem, tid);  \
                 else    map_pass_fast<S, PMV, false, false>(sys, par, s0, ext, ck, W, t, gmask, Tv, smem, tid); } \
  } while (0)
  if (P >= 16)     MAP_DISPATCH(15);
  else if (P >= 4) MAP_DISPATCH(3);
  else             MAP_DISPATCH(0);
#undef MAP_DISPATCH
}

}  // namespace oai
