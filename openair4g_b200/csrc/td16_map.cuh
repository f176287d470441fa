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
#include "td_common.cuh"

namespace oai {

constexpr int MAP_THREADS = 128;     // 32 code blocks per CTA
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
  int guard_b;           // fast path allowed when max_sys + max_in <= guard_b
};

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
__device__ void map_pass(const u32* __restrict__ sys, const u32* __restrict__ par, u32* __restrict__ ext,
                         u32* ck, int W, int t, unsigned gmask, const int16_t* Tv, uint4* abuf, int tid) {
  const int nseg = (W + S - 1) / S;
  u32 a[8];

  // ---- forward sweep (alpha pass 1), checkpoint every S steps ----------------------
#pragma unroll
  for (int s = 0; s < 8; ++s) a[s] = pack2(NEG_INIT, NEG_INIT);
  if (t == 0) a[0] = pack2(0, NEG_INIT);                       // reference :201-208
  for (int c = 0; c * 4 < W; ++c) {
    uint4 s4 = __ldg(reinterpret_cast<const uint4*>(sys + c * 16));
    uint4 p4 = __ldg(reinterpret_cast<const uint4*>(par + c * 16));
    const u32 sv[4] = {s4.x, s4.y, s4.z, s4.w}, pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int k = c * 4 + q;
      if (k < W) {
        if (k % S == 0) ckpt_put(ck + (k / S) * 32, a);
        alpha_step<AR>(a, gamma2<AR>(sv[q], pv[q]));
      }
    }
  }

  // ---- alpha re-run seed: lane l <- final metrics of lane l-1, lane 0 <- start -----
  u32 seed[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    u32 prev = __shfl_sync(gmask, a[s], (t + 3) & 3, 4);      // thread t-1 (lanes 2t-2, 2t-1)
    if (t == 0) prev = pack2(0, (s == 0) ? 0 : NEG_INIT);      // hi half is what gets used
    seed[s] = __byte_perm(prev, a[s], 0x5432);                 // lo <- prev.hi, hi <- mine.lo
  }
  ckpt_put(ck + nseg * 32, seed);
  if (W <= RERUN_STEPS) {
    // the re-run covers the whole lane, so alpha[W] itself is the re-run value (K=40)
#pragma unroll
    for (int s = 0; s < 8; ++s) a[s] = seed[s];
    for (int k = 0; k < W; ++k)
      alpha_step<AR>(a, gamma2<AR>(__ldg(sys + c4_word(k, 0)), __ldg(par + c4_word(k, 0))));
  }

  // ---- beta start: lanes 0..6 <- own alpha[W], lane 7 <- tail metrics ---------------
  u32 b[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    b[s] = a[s];
    if (t == 3) b[s] = (a[s] & 0xffffu) | ((u32)(uint16_t)Tv[s] << 16);
  }

  // ---- backward sweep, pass 1 -----------------------------------------------------
  for (int seg = nseg - 1; seg >= 0; --seg) {
    const int k0 = seg * S, k1 = min(W, k0 + S);
    ckpt_get(ck + seg * 32, a);
    for (int k = k0; k < k1; ++k) {
      abuf_put(abuf, k - k0, tid, a);
      if (k + 1 < k1) alpha_step<AR>(a, gamma2<AR>(__ldg(sys + c4_word(k, 0)), __ldg(par + c4_word(k, 0))));
    }
    if (seg == 0) {          // alpha[0..5] come from the re-run chain
#pragma unroll
      for (int s = 0; s < 8; ++s) a[s] = seed[s];
      for (int k = 0; k <= RERUN_STEPS && k < k1; ++k) {
        abuf_put(abuf, k, tid, a);
        if (k < RERUN_STEPS) alpha_step<AR>(a, gamma2<AR>(__ldg(sys + c4_word(k, 0)), __ldg(par + c4_word(k, 0))));
      }
    }
    for (int k = k1 - 1; k >= k0; --k) {
      Gam<AR> g = gamma2<AR>(__ldg(sys + c4_word(k, 0)), __ldg(par + c4_word(k, 0)));
      if (k <= W - 7) {       // steps whose beta[k+1] is not replaced by the re-run
        abuf_get(abuf, k - k0, tid, a);
        ext[c4_word(k, 0)] = ext_step<AR>(a, b, g);
      }
      beta_step<AR>(b, g);
    }
  }

  // ---- beta re-run: lane l <- beta[0] of lane l+1, lane 7 <- tail metrics ------------
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    u32 next = __shfl_sync(gmask, b[s], (t + 1) & 3, 4);      // thread t+1 (lanes 2t+2, 2t+3)
    if (t == 3) next = (u32)(uint16_t)Tv[s];
    b[s] = __byte_perm(b[s], next, 0x5432);                    // lo <- mine.hi, hi <- next.lo
  }
  {
    const int kk0 = max(W - 6, 0);
    const int sa = kk0 / S;
    ckpt_get(ck + sa * 32, a);
    for (int k = sa * S; k < W; ++k) {
      if (k >= kk0) abuf_put(abuf, k - kk0, tid, a);
      if (k + 1 < W) alpha_step<AR>(a, gamma2<AR>(__ldg(sys + c4_word(k, 0)), __ldg(par + c4_word(k, 0))));
    }
    if (kk0 <= RERUN_STEPS) {
#pragma unroll
      for (int s = 0; s < 8; ++s) a[s] = seed[s];
      for (int k = 0; k <= RERUN_STEPS && k < W; ++k) {
        if (k >= kk0) abuf_put(abuf, k - kk0, tid, a);
        if (k < RERUN_STEPS) alpha_step<AR>(a, gamma2<AR>(__ldg(sys + c4_word(k, 0)), __ldg(par + c4_word(k, 0))));
      }
    }
    for (int k = W - 1; k >= kk0; --k) {
      Gam<AR> g = gamma2<AR>(__ldg(sys + c4_word(k, 0)), __ldg(par + c4_word(k, 0)));
      abuf_get(abuf, k - kk0, tid, a);
      ext[c4_word(k, 0)] = ext_step<AR>(a, b, g);
      if (k >= W - RERUN_STEPS) beta_step<AR>(b, g);             // loopval=(n-40)>>3, reference :585
    }
  }
}

template <int S>
__global__ void __launch_bounds__(MAP_THREADS) k_map16(MapArgs p) {
  extern __shared__ uint4 abuf[];
  const int tid = threadIdx.x;
  const int gt = blockIdx.x * MAP_THREADS + tid;
  const int blk = gt >> 2, t = gt & 3;
  const unsigned gmask = 0xFu << ((tid & 31) & ~3);

  bool active = false, fast = true;
  int W = 0;
  if (blk < p.nblk) {
    const CbMeta m = p.meta[blk];
    const CbState* st = &p.state[blk];
    active = (st->status == 0) && (m.flags & 1) && (p.iter <= m.max_iter);
    W = m.W;
    if (active) fast = (st->max_sys + st->max_in) <= p.guard_b;
  }
  // a warp carries 8 blocks; the exact (saturating) policy is always valid, so one hot
  // block switches its whole warp to it
  const bool warp_fast = __all_sync(0xffffffffu, fast);
  if (!active) return;

  int16_t* slot = p.ws + (long)blk * p.slot_hw;
  const u32* sys = reinterpret_cast<const u32*>(slot + (long)p.sys_arr * p.A) + t * 4;
  const u32* par = reinterpret_cast<const u32*>(slot + (long)p.par_arr * p.A) + t * 4;
  u32* ext = reinterpret_cast<u32*>(slot + (long)p.out_arr * p.A) + t * 4;
  u32* ck = p.ckpt + (long)blk * p.ckpt_words + t * 8;
  const int16_t* Tv = p.state[blk].T[p.term];

  if (warp_fast) map_pass<WrapArith, S>(sys, par, ext, ck, W, t, gmask, Tv, abuf, tid);
  else           map_pass<SatArith, S>(sys, par, ext, ck, W, t, gmask, Tv, abuf, tid);
}

}  // namespace oai
