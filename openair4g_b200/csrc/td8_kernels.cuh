// 8-bit max-log-MAP turbo decoder (reference: openair1/PHY/CODING/3gpplte_turbo_decoder_sse_8bit.c),
// first GPU version: a direct, always-saturating restatement.  int8 arithmetic saturates
// routinely (metrics live in [-128,127]), so there is no non-saturating fast path as in the
// 16-bit kernel; every add/sub is clamped like _mm_adds_epi8/_mm_subs_epi8.
//
// Mapping: 16 threads per code block, thread l = the reference's SIMD lane l (trellis
// positions [l*W,(l+1)*W), W = n/16, TD8:874-881); 2 blocks per warp.  alpha and beta are
// stored for every step in HBM exactly like the reference's stack arrays (TD8:928-929), in
// the reference layout [(step*8+state)*16 + lane] so that the 16 lanes of a block read and
// write 16 consecutive bytes; the boundary re-seeds (TD8:299-316, 652-666) read the
// neighbouring lane's metrics from those arrays.  Per-position arrays are int8 in the
// reference lane layout st8(p) = (p mod W)*16 + p/W.
//
//   k_demux8 : input scaling int16 -> int8 (TD8:1000-1029) and demux (TD8:1062-1077)
//   k_map8   : log_map8 = gamma/alpha/beta/ext (TD8:95-149, 151-827)
//   k_x1_8   : feedback ext = (ext (-) s1) (+) s0 (TD8:1632-1653) and gather s2 = ext o pi (TD8:1341-1379)
//   k_x2_8   : s1 = (ext2 o pi^-1 (-) ext) (+) s0, hard decision (two rules, TD8:1392-1581), CRC, exit
// Parity domain n >= 256, n % 16 == 0; the tail LLRs never influence the
// reference's output and are not read.
#pragma once
#include "td_common.cuh"
#include "td16_xchg.cuh"

namespace oai {

constexpr int MAP8_THREADS = 128;     // 8 code blocks per CTA
constexpr int INIT8 = -63;            // -MAX8/2 (TD8:92,234)
constexpr int RERUN8 = 16;            // L (TD8:211)

enum { A8_S0 = 0, A8_P1 = 1, A8_P2 = 2, A8_SYS = 3, A8_EXT = 4, A8_EXT2 = 5, A8_COUNT = 6 };

struct Td8Args {
  const CbMeta* meta;
  CbState* state;
  int8_t* ws;            // per block: A8_COUNT arrays of `A` bytes
  long slot_b;
  int A;                 // bytes per array (>= n, multiple of 16)
  int8_t* ab;            // per block: alpha then beta, each 128*(W+1) bytes
  long ab_b;             // bytes per block in `ab`
  int nblk;
  const uint16_t* qpp;   // plain QPP tables pi[i]
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
    const int h = ((pos % W) << 4) + pos / W;
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
  for (int i = threadIdx.x; i < n / 16; i += XCHG_THREADS) {
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
// one trellis step of the forward recursion for one lane (TD8:251-297)
__device__ __forceinline__ void alpha8_step(int (&a)[8], int g1, int g0) {
  int n0 = max(s8(a[1] + g1), s8(a[0] - g1));
  int n1 = max(s8(a[3] - g0), s8(a[2] + g0));
  int n2 = max(s8(a[5] + g0), s8(a[4] - g0));
  int n3 = max(s8(a[7] - g1), s8(a[6] + g1));
  int n4 = max(s8(a[1] - g1), s8(a[0] + g1));
  int n5 = max(s8(a[3] + g0), s8(a[2] - g0));
  int n6 = max(s8(a[5] - g0), s8(a[4] + g0));
  int n7 = max(s8(a[7] + g1), s8(a[6] - g1));
  const int mx = max(max(max(n0, n1), max(n2, n3)), max(max(n4, n5), max(n6, n7)));
  a[0] = s8(n0 - mx); a[1] = s8(n1 - mx); a[2] = s8(n2 - mx); a[3] = s8(n3 - mx);
  a[4] = s8(n4 - mx); a[5] = s8(n5 - mx); a[6] = s8(n6 - mx); a[7] = s8(n7 - mx);
}
// backward recursion (TD8:579-650)
__device__ __forceinline__ void beta8_step(int (&b)[8], int g1, int g0) {
  int n0 = max(s8(b[4] + g1), s8(b[0] - g1));
  int n1 = max(s8(b[4] - g1), s8(b[0] + g1));
  int n2 = max(s8(b[5] - g0), s8(b[1] + g0));
  int n3 = max(s8(b[5] + g0), s8(b[1] - g0));
  int n4 = max(s8(b[6] + g0), s8(b[2] - g0));
  int n5 = max(s8(b[6] - g0), s8(b[2] + g0));
  int n6 = max(s8(b[7] - g1), s8(b[3] + g1));
  int n7 = max(s8(b[7] + g1), s8(b[3] - g1));
  const int mx = max(max(max(n0, n1), max(n2, n3)), max(max(n4, n5), max(n6, n7)));
  b[0] = s8(n0 - mx); b[1] = s8(n1 - mx); b[2] = s8(n2 - mx); b[3] = s8(n3 - mx);
  b[4] = s8(n4 - mx); b[5] = s8(n5 - mx); b[6] = s8(n6 - mx); b[7] = s8(n7 - mx);
}

__global__ void __launch_bounds__(MAP8_THREADS) k_map8(Td8Args p) {
  const int gt = blockIdx.x * MAP8_THREADS + threadIdx.x;
  const int blk = gt >> 4, l = gt & 15;
  const unsigned hmask = 0xffffu << (threadIdx.x & 16);        // the 16 threads of this block
  if (blk >= p.nblk) return;
  const CbMeta m = p.meta[blk];
  const CbState* st = &p.state[blk];
  if (st->status != 0 || !(m.flags & 1) || p.iter > m.max_iter) return;
  const int n = m.K, W = n >> 4;
  const int8_t* slot = p.ws + (long)blk * p.slot_b;
  const int8_t* sys = slot + (long)p.sys_arr * p.A + l;
  const int8_t* par = slot + (long)p.par_arr * p.A + l;
  int8_t* ext = const_cast<int8_t*>(slot) + (long)p.out_arr * p.A + l;
  int8_t* alpha = p.ab + (long)blk * p.ab_b + l;               // element (k,s) at (k*8+s)*16
  int8_t* beta = alpha + 128 * (W + 1);
  auto G1 = [&](int k) { return ((int)sys[k * 16] + (int)par[k * 16]) >> 1; };   // TD8:178-185, exact halves
  auto G0 = [&](int k) { return ((int)sys[k * 16] - (int)par[k * 16]) >> 1; };
  int a[8], b[8];

  // ---- alpha: init, W steps, re-seed, 16 steps, re-seed (TD8:234-316) -----------------------
#pragma unroll
  for (int s = 0; s < 8; ++s) a[s] = (l == 0 && s == 0) ? 0 : INIT8;
#pragma unroll
  for (int s = 0; s < 8; ++s) alpha[s * 16] = (int8_t)a[s];
  for (int k = 0; k < W; ++k) {
    alpha8_step(a, G1(k), G0(k));
#pragma unroll
    for (int s = 0; s < 8; ++s) alpha[((k + 1) * 8 + s) * 16] = (int8_t)a[s];
  }
  for (int pass = 0; pass < 2; ++pass) {
    // re-seed: lane l <- alpha[W] of lane l-1, lane 0 <- (0,-63,...): `a` holds this lane's alpha[W]
    int seed[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int prev = __shfl_up_sync(hmask, a[s], 1, 16);
      seed[s] = (l == 0) ? (s == 0 ? 0 : INIT8) : prev;
      alpha[s * 16] = (int8_t)seed[s];
    }
    if (pass == 1) break;
#pragma unroll
    for (int s = 0; s < 8; ++s) b[s] = seed[s];                 // b: scratch for the re-run chain
    for (int k = 0; k < RERUN8; ++k) {
      alpha8_step(b, G1(k), G0(k));
#pragma unroll
      for (int s = 0; s < 8; ++s) alpha[((k + 1) * 8 + s) * 16] = (int8_t)b[s];
    }
    if (W == RERUN8) {                                          // the re-run reached alpha[W] (K=256)
#pragma unroll
      for (int s = 0; s < 8; ++s) a[s] = b[s];
    }
  }

  // ---- beta: from alpha[W]; lane 15 <- 0 before each pass; shift re-seed after each (TD8:505-666) ----
#pragma unroll
  for (int s = 0; s < 8; ++s) b[s] = (l == 15) ? 0 : a[s];
  int b0[8];                                                    // beta[0] of the latest pass that reached step 0
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int s = 0; s < 8; ++s) beta[(W * 8 + s) * 16] = (int8_t)b[s];
    const int kend = (pass == 0) ? 0 : W - RERUN8;
    for (int k = W - 1; k >= kend; --k) {
      beta8_step(b, G1(k), G0(k));
#pragma unroll
      for (int s = 0; s < 8; ++s) beta[(k * 8 + s) * 16] = (int8_t)b[s];
    }
    if (kend == 0) {
#pragma unroll
      for (int s = 0; s < 8; ++s) b0[s] = b[s];
    }
    // re-seed beta[W]: lane l <- beta[0] of lane l+1, lane 15 <- 0
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int next = __shfl_down_sync(hmask, b0[s], 1, 16);
      b[s] = (l == 15) ? 0 : next;
    }
  }
#pragma unroll
  for (int s = 0; s < 8; ++s) beta[(W * 8 + s) * 16] = (int8_t)b[s];
  __syncwarp(hmask);

  // ---- ext (TD8:715-770): alpha[k], beta[k+1]; each lane reads back only its own column ----
  for (int k = 0; k < W; ++k) {
#pragma unroll
    for (int s = 0; s < 8; ++s) { a[s] = alpha[(k * 8 + s) * 16]; b[s] = beta[((k + 1) * 8 + s) * 16]; }
    const int g1 = G1(k), g0 = G0(k);
    int m00 = max(max(s8(a[0] + b[0]), s8(a[1] + b[4])), max(s8(a[6] + b[7]), s8(a[7] + b[3])));
    int m11 = max(max(s8(a[0] + b[4]), s8(a[1] + b[0])), max(s8(a[6] + b[3]), s8(a[7] + b[7])));
    int m01 = max(max(s8(a[2] + b[5]), s8(a[3] + b[1])), max(s8(a[4] + b[2]), s8(a[5] + b[6])));
    int m10 = max(max(s8(a[2] + b[1]), s8(a[3] + b[5])), max(s8(a[4] + b[6]), s8(a[5] + b[2])));
    m01 = s8(m01 - g0); m00 = s8(m00 - g1); m10 = s8(m10 + g0); m11 = s8(m11 + g1);
    ext[k * 16] = (int8_t)s8(max(m10, m11) - max(m01, m00));
  }
}

// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(XCHG_THREADS) k_x1_8(Td8Args p) {
  extern __shared__ int8_t sm8[];
  const int blk = blockIdx.x;
  if (blk >= p.nblk) return;
  const CbMeta m = p.meta[blk];
  if (p.state[blk].status != 0 || p.iter > m.max_iter) return;
  const int n = m.K, W = n >> 4, A = p.A;
  int8_t* slot = p.ws + (long)blk * p.slot_b;
  int8_t* gext = slot + (long)A8_EXT * A, *gsys = slot + (long)A8_SYS * A;
  const int8_t* gs0 = slot + (long)A8_S0 * A;
  for (int i = threadIdx.x; i < n; i += XCHG_THREADS) {
    int e = gext[i];
    if (p.iter > 1) {                 // ext = (ext (-) s1) (+) s0, TD8:1632-1653
      e = s8(s8(e - gsys[i]) + gs0[i]);
      gext[i] = (int8_t)e;
    }
    sm8[i] = (int8_t)e;
  }
  __syncthreads();
  const uint16_t* pi = p.qpp + m.pi_off;
  for (int i = threadIdx.x; i < n; i += XCHG_THREADS) {         // s2[st8(i)] = ext[st8(pi(i))], TD8:1341-1379
    const int j = pi[i];
    gsys[((i % W) << 4) + i / W] = sm8[((j % W) << 4) + j / W];
  }
}

__global__ void __launch_bounds__(XCHG_THREADS) k_x2_8(Td8Args p) {
  extern __shared__ int8_t sm8[];
  __shared__ u32 xred[2 * XCHG_THREADS / 32];
  __shared__ __align__(16) uint8_t sbytes[768 + 32];
  const int blk = blockIdx.x;
  if (blk >= p.nblk) return;
  const CbMeta m = p.meta[blk];
  CbState* st = &p.state[blk];
  if (st->status != 0 || p.iter > m.max_iter) return;
  const int n = m.K, W = n >> 4, A = p.A;
  int8_t* slot = p.ws + (long)blk * p.slot_b;
  const int8_t* gext2 = slot + (long)A8_EXT2 * A, *gext = slot + (long)A8_EXT * A, *gs0 = slot + (long)A8_S0 * A;
  int8_t* gsys = slot + (long)A8_SYS * A;
  int8_t* e2 = sm8, *dec = sm8 + A;           // ext2 (interleaved order); decision variable (same order)
  const bool mode1 = (n & 0x7f) == 0;         // TD8:1392 / 1488
  for (int i = threadIdx.x; i < n; i += XCHG_THREADS) {
    const int v = gext2[i];
    e2[i] = (int8_t)v;
    dec[i] = (int8_t)(mode1 ? v : s8(v + gsys[i]));             // ext2 (+) sys2, TD8:1456
  }
  __syncthreads();
  const uint16_t* pi = p.qpp + m.pi_off;
  for (int i = threadIdx.x; i < n; i += XCHG_THREADS) {         // s1 = (ext2 o pi^-1 (-) ext) (+) s0, TD8:1392-1459
    const int pj = pi[i];
    const int j = ((pj % W) << 4) + pj / W, hi = ((i % W) << 4) + i / W;
    gsys[j] = (int8_t)s8(s8((int)e2[hi] - gext[j]) + gs0[j]);
  }
  bool pass = false;
  if (p.iter > 1) {
    // hard decision at natural position pi(i): written bit by bit through shared-memory atomics
    // would be slow; instead each warp ballots 32 consecutive NATURAL positions, which needs the
    // inverse permutation -- obtained by scattering the decision variable to natural order first
    int8_t* natdec = e2;                       // reuse (e2 no longer needed after the barrier below)
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += XCHG_THREADS) natdec[pi[i]] = dec[((i % W) << 4) + i / W];
    __syncthreads();
    u32 word = 0;                              // thread w packs natural positions 32w..32w+31, MSB first
    if ((int)threadIdx.x < ((n >> 3) + 3) >> 2) {
      const int j0 = threadIdx.x << 5;
      u32 bits = 0;
      for (int q = 0; q < 32; ++q) bits = (bits << 1) | ((j0 + q < n && natdec[j0 + q] > 0) ? 1u : 0u);
      word = bits;
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
