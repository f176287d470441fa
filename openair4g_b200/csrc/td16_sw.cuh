// OPTIONAL sliding-window mode of the 16-bit decoder (OAI_BATCH_SLIDING_WINDOW; north_star: "an optional sliding-window,
// higher-parallelism mode is allowed only if it is separately reported with its BLER delta against the bit-exact mode").
// NOT bit-exact with the reference: never selected unless the caller asks for it.
//
// The bit-exact path keeps the reference's split of a block into 8 lanes (it defines the result), i.e. 4 threads per
// block, >= 10^4 blocks in flight and all per-position state streamed through HBM once per constituent pass (1.28 MB of
// traffic per K=6144 block and decode).  This mode splits a block into NW = 8 / 16 / 32 / 64 windows (K < 512 / < 1024 /
// < 2048 / >= 2048; window length 5...96), two windows per thread, so ONE WARP decodes one K >= 2048 block (or 2 / 4 / 8
// shorter blocks of equal K) from the soft bits to the CRC verdict with everything resident in shared memory:
//   HBM traffic per block = its 3K+12 soft bits in, K/8 bytes out (K=6144: 37.6 KB instead of 1.28 MB),
//   one launch per batch instead of 31, and a single block takes 0.10-0.31 ms instead of 0.23-1.07 ms (K = 40 ... 6144).
// Windows are stitched two ways at once: a window's forward (backward) recursion starts from the more confident (larger
// max - min) of (a) the metrics its left (right) neighbour finished with in the previous iteration ("next-iteration
// initialisation") and (b) the result of a 32-step training recursion over the neighbour's last (first) steps on the
// current inputs, started from equal metrics.  (a) alone lags one iteration behind at every window edge (+0.35 iterations
// on average at rate 1/3), (b) alone is too short for heavily punctured blocks; the pair needs fewer iterations and loses
// fewer blocks than the reference's 8 lanes with their 5-step re-run.  Window 0 starts in state 0 and the last window from
// the tail bits (TD16:474-520), like the reference's lanes 0 and 7.
//
// Arithmetic: the soft bits are scaled to 8 bits (right shift chosen from the block's mean |y|, then clipped to +-127)
// and the extrinsic values are clipped to +-SW_LC = 767, so a systematic input is within +-894, M = 894 + 127 + 1 = 1022
// bounds 2 Gmax and the shifted branch metrics, and the non-saturating DPX fast path of td16_map.cuh (fconst / alpha_fast
// / beta_fast / ext_fast, renormalised every P = 8 steps) cannot wrap: (11 + 2P) M + 276 = 27 870 <= 32 767 (DESIGN.md
// "fast-path guard" (b); the start vectors -- previous final vectors, training results, (0,-3000,...) for window 0, the tail
// metrics -- have spreads within the 10 Gmax + 128 the bound assumes).  A model in plain int arithmetic lives with the test infrastructure;
// this kernel must agree with it bit for bit (tests/test_gpu_sw.py).
//
// Shared memory per warp, window length WL (all arrays [step][lane], one 32-bit word = the thread's two windows):
//   S0, P1, P2 (int8 pairs) 3 x 64 WL | A, B (systematic input / output of the running pass, int16 pairs) 2 x 128 WL |
//   alpha checkpoints every 16 steps 1 KB per segment | window start metrics 4 KB | decoded bits
//   = 54.3 KB at WL = 96 -> 4 warps (blocks) per SM.
#pragma once
#include "td16_map.cuh"
#include "td16_xchg.cuh"

namespace oai {

constexpr int SW_LC = 767;           // extrinsic clip
constexpr int SW_Q = 3000;           // penalty of the states window 0 cannot start in
constexpr int SW_TRAIN = 32;         // steps of the training recursions in front of / behind a window
constexpr int SW_XB = 4;             // window steps per batch of the exchange loops
constexpr int SW_BITS_WORDS = 192;   // decoded bits of the warp's blocks (K/32 words each)

__host__ __device__ inline int sw_windows(int K) { return K >= 2048 ? 64 : (K >= 1024 ? 32 : (K >= 512 ? 16 : 8)); }
__host__ __device__ inline int sw_smem_bytes(int K) {
  const int WL = K / sw_windows(K), nseg = (WL + 15) >> 4;
  return WL * (128 * 3 + 64) + nseg * 1024 + 4096 + SW_BITS_WORDS * 4 + 256;
}

struct SwArgs {
  const CbMeta* meta;
  CbState* state;
  int nblk;
  const int16_t* in_base;
  uint8_t* out_base;
  uint8_t* status_out;
  const u32* tab_pool;      // per K: [WL][NW/2] packed pairs of shared-memory halfword indices of pi(j) for the thread's two windows
  const u32* tab_off;       // [K >> 3] word offset of K's table
  const u32* crc_xp;
};

__device__ __forceinline__ int sw_scale(int v, int sh) { return max(-127, min(127, v >> sh)); }

// shared-memory word of (step o, lane ln): rotated by the step so that a walk along one lane's steps (the fill of the
// arrays) and a walk across the lanes of one step (the recursions, the exchange) are both free of bank conflicts
__device__ __forceinline__ int sw_idx(int o, int ln) { return (o << 5) + ((ln + o) & 31); }

// parity pair of step-word idx; PH points at the int8 pairs of the running decoder, so that ONE copy of the recursion code
// serves both decoders (two copies do not fit the instruction cache once the warps of an SM run out of phase)
__device__ __forceinline__ u32 sw_par(const uint16_t* __restrict__ PH, int idx) { return prmt_sx((u32)PH[idx], 0x9180u); }

// forward recursion over one segment of n <= 16 steps starting at step `base` (FULL: n == 16, no per-step tests, so
// that the instruction scheduler can overlap consecutive steps)
template <bool FULL>
__device__ __forceinline__ void sw_fwd_seg(u32 (&a)[8], const u32* __restrict__ IN, const uint16_t* __restrict__ PH,
                                           int base, int n, int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (FULL || i < n) {
      const int idx = sw_idx(base + i, lane);
      const FC c = fconst(IN[idx], sw_par(PH, idx));
      alpha_fast(a, c);
      if ((i & 7) == 7) renorm(a);
    }
  }
}

// backward recursion + LLRs over one segment: a = alpha in front of the segment (checkpoint), b = beta behind it
//   OUT = 2 * clip(LLR - IN) | (LLR > 0)      (extrinsic value and hard decision; the exchange steps add s0)
template <bool FULL>
__device__ __forceinline__ void sw_bwd_seg(u32 (&a)[8], u32 (&b)[8], const u32* __restrict__ IN, u32* __restrict__ OUT,
                                           const uint16_t* __restrict__ PH, int base, int n, int lane) {
  u32 ae[8][8];
  // recompute the alpha vectors in front of the even steps (the odd ones follow from them in the sweep below)
#pragma unroll
  for (int i = 0; i < 15; ++i) {
    if (FULL || i < n) {
      if (!(i & 1)) {
#pragma unroll
        for (int s = 0; s < 8; ++s) ae[i >> 1][s] = a[s];
      }
      if (i < 14) {                                  // (the vector in front of step 14 is the last one needed)
        const int idx = sw_idx(base + i, lane);
        const FC c = fconst(IN[idx], sw_par(PH, idx));
        alpha_fast(a, c);
        if ((i & 7) == 7) renorm(a);
      }
    }
  }
  // the inputs are read again here instead of being kept in 32 registers across the recomputation
  FC cn;                                             // constants of step i - 1, shared by the odd step i and step i - 1
  u32 xn = 0;
#pragma unroll
  for (int i = 15; i >= 0; --i) {
    if (FULL || i < n) {
      FC c;
      u32 xi;
      if ((i & 1) || !(FULL || i + 1 < n)) {         // not prepared by the step above
        const int idx = sw_idx(base + i, lane);
        xi = IN[idx];
        c = fconst(xi, sw_par(PH, idx));
      } else {
        c = cn; xi = xn;
      }
      u32 ai[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) ai[s] = ae[i >> 1][s];
      if (i & 1) {
        const int idp = sw_idx(base + i - 1, lane);
        xn = IN[idp];
        cn = fconst(xn, sw_par(PH, idp));
        alpha_fast(ai, cn);
      }
      const u32 e = ext_fast(ai, b, c);
      beta_fast(b, c);
      if ((i & 7) == 0) renorm(b);
      u32 le = __vsub2(e, xi);
      le = __vmins2(__vmaxs2(le, 0xFD01FD01u), 0x02FF02FFu);            // +-SW_LC
      OUT[sw_idx(base + i, lane)] = __vadd2(le, le) | (__vcmpgts2(e, 0u) & 0x00010001u);
    }
  }
}

// One constituent pass of the warp's blocks.  IN / OUT: [WL][32] int16 pairs; PH: the decoder's parity (sw_par); nii: [alpha | beta][8 states][32] start metrics of this decoder, updated for the next
// iteration; term: [8 states][8 blocks] tail metrics of this decoder.  The window is cut into a first segment of
// n0 = WL - 16 (nseg - 1) steps and full 16-step segments.
__device__ __noinline__ void sw_pass(const u32* __restrict__ IN, u32* __restrict__ OUT, const uint16_t* __restrict__ PH,
                                     u32* __restrict__ CK, u32* __restrict__ nii, const int16_t* __restrict__ term,
                                     int WL, int lane, int tl, int LPB, int g) {
  const unsigned FULL = 0xffffffffu;
  const int nseg = (WL + 15) >> 4, n0 = WL - ((nseg - 1) << 4);
  const int L = min(SW_TRAIN, WL);
  u32 a[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) a[s] = nii[s * 32 + lane];
  {
    // training recursion over the last L steps of the left neighbours (window 2t-1: upper halfword of lane t-1; window
    // 2t: own lower halfword) from equal metrics; the more confident of (training result, last iteration's final
    // metrics of the neighbour) starts the window.  Window 0 starts in state 0.
    u32 t[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    const int lp = (tl == 0) ? lane : lane - 1;
    int cnt = 0;
#pragma unroll 4
    for (int q = WL - L; q < WL; ++q) {
      const int io = sw_idx(q, lane), ip = sw_idx(q, lp);
      const FC c = fconst(__byte_perm(IN[ip], IN[io], 0x5432), prmt_sx((u32)PH[ip] | ((u32)PH[io] << 16), 0xA291u));
      alpha_fast(t, c);
      if ((++cnt & 7) == 0) renorm(t);
    }
    renorm(t);
    u32 m = __vcmpgeu2(vec_spread(t), vec_spread(a));
    if (tl == 0) m &= 0xffff0000u;
#pragma unroll
    for (int s = 0; s < 8; ++s) a[s] = (t[s] & m) | (a[s] & ~m);
  }
#pragma unroll
  for (int s = 0; s < 8; ++s) CK[s * 32 + lane] = a[s];
  if (n0 == 16) sw_fwd_seg<true>(a, IN, PH, 0, 16, lane);
  else sw_fwd_seg<false>(a, IN, PH, 0, n0, lane);
  for (int seg = 1; seg < nseg; ++seg) {
    renorm(a);
#pragma unroll
    for (int s = 0; s < 8; ++s) CK[(seg * 8 + s) * 32 + lane] = a[s];
    sw_fwd_seg<true>(a, IN, PH, n0 + ((seg - 1) << 4), 16, lane);
  }
  renorm(a);
  // the final metrics of window w start window w+1 in the next iteration
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const u32 prev = __shfl_up_sync(FULL, a[s], 1);
    nii[s * 32 + lane] = (tl == 0) ? pack2(s ? -SW_Q : 0, lo16(a[s])) : __byte_perm(prev, a[s], 0x5432);
  }

  u32 b[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) b[s] = nii[(8 + s) * 32 + lane];
  {
    // ... over the first L steps of the right neighbours (window 2t+1: own upper halfword; window 2t+2: lower halfword of
    // lane t+1); the last window starts from the tail bits
    u32 t[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    const int ln = (tl == LPB - 1) ? lane : lane + 1;
    int cnt = 0;
#pragma unroll 4
    for (int q = L - 1; q >= 0; --q) {
      const int io = sw_idx(q, lane), in = sw_idx(q, ln);
      const FC c = fconst(__byte_perm(IN[io], IN[in], 0x5432), prmt_sx((u32)PH[io] | ((u32)PH[in] << 16), 0xA291u));
      beta_fast(t, c);
      if ((++cnt & 7) == 0) renorm(t);
    }
    renorm(t);
    u32 m = __vcmpgeu2(vec_spread(t), vec_spread(b));
    if (tl == LPB - 1) m &= 0x0000ffffu;
#pragma unroll
    for (int s = 0; s < 8; ++s) b[s] = (t[s] & m) | (b[s] & ~m);
  }
  for (int seg = nseg - 1; seg >= 1; --seg) {
#pragma unroll
    for (int s = 0; s < 8; ++s) a[s] = CK[(seg * 8 + s) * 32 + lane];
    sw_bwd_seg<true>(a, b, IN, OUT, PH, n0 + ((seg - 1) << 4), 16, lane);
  }
#pragma unroll
  for (int s = 0; s < 8; ++s) a[s] = CK[s * 32 + lane];
  if (n0 == 16) sw_bwd_seg<true>(a, b, IN, OUT, PH, 0, 16, lane);
  else sw_bwd_seg<false>(a, b, IN, OUT, PH, 0, n0, lane);
  renorm(b);
  // the metrics at the start of window w start window w-1's backward recursion in the next iteration
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const u32 nxt = __shfl_down_sync(FULL, b[s], 1);
    nii[(8 + s) * 32 + lane] = (tl == LPB - 1) ? pack2(hi16(b[s]), term[s * 8 + g]) : __byte_perm(b[s], nxt, 0x5432);
  }
}

// Exchange steps, batches of SW_XB window steps so that the table reads and the dependent shared-memory gathers of a batch
// overlap (one warp per scheduler: no other warp hides their latency).  Table entries are BYTE offsets into the int16
// arrays (halfword index x 2); MULTI: the warp carries several blocks, block g adds g * LPB to the lane field (mod 32).
template <bool MULTI>
__device__ __forceinline__ u32 sw_tab(const u32* __restrict__ tab, int i, u32 gadd) {
  const u32 v = __ldg(tab + i);
  return MULTI ? ((v & 0xFF83FF83u) | ((v + gadd) & 0x007C007Cu)) : v;
}

// second decoder's systematic input = (s0 + extrinsic) o pi  (TD16:1209-1231, 1354-1375).  The table entries of batch
// i + 1 are requested before batch i is processed: with 217 KB of the SM's 256 KB configured as shared memory the tables
// do not stay in L1, and an L2 round trip per batch was 40 % of this loop's time.
template <bool MULTI>
__device__ __forceinline__ void sw_x1(u32* __restrict__ Aw, const unsigned char* __restrict__ Bb, const signed char* __restrict__ S0B,
                                      const u32* __restrict__ tab, u32 gadd, int WL, int LPB, int lane, int tl) {
  u32 tn[SW_XB];
#pragma unroll
  for (int j = 0; j < SW_XB; ++j) tn[j] = sw_tab<MULTI>(tab, min(j, WL - 1) * LPB + tl, gadd);
  for (int o0 = 0; o0 < WL; o0 += SW_XB) {
    u32 t[SW_XB];
    int e0[SW_XB], e1[SW_XB], s0v[SW_XB], s1v[SW_XB];
#pragma unroll
    for (int j = 0; j < SW_XB; ++j) { t[j] = tn[j]; tn[j] = sw_tab<MULTI>(tab, min(o0 + SW_XB + j, WL - 1) * LPB + tl, gadd); }
#pragma unroll
    for (int j = 0; j < SW_XB; ++j) {
      const u32 k0 = t[j] & 0xffffu, k1 = t[j] >> 16;
      e0[j] = *reinterpret_cast<const int16_t*>(Bb + k0); e1[j] = *reinterpret_cast<const int16_t*>(Bb + k1);
      s0v[j] = S0B[k0 >> 1]; s1v[j] = S0B[k1 >> 1];
    }
#pragma unroll
    for (int j = 0; j < SW_XB; ++j)
      if (o0 + j < WL) Aw[sw_idx(o0 + j, lane)] = pack2(s0v[j] + (e0[j] >> 1), s1v[j] + (e1[j] >> 1));
  }
}

// back to natural order: A = s0 + extrinsic (TD16:1241-1265); hard decisions one byte per natural position (tabk: pi(j))
template <bool MULTI>
__device__ __forceinline__ void sw_x2(unsigned char* __restrict__ Ab, const u32* __restrict__ Bw, const signed char* __restrict__ S0B,
                                      unsigned char* __restrict__ hdb, const u32* __restrict__ tab, const u32* __restrict__ tabk,
                                      u32 gadd, bool hd, int WL, int LPB, int lane, int tl) {
  u32 tn[SW_XB], kn[SW_XB];
#pragma unroll
  for (int j = 0; j < SW_XB; ++j) {
    const int i = min(j, WL - 1) * LPB + tl;
    tn[j] = sw_tab<MULTI>(tab, i, gadd);
    kn[j] = __ldg(tabk + i);
  }
  for (int o0 = 0; o0 < WL; o0 += SW_XB) {
    u32 t[SW_XB], v[SW_XB], kk[SW_XB];
    int s0v[SW_XB], s1v[SW_XB];
#pragma unroll
    for (int j = 0; j < SW_XB; ++j) {
      const int i = min(o0 + SW_XB + j, WL - 1) * LPB + tl;
      t[j] = tn[j]; kk[j] = kn[j];
      tn[j] = sw_tab<MULTI>(tab, i, gadd);
      kn[j] = __ldg(tabk + i);
      v[j] = Bw[sw_idx(min(o0 + j, WL - 1), lane)];
    }
#pragma unroll
    for (int j = 0; j < SW_XB; ++j) { s0v[j] = S0B[(t[j] & 0xffffu) >> 1]; s1v[j] = S0B[t[j] >> 17]; }
#pragma unroll
    for (int j = 0; j < SW_XB; ++j) {
      if (o0 + j < WL) {
        *reinterpret_cast<int16_t*>(Ab + (t[j] & 0xffffu)) = (int16_t)(s0v[j] + (lo16(v[j]) >> 1));
        *reinterpret_cast<int16_t*>(Ab + (t[j] >> 16)) = (int16_t)(s1v[j] + (hi16(v[j]) >> 1));
        if (hd) {
          hdb[kk[j] & 0xffffu] = (unsigned char)(v[j] & 1u);
          hdb[kk[j] >> 16] = (unsigned char)((v[j] >> 16) & 1u);
        }
      }
    }
  }
}

// Grid: one 32-thread CTA per block index.  The blocks of a batch are sorted by K; CTA i leads the group of up to
// G = 64 / NW blocks [i, ...) of equal K that ends at the next index that is a multiple of G, and returns at once when
// block i belongs to an earlier group.
__global__ void __launch_bounds__(32, 4) k_turbo_sw(SwArgs p) {
  extern __shared__ __align__(16) unsigned char sw_smem[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x;
  const int i0 = blockIdx.x;
  const int K = p.meta[i0].K;
  const int NW = sw_windows(K), LPB = NW >> 1, G = 32 / LPB, WL = K / NW;
  if ((i0 % G) != 0 && p.meta[i0 - 1].K == K) return;
  int cnt = 1;
  while (cnt < G && i0 + cnt < p.nblk && ((i0 + cnt) % G) != 0 && p.meta[i0 + cnt].K == K) ++cnt;
  const int g = lane / LPB, tl = lane - g * LPB;
  const bool valid = g < cnt;
  const int blk = i0 + (valid ? g : 0);              // spare lanes shadow block i0 (no outputs)
  const CbMeta m = p.meta[blk];
  CbState* st = &p.state[blk];

  const int nseg = (WL + 15) >> 4;
  signed char* S0B = reinterpret_cast<signed char*>(sw_smem);          // [WL][32] int8 pairs, like the halfwords of A / B
  uint16_t* P1 = reinterpret_cast<uint16_t*>(sw_smem) + WL * 32;
  u32* Aw = reinterpret_cast<u32*>(sw_smem) + WL * 32;
  u32* Bw = Aw + WL * 32;
  u32* CK = Bw + WL * 32;
  u32* NII = CK + nseg * 256;
  u32* BITS = NII + 1024;
  int16_t* TERM = reinterpret_cast<int16_t*>(BITS + SW_BITS_WORDS);
  uint16_t* P2 = reinterpret_cast<uint16_t*>(TERM + 128);
  unsigned char* HD = reinterpret_cast<unsigned char*>(CK);      // hard decisions, one byte per position (not live with CK)

  // ---- soft bits -> shared memory: scale to 8 bits, demultiplex (TD16:1055-1189) into the window layout -----------------
  const int16_t* y = p.in_base + ((((size_t)m.in_off_hi) << 32) | m.in_off_lo);
  const int ny = 3 * K + 12;
  const bool vec = (reinterpret_cast<uintptr_t>(y) & 7) == 0;      // (3K+12 int16 per block: 8-byte alignment is the normal case)
  u32 sum = 0;
  if (vec) {
    const uint2* y2 = reinterpret_cast<const uint2*>(y);
    const int n4 = ny >> 2;
    for (int i = tl; i < n4; i += 12 * LPB) {          // 12 loads in flight per thread (4 warps per SM: bandwidth needs the depth)
      uint2 v[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) v[j] = (i + j * LPB < n4) ? __ldg(y2 + i + j * LPB) : make_uint2(0u, 0u);
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const u32 a0 = __vabs2(v[j].x), a1 = __vabs2(v[j].y);      // |-32768| = 32768
        sum += (a0 & 0xffffu) + (a0 >> 16) + (a1 & 0xffffu) + (a1 >> 16);
      }
    }
  } else {
    for (int i = tl; i < ny; i += LPB) sum += (u32)abs((int)__ldg(y + i));
  }
  for (int o = LPB >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
  const u32 mean = sum / (u32)ny;
  int sh = 0;
  while ((mean >> sh) > 24u) ++sh;
  {
    unsigned char* p1b = reinterpret_cast<unsigned char*>(P1);
    unsigned char* p2b = reinterpret_cast<unsigned char*>(P2);
    int16_t* ah = reinterpret_cast<int16_t*>(Aw);
    const u32 magic = 0xffffffffu / (u32)WL + 1u;                 // k / WL for k < 2^16
    auto put = [&](int k, int s, int pa, int pb) {
      const int w = (int)__umulhi((u32)k, magic), o = k - w * WL;
      const int e = sw_idx(o, g * LPB + (w >> 1)), h = w & 1;
      S0B[e * 2 + h] = (signed char)sw_scale(s, sh);
      p1b[e * 2 + h] = (unsigned char)sw_scale(pa, sh);
      p2b[e * 2 + h] = (unsigned char)sw_scale(pb, sh);
      ah[e * 2 + h] = (int16_t)sw_scale(s, sh);
    };
    if (vec) {
      const uint2* y2 = reinterpret_cast<const uint2*>(y);
      const int nq = K >> 2;                                      // 4 positions = 12 soft bits = 3 x 8 bytes
      for (int q0 = tl; q0 < nq; q0 += 4 * LPB) {
        uint2 v[4][3];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int qd = q0 + j * LPB;
#pragma unroll
          for (int c = 0; c < 3; ++c) v[j][c] = (qd < nq) ? __ldg(y2 + 3 * qd + c) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int qd = q0 + j * LPB, k = qd << 2;
          if (qd < nq) {
            put(k, lo16(v[j][0].x), hi16(v[j][0].x), lo16(v[j][0].y));
            put(k + 1, hi16(v[j][0].y), lo16(v[j][1].x), hi16(v[j][1].x));
            put(k + 2, lo16(v[j][1].y), hi16(v[j][1].y), lo16(v[j][2].x));
            put(k + 3, hi16(v[j][2].x), lo16(v[j][2].y), hi16(v[j][2].y));
          }
        }
      }
    } else {
      for (int k = tl; k < K; k += LPB) put(k, __ldg(y + 3 * k), __ldg(y + 3 * k + 1), __ldg(y + 3 * k + 2));
    }
  }
  if (tl < 2) {                                      // tail metrics of decoder tl (same paths as TD16:474-520), state 0 = 0
    const int16_t* t = y + 3 * K + 6 * tl;
    int ts[3], tp[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { ts[i] = sw_scale(__ldg(t + 2 * i), sh); tp[i] = sw_scale(__ldg(t + 2 * i + 1), sh); }
    int c11 = (ts[2] + tp[2]) >> 1, c10;
    const int b0 = -c11, b1 = c11;
    c11 = (ts[1] + tp[1]) >> 1; c10 = (ts[1] - tp[1]) >> 1;
    const int b0_2 = b0 - c11, b1_2 = b0 + c11, b2_2 = b1 + c10, b3_2 = b1 - c10;
    c11 = (ts[0] + tp[0]) >> 1; c10 = (ts[0] - tp[0]) >> 1;
    const int tv[8] = {b0_2 - c11, b0_2 + c11, b1_2 + c10, b1_2 - c10, b2_2 - c10, b2_2 + c10, b3_2 + c11, b3_2 - c11};
#pragma unroll
    for (int s = 0; s < 8; ++s) TERM[(tl * 8 + s) * 8 + g] = (int16_t)(tv[s] - tv[0]);
  }
  __syncwarp();
#pragma unroll
  for (int d = 0; d < 2; ++d)
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      NII[((d * 2 + 0) * 8 + s) * 32 + lane] = (tl == 0) ? pack2(s ? -SW_Q : 0, 0) : 0u;
      NII[((d * 2 + 1) * 8 + s) * 32 + lane] = (tl == LPB - 1) ? pack2(0, TERM[(d * 8 + s) * 8 + g]) : 0u;
    }

  // table entry (per halfword): 2 x (step' << 6 | ((lane' + step') & 31) << 1 | half), the byte offset of the int16
  // element of pi(j) in A / B, for block 0 of the warp; then, in the same order, the natural positions pi(j)
  const u32* tab = p.tab_pool + p.tab_off[K >> 3];
  const u32* tabk = tab + WL * LPB;
  const u32 gadd = (u32)(g * LPB * 4) * 0x00010001u;
  const int nwb = (K + 31) >> 5;
  const bool run = valid && (m.flags & 1);
  if (valid && !(m.flags & 1) && tl == 0) st->status = 0xFE;          // not to be decoded (like k_demux16)
  bool done = !run;
  if (run && m.max_iter == 0) {                      // no iteration at all: the reference returns 1 (TD16:985,1201)
    if (tl == 0) { st->status = 1; if (p.status_out) p.status_out[blk] = 1; }
    done = true;
  }
  const int my_max = run ? (int)m.max_iter : 0;
  const int mx = __reduce_max_sync(FULL, my_max);
  u32* bits = BITS + g * nwb;
  unsigned char* hdb = HD + g * K;

  for (int it = 1; it <= mx; ++it) {
    sw_pass(Aw, Bw, P1, CK, NII, TERM, WL, lane, tl, LPB, g);
    __syncwarp();
    if (G == 1) sw_x1<false>(Aw, reinterpret_cast<const unsigned char*>(Bw), S0B, tab, 0u, WL, LPB, lane, tl);
    else sw_x1<true>(Aw, reinterpret_cast<const unsigned char*>(Bw), S0B, tab, gadd, WL, LPB, lane, tl);
    __syncwarp();
    sw_pass(Aw, Bw, P2, CK, NII + 512, TERM + 64, WL, lane, tl, LPB, g);
    __syncwarp();
    const bool hd = it > 1;                          // TD16:1267
    if (G == 1) sw_x2<false>(reinterpret_cast<unsigned char*>(Aw), Bw, S0B, hdb, tab, tabk, 0u, hd, WL, LPB, lane, tl);
    else sw_x2<true>(reinterpret_cast<unsigned char*>(Aw), Bw, S0B, hdb, tab, tabk, gadd, hd, WL, LPB, lane, tl);
    __syncwarp();
    if (hd) {                                        // CRC of the block (TD16:1305-1351; see block_crc_check), LPB lanes per block
      typedef unsigned long long u64;
      const int nb = K >> 3;
      for (int wi = tl; wi < nwb; wi += LPB) {       // 32 decisions -> one word, first position in bit 31
        u32 word = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int pos = (wi << 5) + (j << 2);
          u32 x4 = (pos < K) ? *reinterpret_cast<const u32*>(hdb + pos) : 0u;      // K is a multiple of 8
          word = (word << 4) | (((x4 & 0x01010101u) * 0x08040201u) >> 24);
        }
        bits[wi] = word;
      }
      __syncwarp();
      const int ct = m.crc_type;
      const int cw = (ct <= 1) ? 24 : (ct == 2 ? 16 : 8);
      const int j_lo = (ct == 0) ? ((m.F >> 3) << 3) : 0;
      const int j_hi = j_lo + K - cw - ((ct == 0) ? m.F : 0);
      const int Q = j_hi >> 5, r = j_hi & 31;
      u64 acc = 0;
      for (int wi = tl; wi < nwb; wi += LPB) {
        u32 mb = bits[wi];
        const int lead = j_lo - (wi << 5);
        if (lead > 0) mb = (lead < 32) ? (mb & (0xffffffffu >> lead)) : 0u;
        if (wi < Q) acc ^= clmul32(mb, __ldg(p.crc_xp + (ct * 32 + r) * CRC_NM + (Q - wi - 1)));
        else if (wi == Q && r > 0) acc ^= clmul32(mb >> (32 - r), __ldg(p.crc_xp + (ct * 32) * CRC_NM));
      }
      u32 alo = (u32)acc, ahi = (u32)(acc >> 32);
      for (int o = LPB >> 1; o > 0; o >>= 1) { alo ^= __shfl_xor_sync(FULL, alo, o); ahi ^= __shfl_xor_sync(FULL, ahi, o); }
      const u32 poly = (ct == 0) ? 0x864cfbu : (ct == 1) ? 0x800063u : (ct == 2) ? 0x1021u : 0x9Bu;
      const u64 pfull = ((u64)1 << cw) | poly;
      u64 V = ((u64)ahi << 32) | alo;
      for (int bit = 31 + cw; bit >= cw; --bit)
        if ((V >> bit) & 1) V ^= pfull << (bit - cw);
      const u32 crc = (u32)V & (u32)(((u64)1 << cw) - 1);
      auto byte_at = [&](int b) -> u32 { return (bits[b >> 2] >> (24 - 8 * (b & 3))) & 0xffu; };
      u32 rx;
      if (cw == 24) rx = (byte_at(nb - 3) << 16) | (byte_at(nb - 2) << 8) | byte_at(nb - 1);
      else if (cw == 16) rx = (byte_at(nb - 1) << 8) | byte_at(nb - 2);
      else rx = byte_at(nb - 1);
      const bool pass = (crc == rx && crc != 0);     // TD16:1348
      int s = 0;
      if (!done) {
        if (pass) s = it;
        else if (it >= (int)m.max_iter) s = m.max_iter + 1;
      }
      if (s) {
        uint8_t* outp = p.out_base + m.out_off;
        for (int wi = tl; wi < nwb; wi += LPB) {
          const u32 word = __byte_perm(bits[wi], 0, 0x0123);      // first byte in the low 8 bits
          const int b0 = wi << 2;
          if (b0 + 3 < nb && ((reinterpret_cast<uintptr_t>(outp) & 3) == 0)) reinterpret_cast<u32*>(outp)[wi] = word;
          else
            for (int qq = 0; qq < 4 && b0 + qq < nb; ++qq) outp[b0 + qq] = (uint8_t)(word >> (8 * qq));
        }
        if (tl == 0) { st->status = s; if (p.status_out) p.status_out[blk] = (uint8_t)s; }
        done = true;
      }
    } else if (!done && it >= (int)m.max_iter) {     // max_iterations == 1: no CRC test at all (TD16:1267), nothing decoded
      if (tl == 0) { st->status = m.max_iter + 1; if (p.status_out) p.status_out[blk] = (uint8_t)(m.max_iter + 1); }
      done = true;
    }
    if (__all_sync(FULL, done)) break;
    __syncwarp();
  }
}

}  // namespace oai
