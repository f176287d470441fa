// Transmit-side mirror of the path on the GPU: vector generation for the harness and for
// on-device tests -- not on the decoding hot path.
//
//   k_turbo_enc : threegpplte_turbo_encoder      (reference: 3gpplte_sse.c:380-476; termination :  rsc tail, 36.212 5.1.3.2.2)
//   k_sbi_tx    : sub_block_interleaving_turbo   (reference: lte_rate_matching.c:51-130)
//   k_rm_tx     : lte_rate_matching_turbo        (reference: lte_rate_matching.c:464-634), optionally with the sub-block
//                 interleaver folded in (w is built in shared memory from the encoder output d and never stored)
//
// Encoder: the reference walks the two 8-state recursive encoders byte by byte through a 8 x 256 table.  The
// recursion is linear over GF(2): with the state s = (d1,d2,d3) and input u, a = u ^ d2 ^ d3, z = a ^ d1 ^ d3,
// s' = (a,d1,d2) = M s ^ (u,0,0).  One warp encodes one block: every lane takes a contiguous run of input bytes, runs
// it once from the zero state (its zero-state response), the lanes' responses are chained with M^len (M has order 7),
// and a second run from the true start state emits the parity bits.
#pragma once
#include "rm_kernels.cuh"

namespace oai {

constexpr int ENC_WARPS = 4;

struct TxBlock {
  uint32_t K, F;                 // block size; filler bits that are NULL in streams 0/1 of w (0: the reference's TX, which sends them)
  uint32_t RTC, Kpi, ND;
  uint32_t Ncb, k0, E;
  uint32_t f1, f2;               // QPP coefficients of this K: pi(i) = (f1 i + f2 i^2) mod K
  uint32_t w_from_d;             // 1: build w from d (fused interleaver), 0: w is given (w_off)
  uint32_t c_off_lo, c_off_hi;   // byte offset of the K/8 info bytes
  uint32_t d_off_lo, d_off_hi;   // byte offset of d (3K+12 bytes, multiple of 4)
  uint32_t e_off_lo, e_off_hi;   // byte offset of e (E bytes)
  uint32_t w_off_lo, w_off_hi;   // byte offset of a caller-provided w (3*Kpi bytes)
};
__device__ __forceinline__ long off64(uint32_t lo, uint32_t hi) { return (long)(((unsigned long long)hi << 32) | lo); }

// one trellis step of a constituent encoder; state bit 0 = d1 (newest) .. bit 2 = d3
__device__ __forceinline__ uint32_t rsc_step(uint32_t& s, uint32_t u) {
  const uint32_t a = (u ^ (s >> 1) ^ (s >> 2)) & 1u;          // feedback 1 + D^2 + D^3
  const uint32_t z = (a ^ s ^ (s >> 2)) & 1u;                 // parity   1 + D + D^3
  s = ((s << 1) & 6u) | a;
  return z;
}
__device__ __forceinline__ uint32_t lin3(uint32_t cols, uint32_t s) {       // (c1 | c2<<3 | c4<<6) applied to s
  return ((s & 1u) ? (cols & 7u) : 0u) ^ ((s & 2u) ? ((cols >> 3) & 7u) : 0u) ^ ((s & 4u) ? ((cols >> 6) & 7u) : 0u);
}

__global__ void __launch_bounds__(ENC_WARPS * 32) k_turbo_enc(const TxBlock* blocks, int nblk, const uint8_t* c_pool,
                                                               uint8_t* d_pool) {
  __shared__ uint8_t s_c[ENC_WARPS][768];
  __shared__ uint32_t s_i[ENC_WARPS][6][32];     // interleaved input bits of each lane's run: word w of lane l at [w][l]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int blk = blockIdx.x * ENC_WARPS + wid;
  if (blk >= nblk) return;
  const TxBlock b = blocks[blk];
  const uint8_t* c = c_pool + off64(b.c_off_lo, b.c_off_hi);
  uint8_t* d = d_pool + off64(b.d_off_lo, b.d_off_hi);
  const uint32_t K = b.K, KB = K >> 3;
  uint8_t* sc = s_c[wid];
  for (uint32_t i = lane; i < KB; i += 32) sc[i] = c[i];
  __syncwarp();
  const uint32_t nb = (KB + 31) / 32;            // <= 24 bytes = 6 words of bits per lane
  const uint32_t k_lo = min(KB, lane * nb) * 8, k_hi = min(KB, (lane + 1) * nb) * 8;
  // QPP positions by recursion instead of a table read per step (the first version was bound by the shared / global
  // load queues: ncu mio_throttle 65, lg_throttle 19 per issue): pi(k+1) = pi(k) + g(k), g(k+1) = g(k) + 2 f2 (mod K),
  // g(k) = f1 + f2 (2k + 1)
  uint32_t pi = (uint32_t)(((unsigned long long)b.f1 * k_lo + ((unsigned long long)b.f2 * k_lo % K) * k_lo) % K);
  uint32_t g = (uint32_t)((b.f1 + (unsigned long long)b.f2 * (2 * k_lo + 1)) % K);
  const uint32_t g2 = (2 * b.f2) % K;
  // pass 1: zero-state response of this lane's run, both constituent encoders; the interleaved input bits are kept
  uint32_t sa = 0, sb = 0, acc = 0;
  for (uint32_t k = k_lo; k < k_hi; k += 8) {
    const uint32_t byte = sc[k >> 3], sh = (k - k_lo) & 31u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      rsc_step(sa, (byte >> (7 - q)) & 1u);                                   // MSB first
      const uint32_t ui = (sc[pi >> 3] >> (7u - (pi & 7u))) & 1u;
      rsc_step(sb, ui);
      acc |= ui << (sh + q);
      pi += g; pi -= (pi >= K) ? K : 0u;
      g += g2; g -= (g >= K) ? K : 0u;
    }
    if (sh == 24 || k + 8 >= k_hi) { s_i[wid][(k - k_lo) >> 5][lane] = acc; acc = 0; }
  }
  // M^len: images of the three basis states under (k_hi - k_lo) zero-input steps (M^7 = 1)
  uint32_t c1 = 1, c2 = 2, c4 = 4;
  for (uint32_t i = 0; i < (k_hi - k_lo) % 7; ++i) { rsc_step(c1, 0); rsc_step(c2, 0); rsc_step(c4, 0); }
  const uint32_t pack = c1 | (c2 << 3) | (c4 << 6) | (sa << 9) | (sb << 12);
  uint32_t cura = 0, curb = 0, mya = 0, myb = 0;
  for (int l = 0; l < 32; ++l) {
    const uint32_t p = __shfl_sync(0xffffffffu, pack, l);
    if (lane == l) { mya = cura; myb = curb; }
    cura = lin3(p, cura) ^ ((p >> 9) & 7u);
    curb = lin3(p, curb) ^ ((p >> 12) & 7u);
  }
  // pass 2: d[3k] = c_k, d[3k+1] = z_k, d[3k+2] = z'_k; four steps = three aligned words
  sa = mya; sb = myb;
  for (uint32_t k = k_lo; k < k_hi; k += 8) {
    const uint32_t byte = sc[k >> 3], sh = (k - k_lo) & 31u;
    if (sh == 0) acc = s_i[wid][(k - k_lo) >> 5][lane];
    const uint32_t ib = acc >> sh;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t by[12];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t u = (byte >> (7 - 4 * h - q)) & 1u;
        by[3 * q] = u;
        by[3 * q + 1] = rsc_step(sa, u);
        by[3 * q + 2] = rsc_step(sb, (ib >> (4 * h + q)) & 1u);
      }
      uint32_t* o = reinterpret_cast<uint32_t*>(d + 3 * (size_t)(k + 4 * h));
#pragma unroll
      for (int w = 0; w < 3; ++w) o[w] = by[4 * w] | (by[4 * w + 1] << 8) | (by[4 * w + 2] << 16) | (by[4 * w + 3] << 24);
    }
  }
  // termination: x = d2 ^ d3 drives the register to zero, z = d1 ^ d3; x z x z x z of encoder 1, then of encoder 2
  if (lane == 0) {
    uint8_t* x = d + 3 * (size_t)b.K;
    uint32_t s = cura;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        x[6 * e + 2 * t] = (uint8_t)(((s >> 1) ^ (s >> 2)) & 1u);
        x[6 * e + 2 * t + 1] = (uint8_t)((s ^ (s >> 2)) & 1u);
        s = (s << 1) & 6u;
      }
      s = curb;
    }
  }
}

// sub_block_interleaving_turbo on a device copy of the range the reference reads: dbuf[0] = d[-3*ND] (the caller's
// bytes in front of d stand in for the <NULL> padding, :78-80), dbuf[3*ND + 3*D + 2] = d[2] already copied (:76).
__global__ void k_sbi_tx(const uint8_t* dbuf, uint8_t* w, uint32_t RTC, uint32_t Kpi, uint32_t ND) {
  const uint32_t magic = 0xffffffffu / RTC + 1;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < Kpi; k += gridDim.x * blockDim.x) {
    const uint32_t col = __umulhi(k, magic), row = k - col * RTC;
    const uint32_t i3 = 3 * (brev5(col) + 32 * row);
    w[k] = dbuf[i3];
    w[Kpi + 2 * k] = dbuf[i3 + 1];
    w[Kpi + 2 * k + 1] = (ND > 0 && k == Kpi - 1) ? (uint8_t)2 : dbuf[i3 + 5];      // :117-118
  }
}

// lte_rate_matching_turbo: e[k] = k-th non-NULL entry met on the walk of the circular buffer from k0, wrapping until
// E bits are out.  Same parallel form as k_rm_rx (one ballot per 32 slots, one block-wide scan): with N non-NULL slots,
// the slot of rank c on the walk supplies e[c], e[c+N], ...  One CTA per block; w lives in shared memory.
__global__ void __launch_bounds__(RM_THREADS) k_rm_tx(const TxBlock* blocks, int nblk, const uint8_t* d_pool,
                                                      const uint8_t* w_pool, uint8_t* e_pool) {
  extern __shared__ uint8_t swb[];
  __shared__ uint32_t s_bal[RM_GROUPS_MAX], s_pre[RM_GROUPS_MAX], s_w[RM_THREADS / 32];
  const int blk = blockIdx.x;
  if (blk >= nblk) return;
  const TxBlock b = blocks[blk];
  if (b.E == 0 || b.Ncb < 3 * b.Kpi) return;            // "Exiting, RM condition" (:508-511): nothing written
  uint8_t* e = e_pool + off64(b.e_off_lo, b.e_off_hi);
  const int lane = threadIdx.x & 31;
  const uint32_t magic = 0xffffffffu / b.RTC + 1;
  const uint32_t start = (b.k0 < b.Ncb) ? b.k0 : 0;
  const uint32_t ng = (b.Ncb + 31) >> 5;
  const uint8_t* d = d_pool + off64(b.d_off_lo, b.d_off_hi);
  const uint8_t* w = w_pool + off64(b.w_off_lo, b.w_off_hi);
  const uint32_t D = b.K + 4;
  for (uint32_t c0 = 0; c0 < b.Ncb; c0 += RM_THREADS) {
    const uint32_t i = c0 + threadIdx.x;
    uint8_t v = 2;
    if (i < b.Ncb) {
      if (!b.w_from_d) v = w[i];
      else if (!dummy_is_null(i, b.RTC, b.Kpi, b.ND, b.F, magic)) {
        const uint32_t k = (i < b.Kpi) ? i : ((i - b.Kpi) >> 1);
        const uint32_t col = __umulhi(k, magic), row = k - col * b.RTC;
        const uint32_t idx = brev5(col) + 32 * row;
        if (i < b.Kpi) v = d[3 * (idx - b.ND)];
        else if (((i - b.Kpi) & 1) == 0) v = d[3 * (idx - b.ND) + 1];
        else { uint32_t pos = idx + 1 - b.ND; if (pos == D) pos = 0; v = d[3 * pos + 2]; }     // stream 2 is shifted by one (:76,84)
      }
      swb[i] = v;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, v != 2);
    if (lane == 0 && (i >> 5) < ng) s_bal[i >> 5] = bal;
  }
  __syncthreads();
  const uint32_t N = rm_scan_groups(s_bal, s_pre, ng, s_w);
  if (N == 0) return;                                   // (the reference would loop forever)
  const uint32_t before_start = s_pre[start >> 5] + __popc(s_bal[start >> 5] & ((1u << (start & 31)) - 1u));
  for (uint32_t i = threadIdx.x; i < b.Ncb; i += RM_THREADS) {
    const uint32_t bal = s_bal[i >> 5];
    if ((bal >> lane) & 1u) {
      const uint8_t v = swb[i];
      const uint32_t c = s_pre[i >> 5] + __popc(bal & ((1u << lane) - 1u));
      const uint32_t rank = (c >= before_start) ? (c - before_start) : (c + N - before_start);
      for (uint32_t k = rank; k < b.E; k += N) e[k] = v;
    }
  }
}

}  // namespace oai
