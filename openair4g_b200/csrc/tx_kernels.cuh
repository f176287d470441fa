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
  uint32_t qpp_off;              // offset of pi[] of this K in the plain QPP pool
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
                                                               uint8_t* d_pool, const uint16_t* qpp_pool) {
  __shared__ uint8_t s_c[ENC_WARPS][768];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int blk = blockIdx.x * ENC_WARPS + wid;
  if (blk >= nblk) return;
  const TxBlock b = blocks[blk];
  const uint8_t* c = c_pool + off64(b.c_off_lo, b.c_off_hi);
  uint8_t* d = d_pool + off64(b.d_off_lo, b.d_off_hi);
  const uint16_t* pi = qpp_pool + b.qpp_off;
  const uint32_t KB = b.K >> 3;
  uint8_t* sc = s_c[wid];
  for (uint32_t i = lane; i < KB; i += 32) sc[i] = c[i];
  __syncwarp();
  auto bit = [&](uint32_t k) -> uint32_t { return (sc[k >> 3] >> (7u - (k & 7u))) & 1u; };     // MSB first
  const uint32_t nb = (KB + 31) / 32;
  const uint32_t k_lo = min(KB, lane * nb) * 8, k_hi = min(KB, (lane + 1) * nb) * 8;
  // pass 1: zero-state response of this lane's run, both constituent encoders
  uint32_t sa = 0, sb = 0;
  for (uint32_t k = k_lo; k < k_hi; ++k) { rsc_step(sa, bit(k)); rsc_step(sb, bit(pi[k])); }
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
  for (uint32_t k = k_lo; k < k_hi; k += 4) {
    uint32_t by[12];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t u = bit(k + q);
      by[3 * q] = u;
      by[3 * q + 1] = rsc_step(sa, u);
      by[3 * q + 2] = rsc_step(sb, bit(pi[k + q]));
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(d + 3 * (size_t)k);
#pragma unroll
    for (int w = 0; w < 3; ++w) o[w] = by[4 * w] | (by[4 * w + 1] << 8) | (by[4 * w + 2] << 16) | (by[4 * w + 3] << 24);
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
// E bits are out.  Same parallel form as k_rm_rx: with N non-NULL slots, the slot of rank c on the walk supplies
// e[c], e[c+N], ...  One CTA per block; w lives in shared memory.
__global__ void __launch_bounds__(RM_THREADS) k_rm_tx(const TxBlock* blocks, int nblk, const uint8_t* d_pool,
                                                      const uint8_t* w_pool, uint8_t* e_pool) {
  extern __shared__ uint8_t swb[];
  __shared__ uint32_t s_w[RM_THREADS / 32], s_a[RM_THREADS / 32], s_b[RM_THREADS / 32];
  const int blk = blockIdx.x;
  if (blk >= nblk) return;
  const TxBlock b = blocks[blk];
  if (b.E == 0 || b.Ncb < 3 * b.Kpi) return;            // "Exiting, RM condition" (:508-511): nothing written
  uint8_t* e = e_pool + off64(b.e_off_lo, b.e_off_hi);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t magic = 0xffffffffu / b.RTC + 1;
  const uint32_t start = (b.k0 < b.Ncb) ? b.k0 : 0;
  uint32_t cnt = 0, cntb = 0;
  if (b.w_from_d) {
    const uint8_t* d = d_pool + off64(b.d_off_lo, b.d_off_hi);
    const uint32_t D = b.K + 4;
    for (uint32_t i = threadIdx.x; i < b.Ncb; i += RM_THREADS) {
      uint8_t v = 2;
      if (!dummy_is_null(i, b.RTC, b.Kpi, b.ND, b.F, magic)) {
        const uint32_t k = (i < b.Kpi) ? i : ((i - b.Kpi) >> 1);
        const uint32_t col = __umulhi(k, magic), row = k - col * b.RTC;
        const uint32_t idx = brev5(col) + 32 * row;
        if (i < b.Kpi) v = d[3 * (idx - b.ND)];
        else if (((i - b.Kpi) & 1) == 0) v = d[3 * (idx - b.ND) + 1];
        else { uint32_t pos = idx + 1 - b.ND; if (pos == D) pos = 0; v = d[3 * pos + 2]; }     // stream 2 is shifted by one (:76,84)
      }
      swb[i] = v;
      cnt += (v != 2) ? 1u : 0u;
      cntb += (v != 2 && i < start) ? 1u : 0u;
    }
  } else {
    const uint8_t* w = w_pool + off64(b.w_off_lo, b.w_off_hi);
    for (uint32_t i = threadIdx.x; i < b.Ncb; i += RM_THREADS) {
      const uint8_t v = w[i];
      swb[i] = v;
      cnt += (v != 2) ? 1u : 0u;
      cntb += (v != 2 && i < start) ? 1u : 0u;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_xor_sync(0xffffffffu, cnt, o); cntb += __shfl_xor_sync(0xffffffffu, cntb, o); }
  if (lane == 0) { s_a[wid] = cnt; s_b[wid] = cntb; }
  __syncthreads();
  uint32_t N = 0, before_start = 0;
#pragma unroll
  for (int i = 0; i < RM_THREADS / 32; ++i) { N += s_a[i]; before_start += s_b[i]; }
  if (N == 0) return;                                   // (the reference would loop forever)
  uint32_t base = 0;
  for (uint32_t c0 = 0; c0 < b.Ncb; c0 += RM_THREADS) {
    const uint32_t i = c0 + threadIdx.x;
    const uint8_t v = (i < b.Ncb) ? swb[i] : (uint8_t)2;
    const bool f = v != 2;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_w[wid] = __popc(bal);
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < RM_THREADS / 32; ++k) { const uint32_t x = s_w[k]; tot += x; woff += (k < wid) ? x : 0u; }
    if (f) {
      const uint32_t c = base + woff + __popc(bal & ((1u << lane) - 1u));
      const uint32_t rank = (c >= before_start) ? (c - before_start) : (c + N - before_start);
      for (uint32_t k = rank; k < b.E; k += N) e[k] = v;
    }
    base += tot;
    __syncthreads();
  }
}

}  // namespace oai
