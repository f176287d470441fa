// Uplink front of the rate dematcher on the GPU: from the demodulator's soft bits of one PUSCH allocation
// (lte_eNB_pusch_vars->llr, column by column of the channel interleaver matrix) to
//   * the data soft bits e[] in HBM, where k_rm_rx picks them up (no host hop),
//   * the HARQ-ACK and RI soft sums and decisions, and the CQI soft bits q[] for the host's convolutional decoder.
// Reference: openair1/PHY/LTE_TRANSPORT/ulsch_decoding.c:600-733 (Gold sequence, placeholder handling, de-interleaver),
// :775-873 (q_ACK, q_RI), :877-1002 (CQI soft bits, e), :1052-1153 (decisions).  The quirks that define the result are
// listed in DESIGN.md (uplink front end); in short: placeholder signs are patched in the SEQUENCE (y-placeholder = sign of
// the symbol's first bit, x-placeholders = -1), products are stored as int16, ACK positions are zeroed after they were
// summed and stay in e, and the CQI / data walk only ever skips the run of RI symbols at the very start of the matrix.
//
// One CTA per allocation; thread s handles input symbol s = col*R' + r (coalesced reads of Qm soft bits) and writes it to
// its row-major position r*Cmux + col.  Sums go through shared-memory atomics on 32-bit ints and wrap to int16 at the end
// (the reference's int16 `+=` is order-independent mod 2^16).
#pragma once
#include "td_common.cuh"

namespace oai {

constexpr int ULF_THREADS = 256;

struct UlFrontDev {
  uint32_t llr_off_lo, llr_off_hi;   // byte offset of the allocation's soft bits in the llr pool
  uint32_t llr_fmt;                  // 0: int16, 1: int8
  uint32_t e_off_lo, e_off_hi;       // int16 offset of its e[] in the soft-bit pool of the batch
  uint32_t gold_off;                 // word offset of its scrambling sequence in the Gold pool
  uint32_t Qm, Cmux, Rp;             // bits per symbol, columns (Nsymb_pusch), rows (R'mux)
  uint32_t Qprime_RI, Qprime_ACK, Qprime_CQI, Hprime;
  uint32_t Ncp, O_ACK, O_RI, bundling, Nbundled;
  uint32_t cqi_off;                  // byte offset of its q[] in the CQI pool
};
struct UlFrontOut { int16_t q_ACK[18]; int16_t q_RI[6]; uint8_t o_ACK[2]; uint8_t o_RI; uint8_t pad; };

// i-th placeholder of a set of n on the reference's walk (row R'-1-(i>>2), column cs[(4 - (i&3)) & 3]), or -1
__device__ __forceinline__ int ulf_placeholder(uint32_t r, uint32_t col, uint32_t n, uint32_t cs_packed, uint32_t Rp) {
  int jj = -1;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (((cs_packed >> (8 * k)) & 0xffu) == col) jj = k;
  if (jj < 0) return -1;
  const uint32_t i = 4 * (Rp - 1 - r) + ((4 - jj) & 3);
  return (i < n) ? (int)i : -1;
}

__global__ void __launch_bounds__(ULF_THREADS) k_ul_front(const UlFrontDev* tbs, int ntb, const uint8_t* llr_pool, int16_t* e_pool,
                                                          const uint32_t* gold_pool, int8_t* cqi_pool, UlFrontOut* out) {
  __shared__ int s_ack[18], s_ri[6];
  const int tb = blockIdx.x;
  if (tb >= ntb) return;
  const UlFrontDev p = tbs[tb];
  const uint32_t Qm = p.Qm, Cm = p.Cmux, Rp = p.Rp, nsym = Rp * Cm;
  // 36.212 tables 5.2.2.8-1 / -2 (normal, extended cyclic prefix), one column per byte
  const uint32_t cs_ri = p.Ncp ? 0x08050300u : 0x0A070401u, cs_ack = p.Ncp ? 0x07060201u : 0x09080302u;
  const uint32_t len_ACK = (p.O_ACK == 1) ? Qm : (p.O_ACK == 2 ? 3 * Qm : 0), len_RI = (p.O_RI == 1) ? Qm : 0;
  if (threadIdx.x < 18) s_ack[threadIdx.x] = 0;
  if (threadIdx.x < 6) s_ri[threadIdx.x] = 0;
  __syncthreads();
  const uint8_t* lb = llr_pool + (((unsigned long long)p.llr_off_hi << 32) | p.llr_off_lo);
  const int16_t* l16 = reinterpret_cast<const int16_t*>(lb);
  const int8_t* l8 = reinterpret_cast<const int8_t*>(lb);
  const uint32_t* gold = gold_pool + p.gold_off;
  int16_t* e = e_pool + (((unsigned long long)p.e_off_hi << 32) | p.e_off_lo);
  int8_t* qc = cqi_pool + p.cqi_off;
  // leading run of RI symbols in row-major order: all the CQI / data walk ever skips (at most one per row-0 RI column)
  uint32_t L = 0;
  while (L < nsym && ulf_placeholder(L / Cm, L % Cm, p.Qprime_RI, cs_ri, Rp) >= 0) ++L;
  const uint32_t d0 = L + p.Qprime_CQI, d1 = L + p.Hprime;            // data symbols [d0, d1) in row-major order
  for (uint32_t s = threadIdx.x; s < nsym; s += ULF_THREADS) {
    const uint32_t col = s / Rp, r = s - col * Rp, in0 = s * Qm, sym = r * Cm + col;
    const int i_ri = ulf_placeholder(r, col, p.Qprime_RI, cs_ri, Rp), i_ack = ulf_placeholder(r, col, p.Qprime_ACK, cs_ack, Rp);
    int y[6];
    int sign0 = 1;
#pragma unroll
    for (uint32_t q = 0; q < 6; ++q) {
      if (q < Qm) {
        const uint32_t bit = in0 + q;
        int sg = 2 * (int)((gold[bit >> 5] >> (bit & 31)) & 1u) - 1;
        if (q == 0) sign0 = sg;
        if (i_ri >= 0) sg = (q == 1) ? sign0 : (q >= 2 ? -1 : sg);                           // :641-660
        if (i_ack >= 0) {                                                                    // :662-697
          if (p.O_ACK == 1) sg = (q == 1) ? (p.bundling == 0 ? sign0 : sg) : (q >= 2 ? -1 : sg);
          else if (p.O_ACK == 2) sg = (q >= 2) ? -1 : sg;
        }
        const int v = p.llr_fmt ? (int)l8[bit] : (int)l16[bit];
        y[q] = (int)(int16_t)(sg * v);                                                       // stored as int16 (:703-758)
      }
    }
    if (i_ack >= 0 && len_ACK) {
#pragma unroll
      for (uint32_t q = 0; q < 6; ++q)
        if (q < Qm) { atomicAdd(&s_ack[(q + Qm * (uint32_t)i_ack) % len_ACK], y[q]); y[q] = 0; }   // :822-840
    }
    if (i_ri >= 0 && len_RI) {
#pragma unroll
      for (uint32_t q = 0; q < 6; ++q)
        if (q < Qm) atomicAdd(&s_ri[(q + Qm * (uint32_t)i_ri) % len_RI], y[q]);                   // :866-873
    }
    if (sym >= L && sym < d0) {
#pragma unroll
      for (uint32_t q = 0; q < 6; ++q)
        if (q < Qm) qc[(sym - L) * Qm + q] = (int8_t)max(-128, min(127, y[q]));                   // :903-915
    } else if (sym >= d0 && sym < d1) {
#pragma unroll
      for (uint32_t q = 0; q < 6; ++q)
        if (q < Qm) e[(sym - d0) * Qm + q] = (int16_t)y[q];                                       // :925-1002
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {                                                                         // decisions, :1052-1153
    UlFrontOut o;
    for (int i = 0; i < 18; ++i) o.q_ACK[i] = (int16_t)s_ack[i];
    for (int i = 0; i < 6; ++i) o.q_RI[i] = (int16_t)s_ri[i];
    o.o_ACK[0] = o.o_ACK[1] = 0; o.o_RI = 0; o.pad = 0;
    const int widx = (p.bundling == 0) ? 4 : (int)((p.Nbundled - 1) & 3);
    // wACK_RX rows (36.213 table 7.3-1 as +-1): only entries 0 and 1 are used
    const int w0 = (widx == 4) ? 1 : -1, w1 = (widx == 4) ? 1 : ((widx == 1 || widx == 3) ? 1 : -1);
    if (p.O_ACK == 1) {
      o.q_ACK[0] = (int16_t)(o.q_ACK[0] * w0);
      o.q_ACK[0] = (int16_t)(o.q_ACK[0] + ((p.bundling == 0) ? o.q_ACK[1] * w0 : o.q_ACK[1] * w1));
      o.o_ACK[0] = (o.q_ACK[0] < 0) ? 0 : 1;
    } else if (p.O_ACK == 2) {
      const int a = (Qm == 2) ? 3 : (Qm == 4 ? 5 : 7), b = (Qm == 2) ? 4 : (Qm == 4 ? 8 : 12), c2 = (Qm == 2) ? 2 : (Qm == 4 ? 4 : 6),
                d = (Qm == 2) ? 5 : (Qm == 4 ? 9 : 13);
      const int16_t n0 = (int16_t)(o.q_ACK[0] * w0 + o.q_ACK[a] * w1), n1 = (int16_t)(o.q_ACK[1] * w0 + o.q_ACK[b] * w1),
                    n2 = (int16_t)(o.q_ACK[c2] * w0 + o.q_ACK[d] * w1);
      o.q_ACK[0] = n0; o.q_ACK[1] = n1; o.q_ACK[2] = n2;
      o.o_ACK[0] = 1; o.o_ACK[1] = 1;
      int m = n0 + n1 - n2, mn = -n0 + n1 + n2;
      if (mn > m) { o.o_ACK[0] = 0; o.o_ACK[1] = 1; m = mn; }
      mn = n0 - n1 + n2;
      if (mn > m) { o.o_ACK[0] = 1; o.o_ACK[1] = 0; m = mn; }
      mn = -n0 - n1 - n2;
      if (mn > m) { o.o_ACK[0] = 0; o.o_ACK[1] = 0; m = mn; }
    }
    if (p.O_RI == 1 && p.Qprime_RI > 0) o.o_RI = ((o.q_RI[0] + o.q_RI[Qm / 2]) > 0) ? 0 : 1;
    out[tb] = o;
  }
}

}  // namespace oai
