// Receive-side rate dematching, NULL map and sub-block deinterleaving on the GPU.
//
//   k_dummy_w   : generate_dummy_w                (reference: lte_rate_matching.c:293-382)
//   k_rm_rx     : lte_rate_matching_turbo_rx      (reference: lte_rate_matching.c:688-831)
//   k_deint     : sub_block_deinterleaving_turbo  (reference: lte_rate_matching.c:193-243)
//
// One CTA per code block.  The reference's dematching loop is serial: it walks the circular
// buffer from k0, skips NULL slots and adds soft bit k to the k-th non-NULL slot it meets,
// wrapping around until E bits are consumed.  Parallel form: with N non-NULL slots in
// [0,Ncb) and rank(ind) = number of non-NULL slots met before ind on that walk, slot ind
// receives e[rank], e[rank+N], e[rank+2N], ... (< E).  int16 accumulation WRAPS in the
// reference (plain `+=`, :749,765), so the sum is order-independent mod 2^16 and exact.
#pragma once
#include "td_common.cuh"

namespace oai {

constexpr int RM_THREADS = 256;

struct RmBlock {
  uint32_t K, F;            // block size, filler bits (F only used when dummy == nullptr)
  uint32_t RTC, Kpi, ND;    // rows, 32*RTC, Kpi - (K+4)
  uint32_t Ncb, k0, E;      // circular buffer length, start, soft bits of this block
  uint32_t clear;
  uint32_t w_off;           // int16 offset of this block's w in its pool (3*Kpi entries)
  uint32_t w_sel;           // 0: staging pool of the batch (host-authoritative w), 1: device-resident HARQ pool
  uint32_t e_off_lo, e_off_hi;   // int16 offset of this block's soft bits in the input pool
  uint32_t dummy_off;       // byte offset of a caller-provided NULL map, or 0xffffffff: derive from (K,F)
  uint32_t y_off_lo, y_off_hi;   // int16 offset of the decoder input y (3K+12) written by k_deint
  uint32_t gold_off;        // word offset of this block's scrambling sequence in the Gold pool, 0xffffffff: soft bits are not scrambled
  uint32_t scr_off;         // position of this block's first soft bit in that sequence (r_offset, dlsch_decoding.c:333-347)
  uint32_t e_fmt;           // 0: soft bits are int16 (the reference's type), 1: int8 (narrow host feed; same values)
  uint32_t cnt_off;         // halfword offset of this (K,F)'s prefix-count table in the table pool, 0xffffffff: none (scan in the kernel)
};

// Pseudo-random sequences of 36.211 7.2, 32 bits per step like the reference's lte_gold_generic
// (LTE_REFSIG/lte_gold.c:151-180): x1 starts from 1 + 2^31, x2 from c_init completed with its 32nd bit, the
// first 1600 outputs are skipped.  One thread per sequence (a codeword needs at most ~2 900 words); bit j of word i
// is c(32 i + j).
struct GoldSeq { uint32_t c_init, off, nwords; };
__global__ void k_gold(const GoldSeq* seqs, int nseq, uint32_t* pool) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseq) return;
  const GoldSeq q = seqs[i];
  uint32_t x1 = 1u + (1u << 31), x2 = q.c_init;
  x2 = x2 ^ ((x2 ^ (x2 >> 1) ^ (x2 >> 2) ^ (x2 >> 3)) << 31);
  auto step = [&]() {
    x1 = (x1 >> 1) ^ (x1 >> 4);
    x1 = x1 ^ (x1 << 31) ^ (x1 << 28);
    x2 = (x2 >> 1) ^ (x2 >> 2) ^ (x2 >> 3) ^ (x2 >> 4);
    x2 = x2 ^ (x2 << 31) ^ (x2 << 30) ^ (x2 << 29) ^ (x2 << 28);
  };
  for (int n = 1; n < 50; ++n) step();
  for (uint32_t w = 0; w < q.nwords; ++w) { step(); pool[q.off + w] = x1 ^ x2; }
}

__host__ __device__ __forceinline__ uint32_t brev5(uint32_t c) {
#ifdef __CUDA_ARCH__
  return __brev(c) >> 27;
#else
  return ((c & 1) << 4) | ((c & 2) << 2) | (c & 4) | ((c & 8) >> 2) | ((c & 16) >> 4);
#endif
}
__host__ __device__ __forceinline__ uint32_t div_magic(uint32_t x, uint32_t magic) {
#ifdef __CUDA_ARCH__
  return __umulhi(x, magic);
#else
  return (uint32_t)(((unsigned long long)x * magic) >> 32);
#endif
}

// NULL predicate of generate_dummy_w for circular-buffer index `ind` (reference :329-370).
// Only rows 0..2 are ever marked for streams 0/1 and row 0 for stream 2 -- restated as is.
// magic = floor(2^32 / RTC) + 1: ind / RTC as one multiply-high (exact for ind < 2^16, RTC <= 193).
__host__ __device__ __forceinline__ bool dummy_is_null(uint32_t ind, uint32_t RTC, uint32_t Kpi, uint32_t ND, uint32_t F, uint32_t magic) {
  if (ind < Kpi) {                                   // stream 0: w[k], k = col*RTC + row
    const uint32_t col = div_magic(ind, magic), row = ind - col * RTC;
    return row <= 2 && brev5(col) + 32 * row < ND + F;
  }
  const uint32_t j = ind - Kpi, k = j >> 1;
  const uint32_t col = div_magic(k, magic), row = k - col * RTC;
  if ((j & 1) == 0) return row <= 2 && brev5(col) + 32 * row < ND + F;       // stream 1: w[Kpi+2k]
  if (ND > 0 && ind == 3 * Kpi - 1) return true;                               // :369-370
  return row == 0 && brev5(col) + 1 < ND;                                      // stream 2: w[Kpi+2k+1]
}

// generate_dummy_w on a device copy of the caller's buffer: only SETS LTE_NULL (=2)
__global__ void k_dummy_w(uint8_t* w, uint32_t RTC, uint32_t Kpi, uint32_t ND, uint32_t F) {
  const uint32_t magic = 0xffffffffu / RTC + 1;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * Kpi; i += gridDim.x * blockDim.x)
    if (dummy_is_null(i, RTC, Kpi, ND, F, magic)) w[i] = 2;
}

// Ranks on the circular-buffer walk.  The slots are handled in groups of 32 (one warp ballot each, RM_GROUPS_MAX groups for
// 3*Kpi = 18 528 slots); s_bal[g] holds the ballot of the slots that carry a bit, s_pre[g] the number of such slots in all
// groups before g.  One block-wide scan replaces the running prefix that cost two barriers per 256 slots.
constexpr int RM_GROUPS_MAX = 580;
__device__ __forceinline__ uint32_t rm_scan_groups(const uint32_t* s_bal, uint32_t* s_pre, uint32_t ng, uint32_t* s_w) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  static_assert(RM_GROUPS_MAX <= 3 * RM_THREADS, "three groups per thread");
  uint32_t loc[3], tsum = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const uint32_t g = threadIdx.x * 3 + j;
    loc[j] = (g < ng) ? __popc(s_bal[g]) : 0u;
    tsum += loc[j];
  }
  uint32_t x = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) s_w[wid] = x;
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int k = 0; k < RM_THREADS / 32; ++k) { const uint32_t v = s_w[k]; total += v; wbase += (k < wid) ? v : 0u; }
  uint32_t excl = wbase + x - tsum;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const uint32_t g = threadIdx.x * 3 + j;
    if (g < ng) { s_pre[g] = excl; excl += loc[j]; }
  }
  __syncthreads();
  return total;
}

// Pass 1 takes one ballot per 32 circular-buffer slots of the slots that carry soft bits; a block-wide scan gives every
// group its prefix count (and N, and the count before the start index); pass 2 walks the buffer coalesced: every thread
// knows the rank of its slot on the reference's walk and adds e[rank], e[rank+N], ... to it.
// With gold != nullptr and b.gold_off set, the soft bits are descrambled on the fly (dlsch_unscrambling,
// LTE_TRANSPORT/dlsch_scrambling.c:99-138): soft bit k is NEGATED where scrambling bit scr_off + k is 0.  (The reference
// stores the product back as int16, so -32768 stays -32768; under the int16 wrap of the accumulation below, adding
// +32768 instead is the same.)
// cnt_tab: pool of per-(K,F) prefix-count tables (cnt[i] = number of slots in [0,i) that carry a bit; 3*Kpi + 1 entries,
// built once on the host from the same NULL predicate): with one, the ranks are two coalesced table reads per slot and the
// kernel needs neither ballots nor a scan nor a barrier.
__global__ void __launch_bounds__(RM_THREADS) k_rm_rx(const RmBlock* blocks, int nblk, int16_t* w_pool,
                                                      const int16_t* e_pool, const uint8_t* dummy_pool,
                                                      int16_t* harq_pool = nullptr, const uint32_t* gold = nullptr,
                                                      const uint16_t* cnt_tab = nullptr) {
  __shared__ uint32_t s_bal[RM_GROUPS_MAX], s_pre[RM_GROUPS_MAX], s_w[RM_THREADS / 32];
  const int blk = blockIdx.x;
  if (blk >= nblk) return;
  const RmBlock b = blocks[blk];
  int16_t* w = (b.w_sel ? harq_pool : w_pool) + b.w_off;
  const int16_t* e = e_pool + (((long)b.e_off_hi << 32) | b.e_off_lo);
  const int8_t* e8 = reinterpret_cast<const int8_t*>(e);           // e_fmt == 1: the same soft bits, one byte each
  const bool narrow = b.e_fmt != 0;
  const uint8_t* dm = (b.dummy_off == 0xffffffffu) ? nullptr : dummy_pool + b.dummy_off;
  const uint32_t magic = 0xffffffffu / b.RTC + 1;
  const uint32_t* gs = (gold && b.gold_off != 0xffffffffu) ? gold + b.gold_off : nullptr;
  const int lane = threadIdx.x & 31;
  // the first reference loop runs only when k0 < Ncb (:747)
  const uint32_t start = (b.k0 < b.Ncb) ? b.k0 : 0;
  if (cnt_tab && b.cnt_off != 0xffffffffu && !dm) {
    const uint16_t* cnt = cnt_tab + b.cnt_off;
    const uint32_t N = cnt[b.Ncb], bs = cnt[start];
    if (N == 0) return;
    auto slot = [&](uint32_t c, bool data, int old) -> int {
      int acc = (b.clear == 1) ? 0 : old;                  // memset(w,0,Ncb) when clear==1 (:741-742)
      if (data) {
        const uint32_t rank = (c >= bs) ? (c - bs) : (c + N - bs);
        for (uint32_t k = rank; k < b.E; k += N) {
          int v = narrow ? (int)e8[k] : (int)e[k];
          if (gs) { const uint32_t pos = b.scr_off + k; if (!((gs[pos >> 5] >> (pos & 31)) & 1u)) v = -v; }
          acc += v;
        }
      }
      return acc;
    };
    // four slots per thread and step: one 8-byte table read (+ the next entry), one 8-byte read-modify-write of w.
    // (table offsets are multiples of 8 entries; w_off is a multiple of 4 for every legal Kpi / pool slot)
    const bool vec = ((b.w_off & 3u) == 0);
    const uint32_t n4 = vec ? (b.Ncb >> 2) : 0;
    uint2* w2 = reinterpret_cast<uint2*>(w);
    for (uint32_t q = threadIdx.x; q < n4; q += RM_THREADS) {
      const uint2 cc = reinterpret_cast<const uint2*>(cnt)[q];
      const uint32_t c0 = cc.x & 0xffffu, c1 = cc.x >> 16, c2 = cc.y & 0xffffu, c3 = cc.y >> 16, c4 = cnt[4 * q + 4];
      uint2 old = make_uint2(0u, 0u);
      if (b.clear != 1) old = w2[q];
      const int a0 = slot(c0, c1 != c0, (int)(int16_t)(old.x & 0xffffu)), a1 = slot(c1, c2 != c1, (int)(int16_t)(old.x >> 16));
      const int a2 = slot(c2, c3 != c2, (int)(int16_t)(old.y & 0xffffu)), a3 = slot(c3, c4 != c3, (int)(int16_t)(old.y >> 16));
      w2[q] = make_uint2(((u32)a0 & 0xffffu) | ((u32)a1 << 16), ((u32)a2 & 0xffffu) | ((u32)a3 << 16));
    }
    for (uint32_t i = 4 * n4 + threadIdx.x; i < b.Ncb; i += RM_THREADS) {      // tail (Ncb mod 4), or everything when unaligned
      const uint32_t c = cnt[i];
      w[i] = (int16_t)slot(c, cnt[i + 1] != c, (int)w[i]);
    }
    return;
  }
  const uint32_t ng = (b.Ncb + 31) >> 5;
  for (uint32_t c0 = 0; c0 < b.Ncb; c0 += RM_THREADS) {
    const uint32_t i = c0 + threadIdx.x;
    const bool f = (i < b.Ncb) && !(dm ? (dm[i] == 2) : dummy_is_null(i, b.RTC, b.Kpi, b.ND, b.F, magic));
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0 && (i >> 5) < ng) s_bal[i >> 5] = bal;
  }
  __syncthreads();
  const uint32_t N = rm_scan_groups(s_bal, s_pre, ng, s_w);         // non-NULL slots in [0,Ncb)
  if (N == 0) return;                                   // (the reference would loop forever)
  const uint32_t before_start = s_pre[start >> 5] + __popc(s_bal[start >> 5] & ((1u << (start & 31)) - 1u));
  for (uint32_t i = threadIdx.x; i < b.Ncb; i += RM_THREADS) {
    const uint32_t bal = s_bal[i >> 5];
    int acc = (b.clear == 1) ? 0 : (int)w[i];            // memset(w,0,Ncb) when clear==1 (:741-742)
    if ((bal >> lane) & 1u) {
      const uint32_t c = s_pre[i >> 5] + __popc(bal & ((1u << lane) - 1u));
      const uint32_t rank = (c >= before_start) ? (c - before_start) : (c + N - before_start);
      if (gs) {
        for (uint32_t k = rank; k < b.E; k += N) {
          const uint32_t pos = b.scr_off + k;
          const int v = narrow ? (int)e8[k] : (int)e[k];
          acc += ((gs[pos >> 5] >> (pos & 31)) & 1u) ? v : -v;
        }
      } else {
        if (narrow) { for (uint32_t k = rank; k < b.E; k += N) acc += e8[k]; }
        else { for (uint32_t k = rank; k < b.E; k += N) acc += e[k]; }
      }
    }
    w[i] = (int16_t)acc;                                 // wraps like the reference's int16 +=
  }
}

// w (three sub-blocks) -> d triples.  Output element q of d1 = d - 3*ND (reference :216-231):
//   q = 3*idx   <- w[k], q = 3*idx+1 <- w[Kpi+2k], q = 3*idx+5 <- w[Kpi+2k+1],  k = brev5(idx&31)*RTC + (idx>>5)
// Elements q = 2, 3*Kpi, 3*Kpi+1 are never written by the reference.  out[q - q_lo] for q in [q_lo,q_hi).
__global__ void __launch_bounds__(RM_THREADS) k_deint(const RmBlock* blocks, int nblk, const int16_t* w_pool,
                                                      int16_t* y_pool, int api_mode, const int16_t* harq_pool = nullptr) {
  extern __shared__ int16_t sw[];
  const int blk = blockIdx.x;
  if (blk >= nblk) return;
  const RmBlock b = blocks[blk];
  const int16_t* w = (b.w_sel ? harq_pool : w_pool) + b.w_off;
  for (uint32_t i = threadIdx.x; i < 3 * b.Kpi / 2; i += RM_THREADS)
    reinterpret_cast<uint32_t*>(sw)[i] = reinterpret_cast<const uint32_t*>(w)[i];
  __syncthreads();
  int16_t* y = y_pool + (((long)b.y_off_hi << 32) | b.y_off_lo);
  // api_mode: the whole range the reference touches, q in [0, 3*Kpi+3) (y points at d1[0]);
  // batch mode: the decoder input only, q in [3*ND, 3*ND + 3K+12) (y points at d[0])
  const uint32_t q_lo = api_mode ? 0 : 3 * b.ND;
  const uint32_t q_hi = api_mode ? 3 * b.Kpi + 3 : 3 * b.ND + 3 * b.K + 12;
  for (uint32_t q = q_lo + threadIdx.x; q < q_hi; q += RM_THREADS) {
    const uint32_t r = q % 3;
    if (q == 2 || q == 3 * b.Kpi || q == 3 * b.Kpi + 1) continue;
    const uint32_t idx = (r == 2) ? (q - 5) / 3 : q / 3;
    const uint32_t k = brev5(idx & 31) * b.RTC + (idx >> 5);
    const int16_t v = (r == 0) ? sw[k] : (r == 1 ? sw[b.Kpi + 2 * k] : sw[b.Kpi + 2 * k + 1]);
    y[q - q_lo] = v;
  }
}

}  // namespace oai
