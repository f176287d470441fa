// Shared device-side definitions for the B200 turbo-decoding kernels.
//
// Data layout in HBM ("C4 lane layout").  The reference decoder splits the K trellis
// positions of a code block into 8 SIMD lanes of W=K/8 steps (lane l covers positions
// [l*W,(l+1)*W), 3gpplte_turbo_decoder_sse_16bit.c:921-932).  We keep that split --
// it defines the bit-exact result -- but store every per-position int16 array so that
// the thread that owns lanes (2t,2t+1) of a block finds 4 consecutive steps of its
// two lanes in one 16-byte word group:
//     word(k,t)     = (k>>2)*16 + t*4 + (k&3)            (uint32 index, 2 lanes packed)
//     halfword(k,l) = 2*word(k,l>>1) + (l&1)
// so a 4-thread group reads 64 contiguous bytes per 4 steps with one LDG.128 each.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oai {

typedef uint32_t u32;

// ---- per-block metadata (device) ---------------------------------------------------
struct CbMeta {
  uint16_t K;          // block size (one of the 188 QPP sizes)
  uint16_t W;          // K/8 steps per lane
  uint8_t  max_iter;
  uint8_t  crc_type;
  uint8_t  F;
  uint8_t  flags;      // bit0: decode enabled
  uint32_t pi_off;     // offset of this K's H table (natural position -> C4 halfword index)
  uint32_t t_off;      // offset of this K's T table (layout order -> C4 index of the QPP image)
  uint32_t in_off_lo;  // input offset (int16 units) from the batch input base
  uint32_t in_off_hi;
  uint32_t out_off;    // output offset (bytes) from the batch output base
};

// per-block dynamic state (device)
struct CbState {
  int16_t  T[2][8];    // beta start metrics of lane 7 for MAP1 / MAP2 (tail bits)
  int32_t  max_in;     // max |y| over the 3K+12 inputs
  int32_t  max_sys;    // max |systematic input| of the next MAP pass
  int32_t  status;     // 0 = active, otherwise the decoder's return value
  int32_t  max_ext;    // max |ext| (first decoder's a-posteriori LLRs after feedback)
  int32_t  max_ext2;   // max |LLR| of the second decoder's last pass when that pass ran tracked and certified, else -1
  int32_t  cert[2];    // last range-certificate value (sp_a + sp_b + M) of a tracked pass of decoder 1 / 2: spreads grow from
                       // pass to pass, so a value close to the limit sends the next pass of that decoder to the exact policy
  int32_t  retry;      // bit 0: the last MAP pass ran tracked and failed its range certificate -> the retry launch repeats it
                       //        on the exact policy; bit 1: sticky -- later passes of this block go straight to the exact policy
};

// workspace arrays of one block slot, each `A` halfwords long
// ARR_B8A / ARR_B8B hold int8 copies (A bytes each, same C4 index in bytes): [P1 | P2] and [S0 | unused];
// the MAP kernel reads them instead of the int16 arrays when every |y| of the batch is <= 127.
enum { ARR_S0 = 0, ARR_P1 = 1, ARR_P2 = 2, ARR_SYS = 3, ARR_EXT = 4, ARR_EXT2 = 5, ARR_B8A = 6, ARR_B8B = 7, ARR_COUNT = 8 };

__host__ __device__ inline int c4_words(int W) { return ((W + 3) >> 2) << 4; }      // uint32 words per array
__host__ __device__ inline int c4_word(int k, int t) { return ((k >> 2) << 4) + (t << 2) + (k & 3); }
__host__ __device__ inline int c4_hw(int k, int lane) { return (c4_word(k, lane >> 1) << 1) + (lane & 1); }

// ---- packed int16x2 arithmetic policies ---------------------------------------------
// SatArith: the reference's arithmetic (_mm_adds_epi16/_mm_subs_epi16/_mm_max_epi16).
// WrapArith: plain two's-complement halfword ops (VIADD.16x2 / VIMNMX.S16x2 on sm_100a);
// identical results whenever no intermediate leaves the int16 range, which the
// per-pass guard (see DESIGN.md "fast-path guard") proves before this policy is chosen.
struct SatArith {
  static constexpr bool kInv = false;
  static __device__ __forceinline__ u32 enc(u32 v) { return v; }
  static __device__ __forceinline__ u32 dec(u32 r) { return r; }
  static __device__ __forceinline__ u32 add(u32 a, u32 b) { return __vaddss2(a, b); }
  static __device__ __forceinline__ u32 sub(u32 a, u32 b) { return __vsubss2(a, b); }
};
// InvArith: the exact policy of the MAP recursions in an order-reversing unsigned representation R(v) = 16383 - v.
// Normalised metrics are <= 0 and |branch metric| <= 16384, so every candidate a +- g lies in [-49152, 16383], i.e.
// R in [0, 65535] (the single exception, 0 + 16384, is detected per step and handled in signed arithmetic): the add
// never wraps, the reference's saturation at -32768 is a min with 49151, and max becomes min:
//     max(sat(ax + gx), sat(ay + gy))  ->  VIMNMX3.U16x2(ax' - gx, ay' - gy, 49151)      (2 VIADD + 1 ALU-pipe instruction)
//     sat(n - max_s n)                 ->  VIADDMNMX.U16x2(n', 16383 - min_s n', 49151)
// against 6 / 8 instructions for one emulated __vaddss2 / __vsubss2.  enc/dec convert from / to packed signed int16
// (the same involution); add/sub are the signed saturating forms used where values are signed (gamma, ext, feedback).
struct InvArith {
  static constexpr bool kInv = true;
  static __device__ __forceinline__ u32 add(u32 a, u32 b) { return __vaddss2(a, b); }
  static __device__ __forceinline__ u32 sub(u32 a, u32 b) { return __vsubss2(a, b); }
  static __device__ __forceinline__ u32 enc(u32 v) { return __vadd2(~v, 0x40004000u); }     // 16383 - v per halfword
  static __device__ __forceinline__ u32 dec(u32 r) { return __vadd2(~r, 0x40004000u); }
};
struct WrapArith {
  static __device__ __forceinline__ u32 add(u32 a, u32 b) { return __vadd2(a, b); }
  static __device__ __forceinline__ u32 sub(u32 a, u32 b) { return __vsub2(a, b); }
};
__device__ __forceinline__ u32 vmax(u32 a, u32 b) { return __vmaxs2(a, b); }
// per-halfword arithmetic shift right by one (_mm_srai_epi16(x,1))
__device__ __forceinline__ u32 vsra1(u32 x) { return ((x >> 1) & 0x7fff7fffu) | (x & 0x80008000u); }
__device__ __forceinline__ u32 pack2(int lo, int hi) { return ((u32)lo & 0xffffu) | ((u32)hi << 16); }
__device__ __forceinline__ int lo16(u32 x) { return (int)(int16_t)(x & 0xffffu); }
__device__ __forceinline__ int hi16(u32 x) { return (int)(int16_t)(x >> 16); }
__device__ __forceinline__ int sat16i(int v) { return max(-32768, min(32767, v)); }

// prmt with the sign-replication selector bit (__byte_perm only honours 3 bits per nibble)
__device__ __forceinline__ u32 prmt_sx(u32 a, u32 sel) {
  u32 d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
  return d;
}

// 8 int8 (4 steps x 2 lanes, C4 order) -> the 4 packed int16 pairs
__device__ __forceinline__ uint4 widen8(uint2 v) {
  return make_uint4(prmt_sx(v.x, 0x9180u), prmt_sx(v.x, 0xB3A2u), prmt_sx(v.y, 0x9180u), prmt_sx(v.y, 0xB3A2u));
}

}  // namespace oai
