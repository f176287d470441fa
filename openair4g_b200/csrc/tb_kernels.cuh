// Transport-block reassembly and the transport-block return value on the GPU.
//
//   downlink: dlsch_decoding.c:417,448-451 (blocks after the first failing one are not decoded: c[r] zeroed),
//             :455-483 (return value), :486-512 (reassembly, only when every block passed)
//   uplink  : ulsch_decoding.c:1380-1409 (failed blocks are skipped WITHOUT advancing the offset; ret = status of the last
//             passing block unless some block failed)
//
// One CTA per transport block.  Input: the decoder outputs of its C code blocks (d_out at out_off, K/8 bytes each) and
// their status bytes; output: the transport block b (filler bytes of block 0 skipped, CRC24B stripped when C > 1), one
// return value and the number of valid bytes.  With that, one D2H of TBS/8 + 3 bytes and one status byte per transport
// block replace C per-block copies and the host post-pass.
#pragma once
#include "td_common.cuh"

namespace oai {

constexpr int TB_THREADS = 256;
constexpr int TB_MAX_C = 16;              // MAX_NUM_DLSCH_SEGMENTS / MAX_NUM_ULSCH_SEGMENTS (LTE_TRANSPORT/defs.h:67,69)

struct TbMeta {
  uint32_t first;        // index of block r = 0 in the TbBlk array
  uint32_t C;
  uint32_t F;            // filler bits of block 0
  uint32_t b_off;        // byte offset of this transport block in the TB output pool
  uint8_t  uplink;       // 0: dlsch_decoding rule, 1: ulsch_decoding rule
  uint8_t  max_iter;
  uint8_t  stop_after_failure;   // downlink: mark the blocks after the first failing one 0xFE and zero their bytes
  uint8_t  pad;
};
struct TbBlk {
  int32_t  gpu_idx;      // index into the batch status array, -1: the block never reached the GPU (illegal / disabled)
  uint32_t out_off;      // byte offset of its decoded bytes in the batch output
  uint32_t kbytes;       // K/8
};
struct TbResult { uint32_t valid_bytes; uint8_t ret; uint8_t pad[3]; };

__global__ void __launch_bounds__(TB_THREADS) k_tb_assemble(const TbMeta* tbs, int ntb, const TbBlk* blks, uint8_t* d_out,
                                                            uint8_t* d_status, uint8_t* tb_pool, TbResult* res) {
  __shared__ uint32_t s_src[TB_MAX_C], s_dst[TB_MAX_C], s_len[TB_MAX_C];
  __shared__ uint8_t s_zero[TB_MAX_C];
  const int tb = blockIdx.x;
  if (tb >= ntb) return;
  const TbMeta m = tbs[tb];
  const TbBlk* bl = blks + m.first;
  if (threadIdx.x == 0) {
    const uint32_t fail = 1u + m.max_iter, strip = (m.C > 1) ? 3u : 0u;
    uint32_t offset = 0, ret = m.uplink ? 1u : m.max_iter;          // ulsch_decoding.c:1384 / dlsch_decoding.c:253
    bool err = false;
    for (uint32_t r = 0; r < m.C; ++r) {
      const TbBlk b = bl[r];
      uint32_t st = (b.gpu_idx >= 0) ? d_status[b.gpu_idx] : 255u;
      s_zero[r] = 0; s_len[r] = 0;
      if (!m.uplink) {
        if (err) {                                                  // not decoded: c[r] stays zeroed (:400,417)
          if (m.stop_after_failure) { if (b.gpu_idx >= 0) d_status[b.gpu_idx] = 0xFE; s_zero[r] = 1; }
          continue;
        }
        ret = st;                                                   // the last decoded block's value (:424-440)
        if (st >= fail) { err = true; continue; }                   // :448-451 (0xFE / 255 count as failures)
      } else {
        if (st >= fail) { ret = fail; continue; }                   // :1404-1406 (the offset does not advance)
        if (ret != fail) ret = st;                                  // :1401-1402
      }
      const uint32_t skip = (r == 0) ? (m.F >> 3) : 0u;
      s_src[r] = b.out_off + skip;
      s_dst[r] = offset;
      s_len[r] = b.kbytes - skip - strip;
      offset += s_len[r];
    }
    if (!m.uplink && err) {                                         // NACK: no reassembly (:455-469)
      ret = fail; offset = 0;
      for (uint32_t r = 0; r < m.C; ++r) s_len[r] = 0;
    }
    res[tb].valid_bytes = offset;
    res[tb].ret = (uint8_t)ret;
  }
  __syncthreads();
  uint8_t* dst = tb_pool + m.b_off;
  for (uint32_t r = 0; r < m.C; ++r) {
    const uint32_t n = s_len[r];
    const uint8_t* src = d_out + s_src[r];
    uint8_t* d = dst + s_dst[r];
    for (uint32_t i = threadIdx.x; i < n; i += TB_THREADS) d[i] = src[i];
    if (s_zero[r]) {
      uint8_t* z = d_out + bl[r].out_off;
      for (uint32_t i = threadIdx.x; i < bl[r].kbytes; i += TB_THREADS) z[i] = 0;
    }
  }
}

}  // namespace oai
