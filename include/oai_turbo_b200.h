/*
 * oai_turbo_b200.h -- C ABI of the B200-native LTE channel-decoding engine.
 *
 * Drop-in boundary for the turbo-decoding hot path of OpenAirInterface
 * (erlgo/openair4G).  Section 1 exports the reference's own entry points with the
 * reference's exact signatures, so ulsch_decoding.c / dlsch_decoding.c link against
 * this library instead of openair1/PHY/CODING/{3gpplte_turbo_decoder_sse_16bit.c,
 * 3gpplte_turbo_decoder_sse_8bit.c,lte_rate_matching.c}.  Section 2 is the new
 * batched, stream-ordered submit call that replaces the per-code-block loops
 * (ulsch_decoding.c:1222-1369, dlsch_decoding.c:303-453).  Section 3 is the
 * device-resident form used by the throughput harness (inputs already in HBM).
 *
 * All work runs on the GPU (hand-written sm_100a kernels).  There is no CPU
 * fallback: every entry point fails loudly (return code + message on stderr) when
 * no CUDA device is usable.
 */
#ifndef OAI_TURBO_B200_H
#define OAI_TURBO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The seven statistics arguments of the decoder calls.
 *
 * Inside the reference tree (a translation unit that also sees openair1/PHY/CODING/defs.h, e.g. ulsch_decoding.c or
 * dlsch_decoding.c): compile with -DOAI_TURBO_B200_WITH_REFERENCE_HEADERS (the CMake target in
 * cmake_targets/oai_turbo_b200 sets it) or simply include PHY/CODING/defs.h first.  This header then pulls in the
 * reference's own header, oai_time_stats_t IS its time_stats_t (openair1/PHY/TOOLS/time_meas.h:43-52), and the
 * prototypes of section 1 below are compatible re-declarations of defs.h:112,132,152,191-204,239-253,315-320,362,367,
 * 470-484,499-513 -- the compiler checks the two sets against each other.
 *
 * Stand-alone (no reference headers): a struct of the same layout.  Either way this library never dereferences the
 * pointers (the reference only touches them when its global opp_enabled != 0). */
#if defined(OAI_TURBO_B200_WITH_REFERENCE_HEADERS) || defined(__CODING_DEFS__H__)
#ifndef __CODING_DEFS__H__
#include "PHY/CODING/defs.h"
#endif
typedef time_stats_t oai_time_stats_t;
#define OAI_TIME_STATS_T_DEFINED
#elif !defined(OAI_TIME_STATS_T_DEFINED)
#define OAI_TIME_STATS_T_DEFINED
typedef struct {
  long long in, diff, diff_now, p_time, diff_square, max;
  int trials;
} oai_time_stats_t;
#endif

#define OAI_CRC24_A 0   /* openair1/PHY/CODING/defs.h:45-48 */
#define OAI_CRC24_B 1
#define OAI_CRC16   2
#define OAI_CRC8    3
#define OAI_LTE_NULL 2  /* openair1/PHY/CODING/defs.h:53 */

/* ------------------------------------------------------------------------------------
 * 1. Reference entry points (exact signatures)
 * ---------------------------------------------------------------------------------- */

/* openair1/PHY/CODING/defs.h:362,367 ; impl 3gpplte_turbo_decoder_sse_16bit.c:886-943,
 * 3gpplte_turbo_decoder_sse_8bit.c:834-892.  Creates the process-wide GPU context
 * (QPP tables for the 188 block sizes in HBM, CRC tables, streams, staging buffers).
 * Not thread-safe, call once (like the reference); the decode calls are re-entrant. */
void init_td16(void);
void free_td16(void);
void init_td8(void);
void free_td8(void);

/* openair1/PHY/CODING/defs.h:470-484 ; impl 3gpplte_turbo_decoder_sse_16bit.c:945-1385.
 * y: int16[3n+12] = (s,p1,p2) x n, then (x,z) x 3 of encoder 1 and of encoder 2.
 * decoded_bytes: n/8 bytes, MSB first.  Returns iterations used (2..max), max+1 when the
 * CRC never passed, 255 on crc_type > 3 or n not one of the 188 sizes.  f1/f2 are
 * accepted and ignored, like the reference (it looks n up, :1011-1018).
 * Bit-exact (bytes and return value) to the reference for every int16 input. */
unsigned char phy_threegpplte_turbo_decoder16(short *y, unsigned char *decoded_bytes,
    unsigned short n, unsigned short f1, unsigned short f2, unsigned char max_iterations,
    unsigned char crc_type, unsigned char F,
    oai_time_stats_t *init_stats, oai_time_stats_t *alpha_stats, oai_time_stats_t *beta_stats,
    oai_time_stats_t *gamma_stats, oai_time_stats_t *ext_stats, oai_time_stats_t *intl1_stats,
    oai_time_stats_t *intl2_stats);

/* openair1/PHY/CODING/defs.h:499-513 ; impl 3gpplte_turbo_decoder_sse_8bit.c:894-1657.
 * Parity domain: n >= 256 and n % 16 == 0 (outside it the reference
 * itself overruns its buffers); other sizes return 255. */
unsigned char phy_threegpplte_turbo_decoder8(short *y, unsigned char *decoded_bytes,
    unsigned short n, unsigned short f1, unsigned short f2, unsigned char max_iterations,
    unsigned char crc_type, unsigned char F,
    oai_time_stats_t *init_stats, oai_time_stats_t *alpha_stats, oai_time_stats_t *beta_stats,
    oai_time_stats_t *gamma_stats, oai_time_stats_t *ext_stats, oai_time_stats_t *intl1_stats,
    oai_time_stats_t *intl2_stats);

/* openair1/PHY/CODING/defs.h:152 ; impl lte_rate_matching.c:293-382.  Marks (never
 * clears) the NULL positions of the 3*Kpi circular buffer; returns RTC. */
uint32_t generate_dummy_w(uint32_t D, uint8_t *w, uint8_t F);

/* openair1/PHY/CODING/defs.h:239-253 ; impl lte_rate_matching.c:688-831.  w is the
 * caller-owned HARQ soft buffer (host memory, authoritative on every call): it is
 * read, accumulated into with int16 wrap-around, and written back.  Returns 0, or -1
 * when Kmimo, Mdlharq, C, Qm or Nl is zero. */
int lte_rate_matching_turbo_rx(uint32_t RTC, uint32_t G, int16_t *w, uint8_t *dummy_w,
    int16_t *soft_input, uint8_t C, uint32_t Nsoft, uint8_t Mdlharq, uint8_t Kmimo,
    uint8_t rvidx, uint8_t clear, uint8_t Qm, uint8_t Nl, uint8_t r, uint32_t *E_out);

/* openair1/PHY/CODING/defs.h:132 ; impl lte_rate_matching.c:193-243.  Writes d[-3*ND ..
 * 3*D+2] (callers pass &d[r][96], dlsch_decoding.c:380). */
void sub_block_deinterleaving_turbo(uint32_t D, int16_t *d, int16_t *w);

/* Parameter part of lte_segmentation (openair1/PHY/CODING/lte_segmentation.c:52-134; the
 * reference also copies the transport block into c[r], a TX-side job): C, C+, C-, K+, K-, F
 * for a transport block of B bits (CRC24A included).  Host integer rule, no kernel; the
 * batched callers use it to build their descriptors.  Returns 0, or -1 when C > 16 or
 * B'/C > 6144 like the reference.  Keeps the reference's quirk K- = B'/C - 8 for B'/C <= 512. */
int oai_lte_segmentation_params(uint32_t B, uint32_t *C, uint32_t *Cplus, uint32_t *Cminus,
                                uint32_t *Kplus, uint32_t *Kminus, uint32_t *F);

/* ------------------------------------------------------------------------------------
 * 2. Batched submit (new): all code blocks of a subframe / of many subframes and cells
 * ---------------------------------------------------------------------------------- */

typedef struct oai_turbo_batch oai_turbo_batch_t;   /* opaque, one per in-flight batch */

/* One code block.  With dematch_enable == 0, `in` is the decoder input y (int16[3K+12],
 * what the reference passes as &d[r][96]).  With dematch_enable != 0, `in` is this
 * block's slice of the rate-matched soft bits e (int16[E]) and the fused front end
 * (generate_dummy_w + lte_rate_matching_turbo_rx + sub_block_deinterleaving_turbo)
 * runs on the GPU first; `w` is the host HARQ buffer int16[3*Kpi] (read unless clear,
 * written back), may be NULL when clear != 0 and the caller does not keep it. */
typedef struct {
  const int16_t *in;
  uint8_t  *decoded_bytes;     /* K/8 bytes out (host) */
  uint8_t  *status;            /* 1 byte out: the decoder's return value */
  uint16_t  K;
  uint8_t   max_iterations;
  uint8_t   crc_type;
  uint8_t   F;
  uint8_t   llr8;              /* 0: 16-bit decoder, 1: 8-bit decoder */
  uint8_t   decode_enable;     /* 0: front end only (dlsch_decoding.c:417 err_flag) */
  uint8_t   dematch_enable;
  /* front-end parameters (lte_rate_matching_turbo_rx arguments) */
  int16_t  *w;
  uint32_t  G;
  uint32_t  Nsoft;
  uint8_t   C, r, rvidx, clear, Qm, Nl, Mdlharq, Kmimo;
  uint32_t  tb_id;             /* blocks with equal tb_id form one transport block */
  /* device-resident HARQ soft buffer (optional, front end only): when harq_pool != NULL the circular buffer w of
   * this block lives in slot harq_slot of the pool on the GPU, `w` is ignored and nothing of it crosses PCIe */
  struct oai_turbo_harq_pool *harq_pool;
  uint32_t  harq_slot;
  /* descrambling of the soft bits in the front end (optional; downlink: dlsch_unscrambling,
   * openair1/PHY/LTE_TRANSPORT/dlsch_scrambling.c:99-138, which the UE runs on dlsch_llr right before dlsch_decoding):
   * with scr_enable != 0, `in` holds the still scrambled soft bits of this block; scr_c_init is the codeword's c_init
   * (36.211 6.3.1: (rnti<<14) + (q<<13) + ((Ns>>1)<<9) + Nid_cell) and scr_offset the position of the block's first
   * soft bit in the codeword (the reference's r_offset).  Bit 0 of the sequence negates the LLR, like the reference. */
  uint32_t  scr_c_init;
  uint32_t  scr_offset;
  uint8_t   scr_enable;
  /* format of `in` for front-end blocks (dematch_enable != 0): 0 = int16 soft bits (the reference's type), 1 = int8
   * soft bits (int8_t[E] behind the pointer; the values a caller's demapper already clips to 8 bits for the 8-bit
   * decoder).  Same arithmetic on the device (the dematcher accumulates into int16 w); half the bytes on the host link.
   * Must be 0 when dematch_enable == 0. */
  uint8_t   in_fmt;
} oai_cb_desc_t;

/* HARQ soft-buffer pool in HBM.  The reference keeps w[r] (int16[3*Kpi]) per (UE, HARQ process, code block) in host
 * memory across retransmissions (LTE_TRANSPORT/defs.h:428,531) and rate dematching accumulates into it; a GPU front end
 * that treats the host copy as authoritative moves 2 x 3*Kpi int16 over PCIe per block and round.  A pool keeps
 * n_slots buffers of 3*Kpi(max_K) int16 on the device (zeroed at creation; `clear` in the descriptor resets a slot
 * like the reference's memset); the caller maps (cell, UE, harq_pid, r) to a slot index.  All pool-backed blocks of
 * one submit must use the same pool, on the GPU the batch runs on. */
typedef struct oai_turbo_harq_pool oai_turbo_harq_pool_t;
int oai_turbo_harq_pool_create(int gpu, uint32_t n_slots, uint16_t max_K, oai_turbo_harq_pool_t **pool);
/* copies the first n int16 of a slot to host memory (tests, migration of a UE to another GPU); synchronous */
int oai_turbo_harq_pool_read(oai_turbo_harq_pool_t *pool, uint32_t slot, int16_t *w_host, uint32_t n);
void oai_turbo_harq_pool_destroy(oai_turbo_harq_pool_t *pool);

#define OAI_BATCH_DL_STOP_AFTER_FAILURE 1u  /* dlsch_decoding.c:417,448-451: within a
        transport block, blocks after the first failing one report status 0xFE ("not
        decoded") and their decoded_bytes are zeroed */

#define OAI_BATCH_SLIDING_WINDOW 2u  /* OPTIONAL, NOT bit-exact with the reference: the 16-bit blocks of the batch are
        decoded by the sliding-window kernel (8...64 windows per block, stitched by next-iteration initialisation and
        32-step training recursions; soft bits scaled to 8 bits, extrinsic values clipped; one warp decodes a block out
        of shared memory in ONE launch).
        Same outputs and return-value rules; decoded bits / iteration counts may differ from
        phy_threegpplte_turbo_decoder16's near the decoding threshold -- the BLER delta against the default bit-exact
        mode is reported in profiles/ (tools/sw_bler_delta.py).  Never selected implicitly.  8-bit blocks are unaffected. */

/* Copies descriptors and inputs to the GPU and launches the whole pipeline on the
 * batch's stream; returns immediately (0) or a negative error.  gpu < 0: current device. */
int oai_turbo_submit_batch(const oai_cb_desc_t *cbs, int ncb, unsigned flags, int gpu,
                           oai_turbo_batch_t **handle);
/* Blocks until the batch is finished, scatters decoded_bytes/status/w back to the
 * host pointers of the descriptors and frees the handle. */
int oai_turbo_wait(oai_turbo_batch_t *handle);

/* Transport-block outputs.  A transport block is C consecutive descriptors of the cbs array, r = 0..C-1
 * in order.  After the decode, the GPU evaluates the reference's transport-block rules and assembles `b`:
 *   uplink == 0 : dlsch_decoding.c:417,448-451 (err_flag), :455-483 (return value), :486-512 (reassembly only when every
 *                 block passed; on a NACK *ret = 1 + max_iterations, *valid_bytes = 0 and b is left untouched);
 *   uplink != 0 : ulsch_decoding.c:1380-1409 (a failed block is skipped without advancing the offset; *ret = status of
 *                 the last passing block, or 1 + max_iterations once a block has failed).
 * b receives sum(Kr/8) - (F>>3) - (C > 1 ? 3*C : 0) bytes at most (filler bytes of block 0 skipped, CRC24B of every
 * block stripped when C > 1); one device->host copy per batch carries all transport blocks, so the descriptors of
 * their code blocks may leave decoded_bytes NULL (no per-block copy then).  max_iterations and F are taken from the
 * descriptor of block 0. */
/* Optional uplink front end of a transport block: the part of ulsch_decoding that builds e[] from the
 * demodulator's soft bits -- ulsch_decoding.c:600-733 (descrambling with c_init, placeholder handling, channel
 * de-interleaver), :775-873 (HARQ-ACK / RI soft sums), :877-1002 (CQI soft bits, e fill), :1052-1153 (ACK / RI decisions).
 * With it, `in` of the transport block's code-block descriptors is ignored (they must have dematch_enable != 0): e stays
 * on the GPU and block r reads its E soft bits at the running offset of ulsch_decoding.c:1259.
 * Sizes come from oai_ulsch_control_sizes().  llr holds Hpp*Qm soft bits, column by column of the interleaver matrix, as
 * rx_ulsch leaves them in lte_eNB_pusch_vars->llr.  Bit-exact to the reference when Hpp*Qm is a multiple of 32 (any
 * allocation with 12 data symbols); otherwise the reference multiplies its last soft bits by uninitialised stack
 * (ulsch_decoding.c:605-611 fills whole 32-bit words of the sequence only) and this library continues the sequence. */
typedef struct {
  const void *llr;           /* int16_t[Hpp*Qm], or int8_t[Hpp*Qm] with llr_fmt = 1 */
  uint8_t   llr_fmt;
  uint32_t  c_init;          /* (rnti<<14) + (subframe<<9) + Nid_cell, ulsch_decoding.c:296 */
  uint8_t   Qm;              /* 2, 4, 6 */
  uint8_t   Ncp;             /* 0 normal, 1 extended cyclic prefix (selects the RI / ACK column sets) */
  uint8_t   O_ACK;           /* 0, 1, 2 */
  uint8_t   O_RI;            /* 0, 1 */
  uint8_t   bundling;        /* ulsch->bundling */
  uint8_t   Nbundled;
  uint16_t  Cmux;            /* Nsymb_pusch */
  uint32_t  Qprime_RI, Qprime_ACK, Qprime_CQI, Hprime;   /* oai_ulsch_control_sizes */
  /* outputs (host memory, each may be NULL) */
  int16_t  *q_ACK;           /* 18 entries, after the bundling combination of :1059-1092 */
  int16_t  *q_RI;            /* 6 entries */
  int8_t   *q_cqi;           /* Qm*Qprime_CQI soft bits for the convolutional decoder (ulsch_harq->q) */
  uint8_t  *o_ACK;           /* 2 entries */
  uint8_t  *o_RI;            /* 1 entry */
} oai_ul_front_t;

/* Control-information sizes of one PUSCH allocation, ulsch_decoding.c:381-468 (host integer rule).  sumKr = sum of the
 * code-block sizes of the transport block; the beta offsets are the reference's *_times8 values.  G is what remains for
 * the data (the G of lte_rate_matching_turbo_rx), Hprime = (G + Qm*Qprime_CQI)/Qm, Hpp = Hprime + Qprime_RI.
 * Returns 0, or -1 when the control information does not fit (:449-452). */
int oai_ulsch_control_sizes(uint32_t O_RI, uint32_t O_ACK, uint32_t Or1, uint32_t Msc_initial, uint32_t Nsymb_initial,
                            uint32_t beta_offset_ri_times8, uint32_t beta_offset_harqack_times8, uint32_t beta_offset_cqi_times8,
                            uint32_t sumKr, uint32_t nb_rb, uint32_t Qm, uint32_t Nsymb_pusch,
                            uint32_t *Qprime_RI, uint32_t *Qprime_ACK, uint32_t *Qprime_CQI, uint32_t *G, uint32_t *Hprime,
                            uint32_t *Hpp);

typedef struct {
  uint32_t  first_cb;        /* index of block r = 0 in the cbs array */
  uint32_t  C;               /* 1..16 */
  uint8_t  *b;               /* out (host), may be NULL */
  uint32_t  b_capacity;      /* bytes available behind b */
  uint8_t  *ret;             /* out (host), may be NULL: the value dlsch_decoding / ulsch_decoding would return */
  uint32_t *valid_bytes;     /* out (host), may be NULL: bytes written to b (the reference's final `offset`) */
  uint8_t   uplink;
  const oai_ul_front_t *ul_front;   /* NULL: the code blocks bring their own soft bits */
} oai_tb_desc_t;

/* oai_turbo_submit_batch plus transport-block outputs; wait with oai_turbo_wait. */
int oai_turbo_submit_tbs(const oai_cb_desc_t *cbs, int ncb, const oai_tb_desc_t *tbs, int ntb, unsigned flags, int gpu,
                         oai_turbo_batch_t **handle);

/* Page-locked host memory for batch inputs / outputs.  Buffers from this allocator (or any
 * other cudaHostAlloc / cudaHostRegister memory) are copied from and to directly; pageable
 * memory goes through the library's own pinned staging area.  Large batches are pipelined:
 * the input copy of one part of the batch overlaps the decode of the previous part. */
void *oai_turbo_host_alloc(size_t bytes);
void oai_turbo_host_free(void *p);

/* ------------------------------------------------------------------------------------
 * 3. Device-resident decode (throughput mode: inputs already in HBM)
 * ---------------------------------------------------------------------------------- */

typedef struct oai_turbo_dev_plan oai_turbo_dev_plan_t;

/* Plans a batch of ncb equal-parameter code blocks whose inputs y live in device
 * memory at y_dev + i*y_stride (int16 units), outputs at out_dev + i*out_stride bytes
 * and status_dev[i].  Workspace is allocated once here. */
int oai_turbo_dev_plan_create(int ncb, uint16_t K, uint8_t max_iterations, uint8_t crc_type,
                              uint8_t llr8, oai_turbo_dev_plan_t **plan);
/* Enqueues one full decode of the batch on `stream` (a cudaStream_t passed as void*);
 * no host synchronisation.  Returns the number of kernels launched, or < 0. */
int oai_turbo_dev_decode(oai_turbo_dev_plan_t *plan, const int16_t *y_dev, long y_stride,
                         uint8_t *out_dev, long out_stride, uint8_t *status_dev, void *stream);
/* flags = OAI_BATCH_SLIDING_WINDOW: later decodes of the plan run in the optional sliding-window mode (NOT bit-exact,
 * see the flag); 0: back to the default bit-exact mode.  16-bit plans only. */
int oai_turbo_dev_plan_set_mode(oai_turbo_dev_plan_t *plan, unsigned flags);
void oai_turbo_dev_plan_destroy(oai_turbo_dev_plan_t *plan);
/* Per-launch CUDA-event timing of a plan (used by bench.py for the roofline figures).
 * enable != 0 switches it on for subsequent decodes.  When ms4/count4 are non-NULL the
 * totals accumulated so far, per kernel class {demux, MAP, exchange-1, exchange-2}, are
 * returned and reset; the plan's stream must be idle when this is called. */
int oai_turbo_dev_plan_profile(oai_turbo_dev_plan_t *plan, int enable, double *ms4, long *count4);

/* ------------------------------------------------------------------------------------
 * Section 3b.  Transmit-side mirror: test-vector generation on the GPU.
 * The three reference entry points with their exact signatures, and one batched call.
 * ---------------------------------------------------------------------------------- */

/* openair1/PHY/CODING/defs.h:315-320 ; impl 3gpplte_sse.c:380-476.  output = 3*8*input_length_bytes + 12
 * bytes holding one bit each: (systematic, parity 1, parity 2) per input bit, then the 12 termination
 * bits.  F, f1 and f2 are accepted and ignored like in the reference (the interleaver follows from the
 * length; filler bits are ordinary zeros of the input).  Illegal lengths print "Illegal frame length!"
 * and leave output untouched.  For odd byte counts the reference's byte-pair interleaver leaves its last
 * interleaved byte unwritten (3gpplte_sse.c:321), i.e. encodes stack garbage; this call encodes the
 * 36.212 5.1.3.2 code for every length. */
void threegpplte_turbo_encoder(uint8_t *input, uint16_t input_length_bytes, uint8_t *output, uint8_t F,
                               uint16_t interleaver_f1, uint16_t interleaver_f2);

/* openair1/PHY/CODING/defs.h:112 ; impl lte_rate_matching.c:51-130.  Like the reference it READS the
 * 3*ND bytes in front of d (the callers keep LTE_NULL there, dlsch_coding.c:205) and WRITES
 * d[3*D+2] = d[2].  Returns RTC. */
uint32_t sub_block_interleaving_turbo(uint32_t D, uint8_t *d, uint8_t *w);

/* openair1/PHY/CODING/defs.h:191-204 ; impl lte_rate_matching.c:464-634.  Returns E, or 0 (and prints
 * the reference's message) when the soft buffer is smaller than the circular buffer. */
uint32_t lte_rate_matching_turbo(uint32_t RTC, uint32_t G, uint8_t *w, uint8_t *e, uint8_t C,
                                 uint32_t Nsoft, uint8_t Mdlharq, uint8_t Kmimo, uint8_t rvidx, uint8_t Qm,
                                 uint8_t Nl, uint8_t r, uint8_t nb_rb, uint8_t m);

/* One code block of the batched TX call = the body of the per-segment loops of dlsch_encoding /
 * ulsch_encoding (dlsch_coding.c:356-404, ulsch_coding.c:400-550): encoder -> sub-block interleaver ->
 * rate matching, without storing d or w. */
typedef struct {
  const uint8_t *c;     /* K/8 info bytes incl. CRC and leading filler zeros, MSB first (harq->c[r]) */
  uint8_t *e;           /* out: E bits, one per byte (harq->e + r_offset) */
  uint32_t G, Nsoft;
  uint32_t E;           /* out: bits written (0: soft buffer smaller than the circular buffer) */
  uint16_t K;
  uint8_t F;            /* filler bits of this block (first block of a transport block only) */
  uint8_t filler_null;  /* 0: the reference's TX, which transmits the filler bits; 1: 36.212 5.1.4.1.1, filler bits
                           of streams 0/1 are <NULL> -- what generate_dummy_w assumes on the receive side */
  uint8_t C, Mdlharq, Kmimo, rvidx, Qm, Nl, r, reserved;
} oai_tx_desc_t;

#define OAI_TX_DEVICE_POINTERS 1u   /* c and e are device pointers on `gpu` */
/* Synchronous.  Returns 0, or < 0 with oai_turbo_b200_last_error(). */
int oai_turbo_tx_batch(oai_tx_desc_t *blocks, int n, unsigned flags, int gpu);

/* ------------------------------------------------------------------------------------
 * 4. Introspection
 * ---------------------------------------------------------------------------------- */
const char *oai_turbo_b200_version(void);
/* last error message of the calling thread ("" if none) */
const char *oai_turbo_b200_last_error(void);
/* number of kernels this library has launched since load (all threads) */
unsigned long long oai_turbo_b200_launch_count(void);
/* Kernel-level test hook: one max-log-MAP pass (the reference's log_map16,
 * 3gpplte_turbo_decoder_sse_16bit.c:84-119) of a single block on the GPU.  y as for the
 * decoder; the systematic input is y's systematic stream, the parity stream p1 (term=0,
 * tail from y[3K..3K+5]) or p2 (term=1, tail from y[3K+6..3K+11]).  policy 0: guard decides,
 * 1: force the non-saturating fast path, 2: force the exact saturating path.
 * ext_out: K int16 in the reference's lane layout [step*8 + lane]. */
int oai_turbo_debug_map16(const int16_t *y, uint16_t K, int term, int policy, int16_t *ext_out);

#ifdef __cplusplus
}
#endif
#endif /* OAI_TURBO_B200_H */
