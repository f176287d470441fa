/* TEST INFRASTRUCTURE: shadows openair1/PHY/defs.h for a build-time copy of LTE_TRANSPORT/ulsch_decoding.c (the real
 * header needs the ASN.1-generated RRC headers, which this tree cannot produce).  Declares only what that translation
 * unit touches: the members of the eNB / ULSCH / HARQ structures it reads and writes (names as in
 * openair1/PHY/LTE_TRANSPORT/defs.h:370-470 and openair1/PHY/defs.h), sized like there.  Contains no reference code. */
#ifndef ORACLE_SHIM4_PHY_DEFS_H
#define ORACLE_SHIM4_PHY_DEFS_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifndef NO_OPENAIR1
#define NO_OPENAIR1
#endif
typedef int lte_prefix_type_t;
#include "PHY/CODING/defs.h"          /* the reference's own prototypes + time_stats_t (via PHY/TOOLS/time_meas.h) */
unsigned int lte_gold_generic(unsigned int *x1, unsigned int *x2, unsigned char reset);   /* LTE_REFSIG/defs.h:45 */

#define msg(...) ((void)0)
#define LOG_E(c, ...) ((void)0)
#define LOG_D(c, ...) ((void)0)
#define LOG_N(c, ...) ((void)0)
#define LOG_I(c, ...) ((void)0)
#define malloc16(x) aligned_alloc(16, ((x) + 15) & ~(size_t)15)
#define free16(p, n) free(p)

#define NSOFT 1827072
#define MAX_NUM_DLSCH_SEGMENTS 16
#define MAX_NUM_ULSCH_SEGMENTS MAX_NUM_DLSCH_SEGMENTS
#define MAX_ULSCH_PAYLOAD_BYTES (MAX_NUM_ULSCH_SEGMENTS*768)
#define MAX_NUM_CHANNEL_BITS (14*1200*6)
#define MAX_CQI_BITS 128
#define MAX_CQI_BYTES 16
#define MAX_CQI_PAYLOAD (MAX_CQI_BITS*20)
#define MAX_ACK_PAYLOAD 18
#define MAX_RI_PAYLOAD 6

typedef uint32_t frame_t;
typedef int UCI_format_t;
typedef int mod_sym_t;
typedef struct { uint8_t Ncp; uint16_t Nid_cell; uint8_t tdd_config; } LTE_DL_FRAME_PARMS;

typedef struct {
  uint8_t Ndi, status, subframe_scheduling_flag, phich_active, phich_ACK;
  uint16_t nb_rb;
  uint32_t TBS, B;
  uint8_t cqi_crc_status;
  uint8_t o[MAX_CQI_BYTES];
  uint8_t Or1, Or2, o_RI[2], O_RI, o_ACK[4], O_ACK;
  int8_t q[MAX_CQI_PAYLOAD];
  int8_t o_w[(MAX_CQI_BITS+8)*3];
  int8_t o_d[96+((MAX_CQI_BITS+8)*3)];
  int16_t q_ACK[MAX_ACK_PAYLOAD];
  int16_t q_RI[MAX_RI_PAYLOAD];
  int16_t e[MAX_NUM_CHANNEL_BITS];
  uint8_t *b;
  uint8_t *c[MAX_NUM_ULSCH_SEGMENTS];
  uint32_t RTC[MAX_NUM_ULSCH_SEGMENTS];
  uint8_t Nsymb_pusch, round, mcs, rvidx;
  int16_t w[MAX_NUM_ULSCH_SEGMENTS][3*(6144+64)];
  int16_t *d[MAX_NUM_ULSCH_SEGMENTS];
  uint32_t C, Cminus, Cplus, Kminus, Kplus, F;
  uint8_t Nl;
  uint16_t Msc_initial;
  uint8_t Nsymb_initial;
} LTE_UL_eNB_HARQ_t;

typedef struct {
  LTE_UL_eNB_HARQ_t *harq_processes[8];
  uint8_t Mdlharq, max_turbo_iterations, RRCConnRequest_flag, bundling, O_RI, Or1, o_RI[2];
  uint16_t beta_offset_cqi_times8, beta_offset_ri_times8, beta_offset_harqack_times8;
  uint16_t rnti;
  int16_t *e;
} LTE_eNB_ULSCH_t;

typedef struct { int16_t *llr; } LTE_eNB_PUSCH;
typedef struct { int subframe_rx; frame_t frame_rx; } eNB_proc_t;
typedef struct {
  uint8_t Mod_id;
  LTE_DL_FRAME_PARMS lte_frame_parms;
  LTE_eNB_PUSCH *lte_eNB_pusch_vars[4];
  LTE_eNB_ULSCH_t *ulsch_eNB[4];
  eNB_proc_t proc[10];
  time_stats_t ulsch_deinterleaving_stats, ulsch_demultiplexing_stats, ulsch_rate_unmatching_stats, ulsch_turbo_decoding_stats,
               ulsch_tc_alpha_stats, ulsch_tc_beta_stats, ulsch_tc_ext_stats, ulsch_tc_gamma_stats, ulsch_tc_init_stats,
               ulsch_tc_intl1_stats, ulsch_tc_intl2_stats;
} PHY_VARS_eNB;
typedef struct oracle_opaque_ue PHY_VARS_UE;
typedef struct oracle_opaque_rn PHY_VARS_RN;

uint8_t get_Qm_ul(uint8_t I_MCS);
uint8_t subframe2harq_pid(LTE_DL_FRAME_PARMS *frame_parms, frame_t frame, uint8_t subframe);
/* convolutional-code side of the CQI path (not on the pinned path: the driver keeps Or1 = 0 or stubs these) */
uint32_t generate_dummy_w_cc(uint32_t D, uint8_t *w);
void phy_viterbi_lte_sse2(int8_t *y, uint8_t *decoded_bytes, uint16_t n);
extern unsigned short f1f2mat_old[2*188];
#endif
