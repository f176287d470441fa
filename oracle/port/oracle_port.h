/* TEST INFRASTRUCTURE -- CPU oracle for the LTE turbo-decoding hot path.
 *
 * Plain scalar C restatement of the reference algorithms (erlgo/openair4G,
 * openair1/PHY/CODING).  It is the checker for the CUDA product path and the
 * "port" CPU baseline in bench.py; it is never linked into or called from the
 * product library (openair4g_b200/).  Parity pin: every function here is
 * cross-checked against the reference sources compiled in place
 * (oracle/_ref/libref_oai.so, built by oracle/Makefile) and against the golden
 * vectors in tests/golden/ that were generated with that compiled reference.
 */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_CRC24_A 0
#define ORC_CRC24_B 1
#define ORC_CRC16   2
#define ORC_CRC8    3
#define ORC_LTE_NULL 2

/* QPP parameter table (188 rows); returns row index or -1. */
int  orc_qpp_index(int K);
int  orc_qpp_f1(int idx);
int  orc_qpp_f2(int idx);
int  orc_qpp_K(int idx);
/* pi[i] = (f1*i + f2*i*i) mod K for i<K */
int  orc_qpp_table(int K, uint16_t *pi);

/* CRC (crc_byte.c:116-207): result left-aligned in 32 bits like the reference */
uint32_t orc_crc24a(const uint8_t *in, int bitlen);
uint32_t orc_crc24b(const uint8_t *in, int bitlen);
uint32_t orc_crc16(const uint8_t *in, int bitlen);
uint32_t orc_crc8(const uint8_t *in, int bitlen);

/* segmentation parameters (lte_segmentation.c:52-134); returns 0 / -1 */
int orc_lte_segmentation(uint32_t B, uint32_t *C, uint32_t *Cplus, uint32_t *Cminus,
                         uint32_t *Kplus, uint32_t *Kminus, uint32_t *F);

/* rate (de)matching, lte_rate_matching.c */
uint32_t orc_generate_dummy_w(uint32_t D, uint8_t *w, uint8_t F);
int orc_lte_rate_matching_turbo_rx(uint32_t RTC, uint32_t G, int16_t *w, const uint8_t *dummy_w,
                                   const int16_t *soft_input, uint8_t C, uint32_t Nsoft,
                                   uint8_t Mdlharq, uint8_t Kmimo, uint8_t rvidx, uint8_t clear,
                                   uint8_t Qm, uint8_t Nl, uint8_t r, uint32_t *E_out);
void orc_sub_block_deinterleaving_turbo(uint32_t D, int16_t *d, const int16_t *w);

/* TX mirror, used only to make test vectors */
void orc_turbo_encode(const uint8_t *input, int nbytes, uint8_t *out_bits /* 3*K+12 */);
uint32_t orc_sub_block_interleaving_turbo(uint32_t D, const uint8_t *d, uint8_t *w);
uint32_t orc_lte_rate_matching_turbo(uint32_t RTC, uint32_t G, const uint8_t *w, uint8_t *e, uint8_t C,
                                     uint32_t Nsoft, uint8_t Mdlharq, uint8_t Kmimo, uint8_t rvidx,
                                     uint8_t Qm, uint8_t Nl, uint8_t r);

/* 16-bit decoder (3gpplte_turbo_decoder_sse_16bit.c:945-1385).  Returns what the
 * reference returns: iterations used, max+1 on failure, 255 on bad arguments. */
uint8_t orc_turbo_decoder16(const int16_t *y, uint8_t *decoded_bytes, uint16_t n,
                            uint8_t max_iterations, uint8_t crc_type, uint8_t F);
/* one MAP pass in the reference's lane layout, with optional alpha/beta dumps
 * (each 8*(n+16) int16, may be NULL) -- for kernel debugging */
void orc_log_map16(const int16_t *sys, const int16_t *par, int16_t *ext, int n, int term_flag,
                   int16_t *alpha_dump, int16_t *beta_dump);

/* CPU model of the optional sliding-window mode (td16_sw_port.c; not a reference function).  Same contract. */
uint8_t orc_turbo_decoder16_sw(const int16_t *y, uint8_t *decoded_bytes, uint16_t n, uint8_t max_iterations,
                               uint8_t crc_type, uint8_t F, int *dbg_llr);
int orc_sw_windows(int K);
int orc_sw_shift(const int16_t *y, int n);

/* 8-bit decoder (3gpplte_turbo_decoder_sse_8bit.c:894-1657); parity domain
 * n>=256 && n%16==0; returns 254 outside it. */
uint8_t orc_turbo_decoder8(const int16_t *y, uint8_t *decoded_bytes, uint16_t n,
                           uint8_t max_iterations, uint8_t crc_type, uint8_t F);

/* threaded batch drivers for the CPU baseline (OpenMP over code blocks, the
 * reference's own model: ulsch_decoding.c:1306-1310) */
void orc_turbo_decoder16_batch(const int16_t *y, int y_stride, uint8_t *out, int out_stride,
                               uint8_t *ret, int nblk, uint16_t n, uint8_t max_iterations,
                               uint8_t crc_type, int nthreads);

#ifdef __cplusplus
}
#endif
/* Gold sequence (LTE_REFSIG/lte_gold.c:151-180) and DL descrambling (LTE_TRANSPORT/dlsch_scrambling.c:99-138) */
uint32_t orc_lte_gold_generic(uint32_t *x1, uint32_t *x2, uint8_t reset);
void orc_gold_words(uint32_t c_init, uint32_t *words, int nwords);
void orc_dlsch_unscrambling(uint32_t c_init, int16_t *llr, int n);

/* Uplink front end (LTE_TRANSPORT/ulsch_decoding.c:381-1153): control sizes, descrambling + channel de-interleaver,
 * ACK / RI / CQI extraction, e fill, ACK / RI decisions */
typedef struct { uint32_t Qprime_RI, Qprime_ACK, Qprime_CQI, Q_RI, Q_CQI, G, H, Hprime, Hpp, Cmux, Rmux_prime; } orc_ul_sizes_t;
int orc_ulsch_control_sizes(uint32_t O_RI, uint32_t O_ACK, uint32_t Or1, uint32_t Msc_initial, uint32_t Nsymb_initial,
                            uint32_t beta_ri_x8, uint32_t beta_ack_x8, uint32_t beta_cqi_x8, uint32_t sumKr, uint32_t nb_rb,
                            uint32_t Qm, uint32_t Nsymb_pusch, orc_ul_sizes_t *o);
int orc_ulsch_front(const int16_t *llr, uint32_t c_init, uint32_t Qm, const orc_ul_sizes_t *z, uint32_t Ncp, uint32_t O_ACK,
                    uint32_t O_RI, uint32_t bundling, uint32_t Nbundled, int16_t *e, int16_t *q_ACK, int16_t *q_RI,
                    int8_t *q_cqi, uint8_t *o_ACK, uint8_t *o_RI);

#endif
