/* TEST INFRASTRUCTURE (see oracle_port.h).
 * Uplink front of the rate dematcher: control-information sizes, descrambling, channel de-interleaving, HARQ-ACK / RI /
 * CQI extraction and the fill of the data soft bits e[] -- reference: openair1/PHY/LTE_TRANSPORT/ulsch_decoding.c:381-468
 * (sizes), :600-733 (Gold sequence, placeholder handling, de-interleaver), :775-873 (q_ACK, q_RI), :877-1002 (CQI soft
 * bits, e), :1052-1153 (ACK / RI decisions).  Restated index by index instead of with the reference's running pointers;
 * the quirks that define the result are kept:
 *   - the scrambling signs of a placeholder symbol are patched IN THE SEQUENCE before the multiply: y-placeholder =
 *     sign of the symbol's first bit, x-placeholders = -1 (:647-700);
 *   - the products c * llr are stored as int16 (so -(-32768) stays -32768), the q_ACK / q_RI sums wrap in int16;
 *   - HARQ-ACK positions are zeroed in y AFTER they were accumulated, and stay in e as zeros (punctured data);
 *   - the tag index `j` of the CQI / data walk (:877-1002) only advances while it sees a tagged (RI) symbol and is NOT
 *     advanced when a symbol is consumed, so it stops at the first untagged symbol: with L = number of RI-tagged symbols
 *     at the very start of the row-major matrix (0 unless the allocation is one row high), CQI takes symbols
 *     [L, L + Q'_CQI) and e the (H' - Q'_CQI) symbols after them -- RI symbols further down are NOT skipped;
 *   - the reference fills its sign array only for floor(H''*Qm / 32) whole words of the sequence (:605-611): when
 *     H''*Qm is not a multiple of 32 its last signs are uninitialised stack.  Parity domain: H''*Qm mod 32 == 0
 *     (every allocation with 12 data symbols; with 11 or fewer some PRB counts fall outside).  Outside it this port
 *     continues the sequence. */
#include <stdlib.h>
#include "oracle_port.h"

static const uint8_t CS_RI[2][4] = {{1, 4, 7, 10}, {0, 3, 5, 8}};     /* 36.212 table 5.2.2.8-1 (normal, extended CP) */
static const uint8_t CS_ACK[2][4] = {{2, 3, 8, 9}, {1, 2, 6, 7}};     /* 36.212 table 5.2.2.8-2 */
static const int8_t WACK_RX[5][4] = {{-1, -1, -1, -1}, {-1, 1, -1, 1}, {-1, -1, 1, 1}, {-1, 1, 1, -1}, {1, 1, 1, 1}};

static uint32_t ceil_div_cap(uint32_t num, uint32_t den, uint32_t cap, int use_cap)
{
  uint32_t q;
  if (num == 0) return 0;
  q = (num % den) ? 1 + num / den : num / den;
  return (use_cap && q > cap) ? cap : q;
}

/* ulsch_decoding.c:381-468.  Returns 0, or -1 when G would be negative (:449-452). */
int orc_ulsch_control_sizes(uint32_t O_RI, uint32_t O_ACK, uint32_t Or1, uint32_t Msc_initial, uint32_t Nsymb_initial,
                            uint32_t beta_ri_x8, uint32_t beta_ack_x8, uint32_t beta_cqi_x8, uint32_t sumKr, uint32_t nb_rb,
                            uint32_t Qm, uint32_t Nsymb_pusch, orc_ul_sizes_t *o)
{
  const uint32_t Gtot = nb_rb * (12 * Qm) * Nsymb_pusch, L = (Or1 < 12) ? 0 : 8;
  uint32_t Qp;
  o->Qprime_RI = ceil_div_cap(O_RI * Msc_initial * Nsymb_initial * beta_ri_x8, 8 * sumKr, 4 * nb_rb * 12, 1);
  o->Qprime_ACK = ceil_div_cap(O_ACK * Msc_initial * Nsymb_initial * beta_ack_x8, 8 * sumKr, 4 * nb_rb * 12, 1);
  Qp = (Or1 > 0) ? ceil_div_cap((Or1 + L) * Msc_initial * Nsymb_initial * beta_cqi_x8, 8 * sumKr, 0, 0) : 0;
  if (Qp > Gtot - O_RI) Qp = Gtot - O_RI;                                 /* :438-439 (sic: bits against symbols) */
  o->Qprime_CQI = Qp;
  o->Q_RI = Qm * o->Qprime_RI;
  o->Q_CQI = Qm * Qp;
  o->G = Gtot - o->Q_RI - o->Q_CQI;
  if ((int32_t)o->G < 0) return -1;
  o->H = o->G + o->Q_CQI;
  o->Hprime = o->H / Qm;
  o->Hpp = o->Hprime + o->Qprime_RI;
  o->Cmux = Nsymb_pusch;
  o->Rmux_prime = o->Hpp / o->Cmux;
  return 0;
}

/* is symbol (row r, column col) the i-th placeholder of a set of n (column set cs)?  The reference walks i = 0..n-1 with
 * row = R' - 1 - (i >> 2) and column cs[j], j = 0, 3, 2, 1, 0, ... (:641-660).  Returns i or -1. */
static int placeholder_index(uint32_t r, uint32_t col, uint32_t n, const uint8_t *cs, uint32_t Rp)
{
  uint32_t jj, i;
  for (jj = 0; jj < 4; jj++)
    if (cs[jj] == col) break;
  if (jj == 4 || r >= Rp) return -1;
  i = 4 * (Rp - 1 - r) + ((4 - jj) & 3);
  return (i < n) ? (int)i : -1;
}

int orc_ulsch_front(const int16_t *llr, uint32_t c_init, uint32_t Qm, const orc_ul_sizes_t *z, uint32_t Ncp, uint32_t O_ACK,
                    uint32_t O_RI, uint32_t bundling, uint32_t Nbundled, int16_t *e, int16_t *q_ACK, int16_t *q_RI,
                    int8_t *q_cqi, uint8_t *o_ACK, uint8_t *o_RI)
{
  const uint32_t Rp = z->Rmux_prime, Cm = z->Cmux, nsym = Rp * Cm, nbit = nsym * Qm;
  const uint8_t *cs_ri = CS_RI[Ncp ? 1 : 0], *cs_ack = CS_ACK[Ncp ? 1 : 0];
  uint32_t len_ACK = 0, len_RI = 0, i, q, sym, L;
  uint32_t *gold = (uint32_t *)malloc(sizeof(uint32_t) * (nbit / 32 + 2));
  int16_t *y = (int16_t *)malloc(sizeof(int16_t) * (nbit + 8));
  if (O_ACK > 2 || O_RI > 1) { free(gold); free(y); return -1; }          /* :817-820, :857-860 */
  if (O_ACK == 1) len_ACK = Qm;
  if (O_ACK == 2) len_ACK = 3 * Qm;
  if (O_RI == 1) len_RI = Qm;
  orc_gold_words(c_init, gold, (int)(nbit / 32 + 1));
  /* y in row-major symbol order: y[Qm*(r*Cmux + col) + q] <- sign * llr[(col*R' + r)*Qm + q] (:703-758) */
  for (sym = 0; sym < nsym; sym++) {
    const uint32_t r = sym / Cm, col = sym % Cm, in0 = (col * Rp + r) * Qm;
    const int is_ri = placeholder_index(r, col, z->Qprime_RI, cs_ri, Rp) >= 0;
    const int is_ack = placeholder_index(r, col, z->Qprime_ACK, cs_ack, Rp) >= 0;
    int sign[6] = {1, 1, 1, 1, 1, 1};
    for (q = 0; q < Qm; q++) sign[q] = 2 * (int)((gold[(in0 + q) >> 5] >> ((in0 + q) & 31)) & 1) - 1;
    if (is_ri) {                                                          /* :641-660 */
      sign[1] = sign[0];
      for (q = 2; q < Qm; q++) sign[q] = -1;
    }
    if (is_ack) {                                                         /* :662-697 (after the RI pass) */
      if (O_ACK == 1) {
        if (bundling == 0) sign[1] = sign[0];
        for (q = 2; q < Qm; q++) sign[q] = -1;
      } else if (O_ACK == 2) {
        for (q = 2; q < Qm; q++) sign[q] = -1;
      }
    }
    for (q = 0; q < Qm; q++) y[sym * Qm + q] = (int16_t)(sign[q] * (int)llr[in0 + q]);
  }
  /* HARQ-ACK: accumulate, then null the positions (:775-840) */
  for (i = 0; i < len_ACK; i++) q_ACK[i] = 0;
  for (i = 0; i < z->Qprime_ACK; i++) {
    const uint32_t r = Rp - 1 - (i >> 2), col = cs_ack[(4 - (i & 3)) & 3];
    for (q = 0; q < Qm; q++) {
      int16_t *p = &y[Qm * (r * Cm + col) + q];
      q_ACK[(q + Qm * i) % len_ACK] = (int16_t)(q_ACK[(q + Qm * i) % len_ACK] + *p);
      *p = 0;
    }
  }
  /* RI: accumulate; the positions stay in y (:843-873) */
  for (i = 0; i < len_RI; i++) q_RI[i] = 0;
  for (i = 0; i < z->Qprime_RI; i++) {
    const uint32_t r = Rp - 1 - (i >> 2), col = cs_ri[(4 - (i & 3)) & 3];
    for (q = 0; q < Qm; q++) q_RI[(q + Qm * i) % len_RI] = (int16_t)(q_RI[(q + Qm * i) % len_RI] + y[Qm * (r * Cm + col) + q]);
  }
  /* leading run of RI-tagged symbols: the only symbols the CQI / data walk ever skips (see the header) */
  for (L = 0; L < nsym; L++)
    if (placeholder_index(L / Cm, L % Cm, z->Qprime_RI, cs_ri, Rp) < 0) break;
  for (i = 0; i < z->Qprime_CQI * Qm; i++) {                              /* :877-921, saturated to int8 */
    const int v = y[L * Qm + i];
    q_cqi[i] = (int8_t)(v > 127 ? 127 : (v < -128 ? -128 : v));
  }
  for (i = 0; i < (z->Hprime - z->Qprime_CQI) * Qm; i++) e[i] = y[(L + z->Qprime_CQI) * Qm + i];     /* :925-1002 */
  /* decisions (:1052-1153) */
  {
    const int8_t *wa = WACK_RX[(bundling == 0) ? 4 : ((Nbundled - 1) & 3)];
    if (O_ACK == 1) {
      q_ACK[0] = (int16_t)(q_ACK[0] * wa[0]);
      q_ACK[0] = (int16_t)(q_ACK[0] + ((bundling == 0) ? q_ACK[1] * wa[0] : q_ACK[1] * wa[1]));
      o_ACK[0] = (q_ACK[0] < 0) ? 0 : 1;
    }
    if (O_ACK == 2) {
      const int a = (Qm == 2) ? 3 : (Qm == 4 ? 5 : 7), b = (Qm == 2) ? 4 : (Qm == 4 ? 8 : 12), c2 = (Qm == 2) ? 2 : (Qm == 4 ? 4 : 6),
                d = (Qm == 2) ? 5 : (Qm == 4 ? 9 : 13);
      int m, mn;
      q_ACK[0] = (int16_t)(q_ACK[0] * wa[0] + q_ACK[a] * wa[1]);
      q_ACK[1] = (int16_t)(q_ACK[1] * wa[0] + q_ACK[b] * wa[1]);
      q_ACK[2] = (int16_t)(q_ACK[c2] * wa[0] + q_ACK[d] * wa[1]);
      o_ACK[0] = 1; o_ACK[1] = 1;
      m = q_ACK[0] + q_ACK[1] - q_ACK[2];
      mn = -q_ACK[0] + q_ACK[1] + q_ACK[2];
      if (mn > m) { o_ACK[0] = 0; o_ACK[1] = 1; m = mn; }
      mn = q_ACK[0] - q_ACK[1] + q_ACK[2];
      if (mn > m) { o_ACK[0] = 1; o_ACK[1] = 0; m = mn; }
      mn = -q_ACK[0] - q_ACK[1] - q_ACK[2];
      if (mn > m) { o_ACK[0] = 0; o_ACK[1] = 0; m = mn; }
    }
    if (O_RI == 1 && z->Qprime_RI > 0) o_RI[0] = ((q_RI[0] + q_RI[Qm / 2]) > 0) ? 0 : 1;
  }
  free(gold);
  free(y);
  return 0;
}
