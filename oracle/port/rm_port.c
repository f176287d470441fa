/* TEST INFRASTRUCTURE (see oracle_port.h).
 * Turbo rate (de)matching and sub-block (de)interleaving, restated from
 * openair1/PHY/CODING/lte_rate_matching.c:
 *   sub_block_interleaving_turbo   :51-130   (TX, vector generation only)
 *   sub_block_deinterleaving_turbo :193-243
 *   generate_dummy_w               :293-382
 *   lte_rate_matching_turbo        :464-634  (TX, vector generation only)
 *   lte_rate_matching_turbo_rx     :688-831
 * Column permutation = 5-bit bit reversal (36.212 table 5.1.4-1; reference table :42). */
#include <string.h>
#include "oracle_port.h"

static uint32_t brev5(uint32_t c)
{
  return ((c & 1) << 4) | ((c & 2) << 2) | (c & 4) | ((c & 8) >> 2) | ((c & 16) >> 4);
}

static uint32_t rtc_of(uint32_t D) { return (D >> 5) + ((D & 31) ? 1 : 0); }

/* RX: w (three sub-blocks, column-major after the permutation) -> d triples.
 * d may be written at negative offsets down to d[-3*ND] (callers pass &buf[96]) and
 * stream 2 is shifted by one position: its entry lands at triple index+1, slot 2
 * (reference :216-231, `d3 = d1+5`). */
void orc_sub_block_deinterleaving_turbo(uint32_t D, int16_t *d, const int16_t *w)
{
  uint32_t RTC = rtc_of(D), Kpi = RTC << 5, ND = Kpi - D, col, row, k = 0;
  int16_t *d1 = d - 3 * (int32_t)ND;
  for (col = 0; col < 32; col++) {
    uint32_t idx = brev5(col);
    for (row = 0; row < RTC; row++, k++, idx += 32) {
      d1[3 * idx]     = w[k];
      d1[3 * idx + 1] = w[Kpi + 2 * k];
      d1[3 * idx + 5] = w[Kpi + 2 * k + 1];
    }
  }
}

/* Marks NULL positions only (never clears; callers zero the buffer first,
 * dlsch_decoding.c:331). */
uint32_t orc_generate_dummy_w(uint32_t D, uint8_t *w, uint8_t F)
{
  uint32_t RTC = rtc_of(D), Kpi = RTC << 5, ND = Kpi - D, col, k = 0;
  for (col = 0; col < 32; col++, k += RTC) {
    uint32_t idx = brev5(col), k2 = k << 1;
    if (idx < ND + F)        { w[k] = ORC_LTE_NULL;     w[Kpi + k2] = ORC_LTE_NULL; }
    if (idx + 32 < ND + F)   { w[k + 1] = ORC_LTE_NULL; w[Kpi + 2 + k2] = ORC_LTE_NULL; }
    if (idx + 64 < ND + F)   { w[k + 2] = ORC_LTE_NULL; w[Kpi + 4 + k2] = ORC_LTE_NULL; }
    if (idx + 1 < ND)        { w[Kpi + 1 + k2] = ORC_LTE_NULL; }
  }
  if (ND > 0) w[3 * Kpi - 1] = ORC_LTE_NULL;
  return RTC;
}

static uint32_t e_for_block(uint32_t G, uint8_t C, uint8_t Qm, uint8_t Nl, uint8_t r)
{
  uint32_t Gp = G / Nl / Qm, GpmodC = Gp % C;
  if (r < (uint32_t)(C - GpmodC)) return Nl * Qm * (Gp / C);
  return Nl * Qm * ((GpmodC == 0 ? 0 : 1) + (Gp / C));
}

static void rm_params(uint32_t RTC, uint8_t C, uint32_t Nsoft, uint8_t Mdlharq, uint8_t Kmimo,
                      uint8_t rvidx, uint32_t *Ncb, uint32_t *k0)
{
  uint32_t m = Mdlharq < 8 ? Mdlharq : 8;
  uint32_t Nir = Nsoft / Kmimo / m;
  uint32_t ncb = Nir / C, full = 3 * (RTC << 5);
  if (full < ncb) ncb = full;
  *Ncb = ncb;
  *k0 = RTC * (2 + (rvidx * (((ncb % (RTC << 3)) == 0 ? 0 : 1) + (ncb / (RTC << 3))) * 2));
}

/* int16 accumulation WRAPS (plain `+=` on int16_t, reference :749,765). */
int orc_lte_rate_matching_turbo_rx(uint32_t RTC, uint32_t G, int16_t *w, const uint8_t *dummy_w,
                                   const int16_t *soft_input, uint8_t C, uint32_t Nsoft,
                                   uint8_t Mdlharq, uint8_t Kmimo, uint8_t rvidx, uint8_t clear,
                                   uint8_t Qm, uint8_t Nl, uint8_t r, uint32_t *E_out)
{
  uint32_t Ncb, ind, E, k = 0;
  if (Kmimo == 0 || Mdlharq == 0 || C == 0 || Qm == 0 || Nl == 0) return -1;
  rm_params(RTC, C, Nsoft, Mdlharq, Kmimo, rvidx, &Ncb, &ind);
  E = e_for_block(G, C, Qm, Nl, r);
  if (clear == 1) memset(w, 0, Ncb * sizeof(int16_t));
  for (; ind < Ncb && k < E; ind++)
    if (dummy_w[ind] != ORC_LTE_NULL) w[ind] = (int16_t)(w[ind] + soft_input[k++]);
  while (k < E)
    for (ind = 0; ind < Ncb && k < E; ind++)
      if (dummy_w[ind] != ORC_LTE_NULL) w[ind] = (int16_t)(w[ind] + soft_input[k++]);
  *E_out = E;
  return 0;
}

/* ---- TX mirror (test-vector generation) ---------------------------------------- */

/* d: 96 leading pad bytes are NOT assumed; d holds 3*D values (bit or LTE_NULL).
 * w gets 3*Kpi values.  Reference :51-130: the three streams are written row-major
 * into a 32-column matrix after ND leading NULLs, stream 2 with the +1 shift
 * (pi(k) = (P[k/R] + 32*(k mod R) + 1) mod Kpi). */
uint32_t orc_sub_block_interleaving_turbo(uint32_t D, const uint8_t *d, uint8_t *w)
{
  uint32_t RTC = rtc_of(D), Kpi = RTC << 5, ND = Kpi - D, col, row, k = 0;
  for (col = 0; col < 32; col++) {
    uint32_t idx = brev5(col);
    for (row = 0; row < RTC; row++, k++, idx += 32) {
      /* position idx of the padded stream; entries < ND are dummies */
      w[k]           = (idx >= ND) ? d[3 * (idx - ND)] : ORC_LTE_NULL;
      w[Kpi + 2 * k] = (idx >= ND) ? d[3 * (idx - ND) + 1] : ORC_LTE_NULL;
      {
        uint32_t idx2 = (idx + 1) % Kpi;
        w[Kpi + 2 * k + 1] = (idx2 >= ND) ? d[3 * (idx2 - ND) + 2] : ORC_LTE_NULL;
      }
    }
  }
  return RTC;
}

uint32_t orc_lte_rate_matching_turbo(uint32_t RTC, uint32_t G, const uint8_t *w, uint8_t *e, uint8_t C,
                                     uint32_t Nsoft, uint8_t Mdlharq, uint8_t Kmimo, uint8_t rvidx,
                                     uint8_t Qm, uint8_t Nl, uint8_t r)
{
  uint32_t Ncb, ind, E, k = 0;
  rm_params(RTC, C, Nsoft, Mdlharq, Kmimo, rvidx, &Ncb, &ind);
  if (Ncb < 3 * (RTC << 5)) return 0;      /* the TX side gives up on a limited soft buffer (RM:508-511) */
  E = e_for_block(G, C, Qm, Nl, r);
  while (k < E) {
    if (ind >= Ncb) ind = 0;
    if (w[ind] != ORC_LTE_NULL) e[k++] = w[ind];
    ind++;
  }
  return E;
}
