/* TEST INFRASTRUCTURE (see oracle_port.h).
 * Pseudo-random (Gold) sequence of 36.211 7.2 as the reference generates it, 32 bits per
 * call (reference: openair1/PHY/LTE_REFSIG/lte_gold.c:151-180), and the downlink
 * descrambling of soft bits (reference: openair1/PHY/LTE_TRANSPORT/dlsch_scrambling.c:99-138).
 * Restated as they are:
 *   - x1 starts from the fixed word 1 + 2^31, x2 from c_init completed with its 32nd bit; the
 *     first 1600 outputs are skipped by 49 + 1 word steps;
 *   - the sign applied to soft bit k is 2*c(k) - 1, i.e. the LLR is NEGATED where the
 *     scrambling bit is 0; the product is stored back as int16 (so -32768 stays -32768);
 *   - the reference loop covers 32*(1 + G/32) soft bits, i.e. up to 31 beyond G. */
#include "oracle_port.h"

static void gold_step(uint32_t *x1, uint32_t *x2)
{
  *x1 = (*x1 >> 1) ^ (*x1 >> 4);
  *x1 = *x1 ^ (*x1 << 31) ^ (*x1 << 28);
  *x2 = (*x2 >> 1) ^ (*x2 >> 2) ^ (*x2 >> 3) ^ (*x2 >> 4);
  *x2 = *x2 ^ (*x2 << 31) ^ (*x2 << 30) ^ (*x2 << 29) ^ (*x2 << 28);
}

uint32_t orc_lte_gold_generic(uint32_t *x1, uint32_t *x2, uint8_t reset)
{
  if (reset) {
    int n;
    *x1 = 1u + (1u << 31);
    *x2 = *x2 ^ ((*x2 ^ (*x2 >> 1) ^ (*x2 >> 2) ^ (*x2 >> 3)) << 31);
    for (n = 1; n < 50; n++) gold_step(x1, x2);
  }
  gold_step(x1, x2);
  return *x1 ^ *x2;
}

/* words[i] = scrambling bits 32i .. 32i+31 (bit j of the word = c(32i + j)) */
void orc_gold_words(uint32_t c_init, uint32_t *words, int nwords)
{
  uint32_t x1 = 0, x2 = c_init;
  int i;
  for (i = 0; i < nwords; i++) words[i] = orc_lte_gold_generic(&x1, &x2, i == 0);
}

/* dlsch_unscrambling over the first n soft bits (n = 32*(1 + G/32) reproduces the reference loop) */
void orc_dlsch_unscrambling(uint32_t c_init, int16_t *llr, int n)
{
  uint32_t x1 = 0, x2 = c_init, s = 0;
  int k;
  for (k = 0; k < n; k++) {
    if ((k & 31) == 0) s = orc_lte_gold_generic(&x1, &x2, k == 0);
    llr[k] = (int16_t)((2 * (int)((s >> (k & 31)) & 1) - 1) * (int)llr[k]);
  }
}
