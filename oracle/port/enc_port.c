/* TEST INFRASTRUCTURE (see oracle_port.h).
 * 36.212 5.1.3.2 turbo encoder, scalar, with the reference's output order
 * (reference: openair1/PHY/CODING/3gpplte_sse.c:96-109 RSC + termination,
 * :380-476 encoder): out[3i..3i+2] = (c_i, z_i, z'_i), then 12 tail bits
 * x0 z0 x1 z1 x2 z2 of encoder 1 followed by the same for encoder 2.
 * One output byte per coded bit (0/1).  Used only to make test vectors. */
#include <stdlib.h>
#include "oracle_port.h"

/* registers D1 D2 D3 held as state bits 2,1,0; g0 = 1+D^2+D^3 feedback, g1 = 1+D+D^3 */
static uint8_t rsc(uint8_t in, uint8_t *st)
{
  uint8_t d1 = (*st >> 2) & 1, d2 = (*st >> 1) & 1, d3 = *st & 1;
  uint8_t a = in ^ d2 ^ d3;
  *st = (uint8_t)((a << 2) | (d1 << 1) | d2);
  return a ^ d1 ^ d3;
}

static void rsc_term(uint8_t *x, uint8_t *z, uint8_t *st)
{
  uint8_t d1 = (*st >> 2) & 1, d2 = (*st >> 1) & 1, d3 = *st & 1;
  *x = d2 ^ d3;          /* makes the feedback sum zero */
  *z = d1 ^ d3;
  *st >>= 1;
}

void orc_turbo_encode(const uint8_t *input, int nbytes, uint8_t *out)
{
  int K = nbytes * 8, i, t;
  uint16_t *pi = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)K);
  uint8_t s1 = 0, s2 = 0;
  if (orc_qpp_table(K, pi) != 0) { free(pi); return; }
  for (i = 0; i < K; i++) {
    uint8_t c  = (input[i >> 3] >> (7 - (i & 7))) & 1;
    uint8_t ci = (input[pi[i] >> 3] >> (7 - (pi[i] & 7))) & 1;
    out[3 * i]     = c;
    out[3 * i + 1] = rsc(c, &s1);
    out[3 * i + 2] = rsc(ci, &s2);
  }
  for (t = 0; t < 3; t++) rsc_term(&out[3 * K + 2 * t], &out[3 * K + 2 * t + 1], &s1);
  for (t = 0; t < 3; t++) rsc_term(&out[3 * K + 6 + 2 * t], &out[3 * K + 7 + 2 * t], &s2);
  free(pi);
}
