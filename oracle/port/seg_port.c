/* TEST INFRASTRUCTURE (see oracle_port.h).
 * Code-block segmentation parameters, following the reference rule including its
 * quirks (reference: openair1/PHY/CODING/lte_segmentation.c:52-134):
 *   - K- for B'/C <= 512 is B'/C - 8, not K+ - 8 (:81);
 *   - only C > 16 and B'/C > 6144 are errors. */
#include "oracle_port.h"

int orc_lte_segmentation(uint32_t B, uint32_t *C, uint32_t *Cplus, uint32_t *Cminus,
                         uint32_t *Kplus, uint32_t *Kminus, uint32_t *F)
{
  uint32_t L, Bp, q;
  if (B <= 6144) {
    L = 0; *C = 1; Bp = B;
  } else {
    L = 24;
    *C = B / (6144 - L);
    if ((6144 - L) * (*C) < B) (*C)++;
    Bp = B + (*C) * L;
  }
  if (*C > 16) return -1;
  q = Bp / (*C);
  if (q <= 40) { *Kplus = 40; *Kminus = 0; }
  else if (q <= 512)  { *Kplus = (q >> 3) << 3; *Kminus = q - 8; }
  else if (q <= 1024) { *Kplus = (q >> 4) << 4; if (*Kplus < q) *Kplus += 16; *Kminus = *Kplus - 16; }
  else if (q <= 2048) { *Kplus = (q >> 5) << 5; if (*Kplus < q) *Kplus += 32; *Kminus = *Kplus - 32; }
  else if (q <= 6144) { *Kplus = (q >> 6) << 6; if (*Kplus < q) *Kplus += 64; *Kminus = *Kplus - 64; }
  else return -1;
  if (*C == 1) { *Cplus = 1; *Kminus = 0; *Cminus = 0; }
  else {
    *Cminus = ((*C) * (*Kplus) - Bp) / (*Kplus - *Kminus);
    *Cplus = *C - *Cminus;
  }
  *F = (*Cplus) * (*Kplus) + (*Cminus) * (*Kminus) - Bp;
  return 0;
}
