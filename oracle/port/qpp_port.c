/* TEST INFRASTRUCTURE (see oracle_port.h).
 * QPP interleaver parameters: 3GPP TS 36.212 table 5.1.3-3.  The reference keeps the
 * same 188 (f1,f2) pairs in f1f2mat_old[] (openair1/PHY/CODING/lte_interleaver2.h:29-217)
 * and indexes them with the bucket rule of lte_interleaver_inline.h:54-70 /
 * dlsch_decoding.c:314-325.  The numbers below are the standard's table; the test
 * tests/test_oracle_pin.py checks them against the reference header. */
#include "oracle_port.h"

static const uint16_t qpp_f1f2[188][2] = {
#include "qpp_table.inc"
};

int orc_qpp_K(int idx)
{
  if (idx < 0 || idx >= 188) return -1;
  if (idx < 60)  return 40 + 8 * idx;
  if (idx < 92)  return 512 + 16 * (idx - 59);
  if (idx < 124) return 1024 + 32 * (idx - 91);
  return 2048 + 64 * (idx - 123);
}

int orc_qpp_index(int K)
{
  int kb;
  if (K < 40 || K > 6144 || (K & 7)) return -1;
  kb = K >> 3;
  if (kb <= 64) return kb - 5;
  if (kb <= 128) return (K & 15) ? -1 : 59 + ((kb - 64) >> 1);
  if (kb <= 256) return (K & 31) ? -1 : 91 + ((kb - 128) >> 2);
  return (K & 63) ? -1 : 123 + ((kb - 256) >> 3);
}

int orc_qpp_f1(int idx) { return (idx < 0 || idx >= 188) ? -1 : qpp_f1f2[idx][0]; }
int orc_qpp_f2(int idx) { return (idx < 0 || idx >= 188) ? -1 : qpp_f1f2[idx][1]; }

int orc_qpp_table(int K, uint16_t *pi)
{
  int idx = orc_qpp_index(K), i;
  uint64_t f1, f2;
  if (idx < 0) return -1;
  f1 = qpp_f1f2[idx][0];
  f2 = qpp_f1f2[idx][1];
  for (i = 0; i < K; i++)
    pi[i] = (uint16_t)((f1 * (uint64_t)i + f2 * (uint64_t)i * (uint64_t)i) % (uint64_t)K);
  return 0;
}
