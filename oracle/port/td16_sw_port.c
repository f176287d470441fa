/* TEST INFRASTRUCTURE -- CPU model of the OPTIONAL sliding-window decoding mode (OAI_BATCH_SLIDING_WINDOW).
 *
 * PARITY UNPINNED (by construction, for this file only; every other file under oracle/port is pinned on the compiled
 * reference): this file does NOT restate a reference function: the sliding-window mode is this repo's own higher-parallelism
 * variant of the 16-bit max-log-MAP decoder (north_star: "allowed only if it is separately reported with its BLER
 * delta against the bit-exact mode"), so there is nothing in /root/reference to pin it on.  What it shares with the
 * reference (3gpplte_turbo_decoder_sse_16bit.c) is the trellis (:292-322, :592-636), the branch metrics (:121-169,
 * floor halves of s+p and s-p), the LLR (:757-818), the tail-bit start metrics (:474-520), the extrinsic exchange
 * (:1209-1265, :1354-1375), the hard decision / CRC / early-exit rule (:1267-1351) and the return value (:985,:1348).
 * What differs, by design:
 *   - the K trellis positions are split into NW = 8 / 16 / 32 / 64 windows (K < 512 / < 1024 / < 2048 / >= 2048) instead
 *     of 8 lanes.  A window's forward (backward) recursion starts from the MORE CONFIDENT (larger max - min) of two
 *     candidates: the metrics its left (right) neighbour ended with in the PREVIOUS iteration ("next-iteration
 *     initialisation") and the result of a training recursion over the neighbour's last (first) min(32, WL) steps, started
 *     from equal metrics on the CURRENT inputs -- instead of the reference's 5-step re-run.  Window 0 always starts in
 *     state 0, the last window's backward recursion from the tail bits;
 *   - the soft bits are scaled to 8 bits first (right shift chosen from the block's mean |y|, then clipped to +-127)
 *     and the extrinsic values are clipped to +-767: with these bounds the int16 recursions cannot overflow, so the
 *     arithmetic is plain (non-saturating) integer arithmetic -- computed here in int, on the GPU in int16x2.
 * The GPU kernel (openair4g_b200/csrc/td16_sw.cuh) must reproduce this model bit for bit (tests/test_gpu_sw.py);
 * its BLER against the bit-exact mode is measured by tools/sw_bler_delta.py.
 */
#include <stdlib.h>
#include <string.h>
#include "oracle_port.h"

#define SW_LC 767
#define SW_Q  3000
#define SW_TRAIN 32

int orc_sw_windows(int K) { return K >= 2048 ? 64 : (K >= 1024 ? 32 : (K >= 512 ? 16 : 8)); }

int orc_sw_shift(const int16_t *y, int n)
{
  long sum = 0;
  int i, sh = 0;
  for (i = 0; i < n; i++) sum += (y[i] < 0) ? -(long)y[i] : (long)y[i];
  {
    long mean = sum / n;
    while ((mean >> sh) > 24) sh++;
  }
  return sh;
}

static int sw_scale(int v, int sh)
{
  v >>= sh;                                   /* arithmetic shift (floor) */
  return v > 127 ? 127 : (v < -127 ? -127 : v);
}

static int imax(int a, int b) { return a > b ? a : b; }

static void alpha_step(int *a, int g1, int g0)
{
  int nw[8], s;
  nw[0] = imax(a[1] + g1, a[0] - g1);
  nw[1] = imax(a[3] - g0, a[2] + g0);
  nw[2] = imax(a[5] + g0, a[4] - g0);
  nw[3] = imax(a[7] - g1, a[6] + g1);
  nw[4] = imax(a[1] - g1, a[0] + g1);
  nw[5] = imax(a[3] + g0, a[2] - g0);
  nw[6] = imax(a[5] - g0, a[4] + g0);
  nw[7] = imax(a[7] + g1, a[6] - g1);
  for (s = 0; s < 8; s++) a[s] = nw[s] - nw[0];      /* state 0 = 0: any uniform shift leaves the LLRs unchanged */
}

static void beta_step(int *b, int g1, int g0)
{
  int nw[8], s;
  nw[0] = imax(b[4] + g1, b[0] - g1);
  nw[1] = imax(b[4] - g1, b[0] + g1);
  nw[2] = imax(b[5] - g0, b[1] + g0);
  nw[3] = imax(b[5] + g0, b[1] - g0);
  nw[4] = imax(b[6] + g0, b[2] - g0);
  nw[5] = imax(b[6] - g0, b[2] + g0);
  nw[6] = imax(b[7] - g1, b[3] + g1);
  nw[7] = imax(b[7] + g1, b[3] - g1);
  for (s = 0; s < 8; s++) b[s] = nw[s] - nw[0];
}

static int spread(const int *v)
{
  int mx = v[0], mn = v[0], s;
  for (s = 1; s < 8; s++) { if (v[s] > mx) mx = v[s]; if (v[s] < mn) mn = v[s]; }
  return mx - mn;
}

static int llr_step(const int *a, const int *b, int g1, int g0)
{
  int m00 = imax(imax(a[0] + b[0], a[1] + b[4]), imax(a[6] + b[7], a[7] + b[3]));
  int m11 = imax(imax(a[0] + b[4], a[1] + b[0]), imax(a[6] + b[3], a[7] + b[7]));
  int m01 = imax(imax(a[2] + b[5], a[3] + b[1]), imax(a[4] + b[2], a[5] + b[6]));
  int m10 = imax(imax(a[2] + b[1], a[3] + b[5]), imax(a[4] + b[6], a[5] + b[2]));
  return imax(m10 + g0, m11 + g1) - imax(m01 - g0, m00 - g1);
}

/* start metrics of the last window's backward recursion from the three tail steps (same paths as :474-520), state 0 = 0 */
static void tail_beta(const int *ts, const int *tp, int *t)
{
  int c11, c10, b0, b1, b0_2, b1_2, b2_2, b3_2, s, t0;
  c11 = (ts[2] + tp[2]) >> 1;
  b0 = -c11; b1 = c11;
  c11 = (ts[1] + tp[1]) >> 1; c10 = (ts[1] - tp[1]) >> 1;
  b0_2 = b0 - c11; b1_2 = b0 + c11; b2_2 = b1 + c10; b3_2 = b1 - c10;
  c11 = (ts[0] + tp[0]) >> 1; c10 = (ts[0] - tp[0]) >> 1;
  t[0] = b0_2 - c11; t[1] = b0_2 + c11; t[2] = b1_2 + c10; t[3] = b1_2 - c10;
  t[4] = b2_2 - c10; t[5] = b2_2 + c10; t[6] = b3_2 + c11; t[7] = b3_2 - c11;
  t0 = t[0];
  for (s = 0; s < 8; s++) t[s] -= t0;
}

/* one constituent pass over all windows.  in/par/out are indexed by trellis position of THIS decoder (natural order
 * for decoder 1, interleaved order for decoder 2).  llr receives the a-posteriori LLR. */
static void sw_pass(const int *in, const int *par, int *llr, int K, int NW, int *nii_a, int *nii_b, const int *term)
{
  int WL = K / NW, L = WL < SW_TRAIN ? WL : SW_TRAIN, w, o, s;
  int *alpha = (int *)malloc(sizeof(int) * 8 * (size_t)(WL + 1));
  int *na = (int *)malloc(sizeof(int) * 8 * (size_t)NW), *nb = (int *)malloc(sizeof(int) * 8 * (size_t)NW);
  memcpy(na, nii_a, sizeof(int) * 8 * (size_t)NW);
  memcpy(nb, nii_b, sizeof(int) * 8 * (size_t)NW);
  for (w = 0; w < NW; w++) {
    int a[8], b[8];
    const int *x = in + w * WL, *p = par + w * WL;
    memcpy(a, nii_a + 8 * w, sizeof(a));
    if (w > 0) {                                                     /* training over the left neighbour's last steps */
      int t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q;
      for (q = w * WL - L; q < w * WL; q++) alpha_step(t, (in[q] + par[q]) >> 1, (in[q] - par[q]) >> 1);
      if (spread(t) >= spread(a)) memcpy(a, t, sizeof(a));
    }
    for (o = 0; o < WL; o++) {
      memcpy(alpha + 8 * o, a, sizeof(a));
      alpha_step(a, (x[o] + p[o]) >> 1, (x[o] - p[o]) >> 1);
    }
    if (w + 1 < NW) memcpy(na + 8 * (w + 1), a, sizeof(a));          /* next iteration: start of the right neighbour */
    memcpy(b, nii_b + 8 * w, sizeof(b));
    if (w + 1 < NW) {                                                /* ... over the right neighbour's first steps */
      int t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q;
      for (q = (w + 1) * WL + L - 1; q >= (w + 1) * WL; q--) beta_step(t, (in[q] + par[q]) >> 1, (in[q] - par[q]) >> 1);
      if (spread(t) >= spread(b)) memcpy(b, t, sizeof(b));
    }
    for (o = WL - 1; o >= 0; o--) {
      int g1 = (x[o] + p[o]) >> 1, g0 = (x[o] - p[o]) >> 1;
      llr[w * WL + o] = llr_step(alpha + 8 * o, b, g1, g0);
      beta_step(b, g1, g0);
    }
    if (w > 0) memcpy(nb + 8 * (w - 1), b, sizeof(b));               /* next iteration: start of the left neighbour */
  }
  for (s = 0; s < 8; s++) { na[s] = s ? -SW_Q : 0; nb[8 * (NW - 1) + s] = term[s]; }
  memcpy(nii_a, na, sizeof(int) * 8 * (size_t)NW);
  memcpy(nii_b, nb, sizeof(int) * 8 * (size_t)NW);
  free(alpha); free(na); free(nb);
}

static int clampi(int v, int lim) { return v > lim ? lim : (v < -lim ? -lim : v); }

/* Same contract as orc_turbo_decoder16.  dbg_llr (optional, K ints): a-posteriori LLRs of the last second-decoder pass
 * in interleaved order (kernel debugging). */
uint8_t orc_turbo_decoder16_sw(const int16_t *y, uint8_t *decoded_bytes, uint16_t n, uint8_t max_iterations,
                               uint8_t crc_type, uint8_t F, int *dbg_llr)
{
  int K = n, NW, i, s, sh, crc_len, it = 0, done = 0;
  uint8_t ret = 0;
  uint16_t *pi;
  int *s0, *p1, *p2, *A, *B, *llr, *nii, term[2][8], ts[3], tp[3];
  if (crc_type > 3) return 255;
  if (orc_qpp_index(K) < 0) return 255;
  crc_len = (crc_type == ORC_CRC16) ? 2 : (crc_type == ORC_CRC8) ? 1 : 3;
  NW = orc_sw_windows(K);
  pi = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)K);
  orc_qpp_table(K, pi);
  s0 = (int *)malloc(sizeof(int) * 6 * (size_t)K);
  p1 = s0 + K; p2 = p1 + K; A = p2 + K; B = A + K; llr = B + K;
  nii = (int *)calloc((size_t)4 * 8 * NW, sizeof(int));          /* [decoder][alpha, beta][window][state] */
  sh = orc_sw_shift(y, 3 * K + 12);
  for (i = 0; i < K; i++) {
    s0[i] = sw_scale(y[3 * i], sh);
    p1[i] = sw_scale(y[3 * i + 1], sh);
    p2[i] = sw_scale(y[3 * i + 2], sh);
    A[i] = s0[i];
  }
  for (s = 0; s < 2; s++) {
    for (i = 0; i < 3; i++) { ts[i] = sw_scale(y[3 * K + 6 * s + 2 * i], sh); tp[i] = sw_scale(y[3 * K + 6 * s + 2 * i + 1], sh); }
    tail_beta(ts, tp, term[s]);
  }
  for (s = 0; s < 2; s++) {
    int *na = nii + (size_t)s * 16 * NW, *nb = na + 8 * NW;
    for (i = 1; i < 8; i++) na[i] = -SW_Q;                        /* window 0 starts in state 0 */
    memcpy(nb + 8 * (NW - 1), term[s], sizeof(term[s]));
  }
  while (it++ < max_iterations) {
    /* first decoder: A = systematic + a-priori in natural order -> B = s0 + extrinsic (feedback fused, :1354-1375) */
    sw_pass(A, p1, llr, K, NW, nii, nii + 8 * NW, term[0]);
    for (i = 0; i < K; i++) B[i] = clampi(llr[i] - A[i], SW_LC) + s0[i];
    for (i = 0; i < K; i++) A[i] = B[pi[i]];                       /* :1209-1231 */
    sw_pass(A, p2, llr, K, NW, nii + 16 * NW, nii + 24 * NW, term[1]);
    for (i = 0; i < K; i++) B[pi[i]] = s0[pi[i]] + clampi(llr[i] - A[i], SW_LC);       /* :1241-1265 */
    if (dbg_llr) memcpy(dbg_llr, llr, sizeof(int) * (size_t)K);
    if (it > 1) {                                                  /* :1267-1351 */
      uint32_t crc = 0, oldcrc = 0;
      int nb = K >> 3;
      memset(decoded_bytes, 0, (size_t)nb);
      for (i = 0; i < K; i++)
        if (llr[i] > 0) decoded_bytes[pi[i] >> 3] |= (uint8_t)(0x80 >> (pi[i] & 7));
      for (i = 0; i < crc_len; i++) oldcrc |= (uint32_t)decoded_bytes[nb - crc_len + i] << (8 * i);
      switch (crc_type) {
      case ORC_CRC24_A:
        crc = orc_crc24a(&decoded_bytes[F >> 3], K - 24 - F) >> 8;
        crc = ((crc & 0xff) << 16) | (crc & 0xff00) | ((crc >> 16) & 0xff);
        break;
      case ORC_CRC24_B:
        crc = orc_crc24b(decoded_bytes, K - 24) >> 8;
        crc = ((crc & 0xff) << 16) | (crc & 0xff00) | ((crc >> 16) & 0xff);
        break;
      case ORC_CRC16:
        crc = orc_crc16(decoded_bytes, K - 16) >> 16;
        break;
      default:
        crc = orc_crc8(decoded_bytes, K - 8) >> 24;
        break;
      }
      if (crc == oldcrc && crc != 0) { ret = (uint8_t)it; done = 1; break; }
    }
    memcpy(A, B, sizeof(int) * (size_t)K);
  }
  if (!done) ret = (uint8_t)it;
  free(pi); free(s0); free(nii);
  return ret;
}
