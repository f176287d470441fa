/* TEST INFRASTRUCTURE (see oracle_port.h).
 * Scalar restatement of the reference 8-bit max-log-MAP turbo decoder
 * (reference: openair1/PHY/CODING/3gpplte_turbo_decoder_sse_8bit.c):
 *   gamma :151-209, alpha :213-409, beta :412-680, ext :682-827, tables :846-892,
 *   driver :894-1657 (input scaling :1000-1029, demux :1062-1077, exchange/hard
 *   decision :1341-1581, CRC :1583-1627, feedback :1632-1653).
 * Parity domain: n >= 256 and n % 16 == 0: outside it the reference
 * itself overruns its stack buffers.  16 SIMD lanes of int8, lane l covering trellis
 * positions [l*W,(l+1)*W), W = n/16; element (step k, lane l) sits at k*16+l.
 * Peculiarities restated as they are:
 *   - the 12 tail LLRs never influence the output (beta termination is disabled, the
 *     last lane starts from all-zero metrics, :519-532);
 *   - the alpha and beta re-runs cover 16 steps and the boundary metrics are re-seeded
 *     AFTER each of the two passes (:299-316, :652-666);
 *   - input scaling uses only the first 3*(n>>4)+1 vectors of 8 for its average and
 *     weights elements 4 and 5 of each vector twice (:1001-1008), |-32768| stays negative
 *     (_mm_abs_epi16), and the top bracket alternates shifts 3 and 4 per group of 8 (:1027-1029);
 *   - two hard-decision rules: n % 128 == 0 -> sign(ext2), else sign(ext2 (+) sys2) (:1392-1581). */
#include <stdlib.h>
#include <string.h>
#include "oracle_port.h"

typedef int8_t llr8_t;
#define INIT8 (-63)      /* -MAX8/2, MAX8 = 127 (:92,234) */
#define RERUN8 16        /* L (:211) */

static inline llr8_t sat8(int v) { return (llr8_t)(v > 127 ? 127 : (v < -128 ? -128 : v)); }
static inline llr8_t adds8(llr8_t a, llr8_t b) { return sat8((int)a + b); }
static inline llr8_t subs8(llr8_t a, llr8_t b) { return sat8((int)a - b); }
static inline llr8_t max8(llr8_t a, llr8_t b) { return a > b ? a : b; }

#define A8(k, s, l) alpha[(((k) * 8 + (s)) * 16) + (l)]
#define B8(k, s, l) beta[(((k) * 8 + (s)) * 16) + (l)]

static void gamma8(llr8_t *m11, llr8_t *m10, const llr8_t *sys, const llr8_t *par, int n)
{
  int j;                                   /* widen, add, >>1, pack: exact floor halves (:178-185) */
  for (j = 0; j < n; j++) {
    m11[j] = sat8(((int)sys[j] + par[j]) >> 1);
    m10[j] = sat8(((int)sys[j] - par[j]) >> 1);
  }
}

static void alpha8_steps(llr8_t *alpha, const llr8_t *m11, const llr8_t *m10, int steps)
{
  int k, l, s;
  for (k = 0; k < steps; k++)
    for (l = 0; l < 16; l++) {
      llr8_t g1 = m11[k * 16 + l], g0 = m10[k * 16 + l], a[8], nw[8], mx;
      for (s = 0; s < 8; s++) a[s] = A8(k, s, l);
      nw[0] = max8(adds8(a[1], g1), subs8(a[0], g1));      /* :251-284 */
      nw[1] = max8(subs8(a[3], g0), adds8(a[2], g0));
      nw[2] = max8(adds8(a[5], g0), subs8(a[4], g0));
      nw[3] = max8(subs8(a[7], g1), adds8(a[6], g1));
      nw[4] = max8(subs8(a[1], g1), adds8(a[0], g1));
      nw[5] = max8(adds8(a[3], g0), subs8(a[2], g0));
      nw[6] = max8(subs8(a[5], g0), adds8(a[4], g0));
      nw[7] = max8(adds8(a[7], g1), subs8(a[6], g1));
      mx = nw[0];
      for (s = 1; s < 8; s++) mx = max8(mx, nw[s]);
      for (s = 0; s < 8; s++) A8(k + 1, s, l) = subs8(nw[s], mx);
    }
}

static void alpha8_reseed(llr8_t *alpha, int W)
{
  int s, l;                                /* :299-316: byte shift of the 16-lane vector */
  for (s = 0; s < 8; s++) {
    for (l = 15; l >= 1; l--) A8(0, s, l) = A8(W, s, l - 1);
    A8(0, s, 0) = (s == 0) ? 0 : INIT8;
  }
}

static void beta8_steps(llr8_t *beta, const llr8_t *m11, const llr8_t *m10, int from, int to)
{
  int k, l, s;
  for (k = from; k >= to; k--)
    for (l = 0; l < 16; l++) {
      llr8_t g1 = m11[k * 16 + l], g0 = m10[k * 16 + l], b[8], nw[8], mx;
      for (s = 0; s < 8; s++) b[s] = B8(k + 1, s, l);
      nw[0] = max8(adds8(b[4], g1), subs8(b[0], g1));      /* :579-612 */
      nw[1] = max8(subs8(b[4], g1), adds8(b[0], g1));
      nw[2] = max8(subs8(b[5], g0), adds8(b[1], g0));
      nw[3] = max8(adds8(b[5], g0), subs8(b[1], g0));
      nw[4] = max8(adds8(b[6], g0), subs8(b[2], g0));
      nw[5] = max8(subs8(b[6], g0), adds8(b[2], g0));
      nw[6] = max8(subs8(b[7], g1), adds8(b[3], g1));
      nw[7] = max8(adds8(b[7], g1), subs8(b[3], g1));
      mx = nw[0];
      for (s = 1; s < 8; s++) mx = max8(mx, nw[s]);
      for (s = 0; s < 8; s++) B8(k, s, l) = subs8(nw[s], mx);
    }
}

static void log_map8(const llr8_t *sys, const llr8_t *par, llr8_t *ext, int n,
                     llr8_t *alpha, llr8_t *beta, llr8_t *m11, llr8_t *m10)
{
  int W = n >> 4, k, l, s, pass;
  gamma8(m11, m10, sys, par, n);
  /* alpha (:234-316): init, W steps, re-seed, 16 steps, re-seed */
  for (s = 0; s < 8; s++)
    for (l = 0; l < 16; l++) A8(0, s, l) = INIT8;
  A8(0, 0, 0) = 0;
  alpha8_steps(alpha, m11, m10, W);
  alpha8_reseed(alpha, W);
  alpha8_steps(alpha, m11, m10, RERUN8);
  alpha8_reseed(alpha, W);
  /* beta (:505-666): start from alpha[W]; before each pass lane 15 <- 0; after each pass
   * lane l <- beta[0] of lane l+1 (lane 15 <- 0 from the byte shift) */
  for (s = 0; s < 8; s++)
    for (l = 0; l < 16; l++) B8(W, s, l) = A8(W, s, l);
  for (pass = 0; pass < 2; pass++) {
    for (s = 0; s < 8; s++) B8(W, s, 15) = 0;
    beta8_steps(beta, m11, m10, W - 1, pass == 0 ? 0 : W - RERUN8);
    for (s = 0; s < 8; s++) {
      for (l = 0; l < 15; l++) B8(W, s, l) = B8(0, s, l + 1);
      B8(W, s, 15) = 0;
    }
  }
  /* ext (:715-770) */
  for (k = 0; k < W; k++)
    for (l = 0; l < 16; l++) {
      llr8_t a[8], b[8], g1 = m11[k * 16 + l], g0 = m10[k * 16 + l], m00, m01, mm10, mm11;
      for (s = 0; s < 8; s++) { a[s] = A8(k, s, l); b[s] = B8(k + 1, s, l); }
      m00  = max8(max8(adds8(a[0], b[0]), adds8(a[1], b[4])), max8(adds8(a[6], b[7]), adds8(a[7], b[3])));
      mm11 = max8(max8(adds8(a[0], b[4]), adds8(a[1], b[0])), max8(adds8(a[6], b[3]), adds8(a[7], b[7])));
      m01  = max8(max8(adds8(a[2], b[5]), adds8(a[3], b[1])), max8(adds8(a[4], b[2]), adds8(a[5], b[6])));
      mm10 = max8(max8(adds8(a[2], b[1]), adds8(a[3], b[5])), max8(adds8(a[4], b[6]), adds8(a[5], b[2])));
      m01 = subs8(m01, g0); m00 = subs8(m00, g1); mm10 = adds8(mm10, g0); mm11 = adds8(mm11, g1);
      ext[k * 16 + l] = subs8(max8(mm10, mm11), max8(m01, m00));
    }
}

static int abs16_sse(int v) { return v == -32768 ? -32768 : (v < 0 ? -v : v); }   /* _mm_abs_epi16 */

uint8_t orc_turbo_decoder8(const int16_t *y, uint8_t *decoded_bytes, uint16_t n,
                           uint8_t max_iterations, uint8_t crc_type, uint8_t F)
{
  int W, i, v, crc_len, round_avg, nb = n >> 3;
  int32_t lane_sum[4] = {0, 0, 0, 0};
  uint16_t *pi;
  llr8_t *y8, *s0, *s1, *s2, *yp1, *yp2, *ext, *ext2, *tmp, *alpha, *beta, *m11, *m10;
  uint8_t it = 0, ret = 0;
  int done = 0;

  if (crc_type > 3) return 255;
  if (orc_qpp_index(n) < 0) return 255;
  if (n < 256 || (n & 15)) return 254;                   /* outside the parity domain */
  crc_len = (crc_type == ORC_CRC16) ? 2 : (crc_type == ORC_CRC8) ? 1 : 3;
  W = n >> 4;
#define ST8(pos) ((((pos) % W) << 4) + ((pos) / W))       /* :874-881 */

  /* ---- input scaling (:1000-1029) ---- */
  for (v = 0; v < 3 * (n >> 4) + 1; v++) {
    const int16_t *p = y + 8 * v;
    lane_sum[0] += abs16_sse(p[0]) + abs16_sse(p[4]);
    lane_sum[1] += abs16_sse(p[1]) + abs16_sse(p[4]);
    lane_sum[2] += abs16_sse(p[2]) + abs16_sse(p[5]);
    lane_sum[3] += abs16_sse(p[3]) + abs16_sse(p[5]);
  }
  round_avg = (int)((int32_t)((uint32_t)lane_sum[0] + (uint32_t)lane_sum[1] + (uint32_t)lane_sum[2] + (uint32_t)lane_sum[3]) / (n * 3));
  y8 = (llr8_t *)calloc((size_t)3 * (n + 16), 1);
  for (i = 0; i < 3 * n; i++) {
    int sh;
    if (round_avg < 16) sh = 0;
    else if (round_avg < 32) sh = 1;
    else if (round_avg < 64) sh = 2;
    else if (round_avg < 128) sh = 3;
    else sh = ((i >> 3) & 1) ? 4 : 3;                     /* even vectors >>3, odd vectors >>4 */
    y8[i] = sat8(y[i] >> sh);
  }

  pi    = (uint16_t *)malloc(sizeof(uint16_t) * n);
  s0    = (llr8_t *)calloc((size_t)n + 64, 1);
  s1    = (llr8_t *)calloc((size_t)n + 64, 1);
  s2    = (llr8_t *)calloc((size_t)n + 64, 1);
  yp1   = (llr8_t *)calloc((size_t)n + 64, 1);
  yp2   = (llr8_t *)calloc((size_t)n + 64, 1);
  ext   = (llr8_t *)calloc((size_t)n + 128, 1);
  ext2  = (llr8_t *)calloc((size_t)n + 128, 1);
  tmp   = (llr8_t *)calloc((size_t)n + 128, 1);
  m11   = (llr8_t *)calloc((size_t)n + 64, 1);
  m10   = (llr8_t *)calloc((size_t)n + 64, 1);
  alpha = (llr8_t *)calloc((size_t)8 * (n + 64), 1);
  beta  = (llr8_t *)calloc((size_t)8 * (n + 64), 1);
  orc_qpp_table(n, pi);

  /* ---- demux into lane layout (:1062-1077); the tail LLRs are never used ---- */
  for (i = 0; i < n; i++) {
    s0[ST8(i)]  = y8[3 * i];
    yp1[ST8(i)] = y8[3 * i + 1];
    yp2[ST8(i)] = y8[3 * i + 2];
  }

  log_map8(s0, yp1, ext, n, alpha, beta, m11, m10);
  while (it++ < max_iterations) {
    for (i = 0; i < n; i++) s2[ST8(i)] = ext[ST8(pi[i])];                      /* :1341-1379 */
    log_map8(s2, yp2, ext2, n, alpha, beta, m11, m10);
    if ((n & 0x7f) != 0)                                                        /* :1456 */
      for (i = 0; i < n; i++) tmp[i] = adds8(ext2[i], s2[i]);
    for (i = 0; i < n; i++) {                                                   /* :1392-1459 */
      int j = ST8(pi[i]);
      s1[j] = adds8(subs8(ext2[ST8(i)], ext[j]), s0[j]);
    }
    if (it > 1) {
      uint32_t crc = 0, oldcrc = 0;
      memset(decoded_bytes, 0, (size_t)nb);
      for (i = 0; i < n; i++) {
        llr8_t d = ((n & 0x7f) == 0) ? ext2[ST8(i)] : tmp[ST8(i)];             /* :1488-1581 */
        if (d > 0) decoded_bytes[pi[i] >> 3] |= (uint8_t)(0x80 >> (pi[i] & 7));
      }
      for (i = 0; i < crc_len; i++) oldcrc |= (uint32_t)decoded_bytes[nb - crc_len + i] << (8 * i);
      switch (crc_type) {
      case ORC_CRC24_A:
        crc = orc_crc24a(&decoded_bytes[F >> 3], n - 24 - F) >> 8;
        crc = ((crc & 0xff) << 16) | (crc & 0xff00) | ((crc >> 16) & 0xff);
        break;
      case ORC_CRC24_B:
        crc = orc_crc24b(decoded_bytes, n - 24) >> 8;
        crc = ((crc & 0xff) << 16) | (crc & 0xff00) | ((crc >> 16) & 0xff);
        break;
      case ORC_CRC16:
        crc = orc_crc16(decoded_bytes, n - 16) >> 16;
        break;
      default:
        crc = orc_crc8(decoded_bytes, n - 8) >> 24;
        break;
      }
      if (crc == oldcrc && crc != 0) { ret = it; done = 1; break; }
    }
    if (it < max_iterations) {                                                  /* :1632-1653 */
      log_map8(s1, yp1, ext, n, alpha, beta, m11, m10);
      for (i = 0; i < n; i++) ext[i] = adds8(subs8(ext[i], s1[i]), s0[i]);
    }
  }
  if (!done) ret = it;
  free(y8); free(pi); free(s0); free(s1); free(s2); free(yp1); free(yp2); free(ext); free(ext2);
  free(tmp); free(m11); free(m10); free(alpha); free(beta);
  return ret;
}
