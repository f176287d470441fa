/* TEST INFRASTRUCTURE (see oracle_port.h).
 * Scalar restatement of the reference 16-bit max-log-MAP turbo decoder
 * (reference: openair1/PHY/CODING/3gpplte_turbo_decoder_sse_16bit.c):
 *   gamma :121-169, alpha :173-439, beta :442-693, ext :695-879,
 *   lane tables :898-943, driver :945-1385.
 * The reference runs 8 SIMD lanes, lane l covering trellis positions
 * [l*W,(l+1)*W), W=n/8; element (step k, lane l) of a lane-layout array sits at
 * k*8+l.  alpha/beta are stored as [(k*8+state)*8+lane].  Everything below
 * mirrors the reference's order of saturating operations, its two-pass
 * boundary heuristic (full pass with fixed start metrics, then a 5-step re-run
 * seeded from the neighbouring lane) and its wrapping scalar tail initialisation. */
#include <stdlib.h>
#include <string.h>
#include "oracle_port.h"

typedef int16_t llr_t;
#define NEG_INIT (-128)   /* -MAX/2 with MAX=256, reference :79,201 */
#define RERUN 5           /* L>>3 with L=40, reference :171,189 */

static inline llr_t sat16(int32_t v) { return (llr_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }
static inline llr_t adds(llr_t a, llr_t b) { return sat16((int32_t)a + b); }
static inline llr_t subs(llr_t a, llr_t b) { return sat16((int32_t)a - b); }
static inline llr_t max16(llr_t a, llr_t b) { return a > b ? a : b; }

/* gamma, reference :121-169.  Index W (the "termination" vector) reads systematic
 * vector W+term_flag; only its first 3 lanes are ever consumed. */
static void gamma16(llr_t *m11, llr_t *m10, const llr_t *sys, const llr_t *par, int n, int term)
{
  int j;
  for (j = 0; j < n; j++) {
    m11[j] = (llr_t)(adds(sys[j], par[j]) >> 1);
    m10[j] = (llr_t)(subs(sys[j], par[j]) >> 1);
  }
  for (j = 0; j < 8; j++) {
    m11[n + j] = (llr_t)(adds(sys[n + 8 * term + j], par[n + j]) >> 1);
    m10[n + j] = (llr_t)(subs(sys[n + 8 * term + j], par[n + j]) >> 1);
  }
}

#define A(k, s, l) alpha[(((k) * 8 + (s)) * 8) + (l)]
#define B(k, s, l) beta[(((k) * 8 + (s)) * 8) + (l)]

static void alpha_steps(llr_t *alpha, const llr_t *m11, const llr_t *m10, int steps)
{
  int k, l, s;
  for (k = 0; k < steps; k++)
    for (l = 0; l < 8; l++) {
      llr_t g1 = m11[k * 8 + l], g0 = m10[k * 8 + l], a[8], nw[8], mx;
      for (s = 0; s < 8; s++) a[s] = A(k, s, l);
      /* reference :292-322 */
      nw[0] = max16(adds(a[1], g1), subs(a[0], g1));
      nw[1] = max16(subs(a[3], g0), adds(a[2], g0));
      nw[2] = max16(adds(a[5], g0), subs(a[4], g0));
      nw[3] = max16(subs(a[7], g1), adds(a[6], g1));
      nw[4] = max16(subs(a[1], g1), adds(a[0], g1));
      nw[5] = max16(adds(a[3], g0), subs(a[2], g0));
      nw[6] = max16(subs(a[5], g0), adds(a[4], g0));
      nw[7] = max16(adds(a[7], g1), subs(a[6], g1));
      mx = nw[0];
      for (s = 1; s < 8; s++) mx = max16(mx, nw[s]);
      for (s = 0; s < 8; s++) A(k + 1, s, l) = subs(nw[s], mx);
    }
}

static void alpha16(llr_t *alpha, const llr_t *m11, const llr_t *m10, int n)
{
  int W = n >> 3, l, s;
  /* pass 1, reference :201-208 */
  for (s = 0; s < 8; s++)
    for (l = 0; l < 8; l++) A(0, s, l) = NEG_INIT;
  A(0, 0, 0) = 0;
  alpha_steps(alpha, m11, m10, W);
  /* pass 2, reference :232-259: lane l <- final metrics of lane l-1 (byte shift of
   * the 8-lane vector), lane 0 <- (0,-128,...) */
  for (s = 0; s < 8; s++) {
    for (l = 7; l >= 1; l--) A(0, s, l) = A(W, s, l - 1);
    A(0, s, 0) = (s == 0) ? 0 : NEG_INIT;
  }
  alpha_steps(alpha, m11, m10, RERUN);
}

static void beta_steps(llr_t *beta, const llr_t *m11, const llr_t *m10, int from, int to)
{
  int k, l, s;
  for (k = from; k >= to; k--)
    for (l = 0; l < 8; l++) {
      llr_t g1 = m11[k * 8 + l], g0 = m10[k * 8 + l], b[8], nw[8], mx;
      for (s = 0; s < 8; s++) b[s] = B(k + 1, s, l);
      /* reference :592-636 */
      nw[0] = max16(adds(b[4], g1), subs(b[0], g1));
      nw[1] = max16(subs(b[4], g1), adds(b[0], g1));
      nw[2] = max16(subs(b[5], g0), adds(b[1], g0));
      nw[3] = max16(adds(b[5], g0), subs(b[1], g0));
      nw[4] = max16(adds(b[6], g0), subs(b[2], g0));
      nw[5] = max16(subs(b[6], g0), adds(b[2], g0));
      nw[6] = max16(subs(b[7], g1), adds(b[3], g1));
      nw[7] = max16(adds(b[7], g1), subs(b[3], g1));
      mx = nw[0];
      for (s = 1; s < 8; s++) mx = max16(mx, nw[s]);
      for (s = 0; s < 8; s++) B(k, s, l) = subs(nw[s], mx);
    }
}

static void beta16(const llr_t *alpha, llr_t *beta, const llr_t *m11, const llr_t *m10, int n)
{
  int W = n >> 3, l, s, loopval;
  int16_t c11, c10, b0, b1, b0_2, b1_2, b2_2, b3_2, t[8], bm;
  /* tail-bit initialisation in plain (WRAPPING) int16, reference :474-520 */
  c11 = m11[n + 2];
  b0 = (int16_t)(-c11);
  b1 = c11;
  c11 = m11[n + 1]; c10 = m10[n + 1];
  b0_2 = (int16_t)(b0 - c11);
  b1_2 = (int16_t)(b0 + c11);
  b2_2 = (int16_t)(b1 + c10);
  b3_2 = (int16_t)(b1 - c10);
  c11 = m11[n]; c10 = m10[n];
  t[0] = (int16_t)(b0_2 - c11);
  t[1] = (int16_t)(b0_2 + c11);
  t[2] = (int16_t)(b1_2 + c10);
  t[3] = (int16_t)(b1_2 - c10);
  t[4] = (int16_t)(b2_2 - c10);
  t[5] = (int16_t)(b2_2 + c10);
  t[6] = (int16_t)(b3_2 + c11);
  t[7] = (int16_t)(b3_2 - c11);
  bm = t[0];
  for (s = 1; s < 8; s++) bm = (bm > t[s]) ? bm : t[s];
  for (s = 0; s < 8; s++) t[s] = (int16_t)(t[s] - bm);

  /* pass 1, reference :531-538,566-573: lanes 0..6 start from their own final alpha */
  for (s = 0; s < 8; s++) {
    for (l = 0; l < 7; l++) B(W, s, l) = alpha[((W * 8 + s) * 8) + l];
    B(W, s, 7) = t[s];
  }
  beta_steps(beta, m11, m10, W - 1, 0);
  /* pass 2, reference :541-549,585-587: lane l <- beta[0] of lane l+1, last 5 steps */
  for (s = 0; s < 8; s++) {
    for (l = 0; l < 7; l++) B(W, s, l) = B(0, s, l + 1);
    B(W, s, 7) = t[s];
  }
  loopval = (n - 40) >> 3;
  beta_steps(beta, m11, m10, W - 1, loopval);
}

/* reference :757-818 */
static void ext16(const llr_t *alpha, const llr_t *beta, const llr_t *m11, const llr_t *m10,
                  llr_t *ext, int n)
{
  int W = n >> 3, k, l;
  for (k = 0; k < W; k++)
    for (l = 0; l < 8; l++) {
      llr_t a[8], b[8], g1 = m11[k * 8 + l], g0 = m10[k * 8 + l];
      llr_t m00, m01, mm10, mm11;
      int s;
      for (s = 0; s < 8; s++) { a[s] = alpha[((k * 8 + s) * 8) + l]; b[s] = B(k + 1, s, l); }
      m00  = max16(max16(adds(a[0], b[0]), adds(a[1], b[4])), max16(adds(a[6], b[7]), adds(a[7], b[3])));
      mm11 = max16(max16(adds(a[0], b[4]), adds(a[1], b[0])), max16(adds(a[6], b[3]), adds(a[7], b[7])));
      m01  = max16(max16(adds(a[2], b[5]), adds(a[3], b[1])), max16(adds(a[4], b[2]), adds(a[5], b[6])));
      mm10 = max16(max16(adds(a[2], b[1]), adds(a[3], b[5])), max16(adds(a[4], b[6]), adds(a[5], b[2])));
      m01  = subs(m01, g0);
      m00  = subs(m00, g1);
      mm10 = adds(mm10, g0);
      mm11 = adds(mm11, g1);
      ext[k * 8 + l] = subs(max16(mm10, mm11), max16(m01, m00));
    }
}

void orc_log_map16(const int16_t *sys, const int16_t *par, int16_t *ext, int n, int term_flag,
                   int16_t *alpha_dump, int16_t *beta_dump)
{
  size_t ab = (size_t)8 * (n + 16);
  llr_t *alpha = (llr_t *)calloc(ab, sizeof(llr_t));
  llr_t *beta  = (llr_t *)calloc(ab, sizeof(llr_t));
  llr_t *m11   = (llr_t *)calloc((size_t)n + 16, sizeof(llr_t));
  llr_t *m10   = (llr_t *)calloc((size_t)n + 16, sizeof(llr_t));
  gamma16(m11, m10, sys, par, n, term_flag);
  alpha16(alpha, m11, m10, n);
  beta16(alpha, beta, m11, m10, n);
  ext16(alpha, beta, m11, m10, ext, n);
  if (alpha_dump) memcpy(alpha_dump, alpha, ab * sizeof(llr_t));
  if (beta_dump)  memcpy(beta_dump, beta, ab * sizeof(llr_t));
  free(alpha); free(beta); free(m11); free(m10);
}

uint8_t orc_turbo_decoder16(const int16_t *y, uint8_t *decoded_bytes, uint16_t n,
                            uint8_t max_iterations, uint8_t crc_type, uint8_t F)
{
  int W, i, p, crc_len;
  uint16_t *pi;
  llr_t *s0, *s1, *s2, *yp1, *yp2, *ext, *ext2;
  uint8_t it = 0, ret = 0;
  int done = 0;

  if (crc_type > 3) return 255;                 /* reference :1003-1006 */
  if (orc_qpp_index(n) < 0) return 255;         /* reference :1011-1018 */
  crc_len = (crc_type == ORC_CRC16) ? 2 : (crc_type == ORC_CRC8) ? 1 : 3;
  W = n >> 3;
#define ST(pos) ((((pos) % W) << 3) + ((pos) / W))   /* reference :921-932 */

  pi   = (uint16_t *)malloc(sizeof(uint16_t) * n);
  s0   = (llr_t *)calloc((size_t)n + 16, sizeof(llr_t));
  s1   = (llr_t *)calloc((size_t)n + 16, sizeof(llr_t));
  s2   = (llr_t *)calloc((size_t)n + 16, sizeof(llr_t));
  yp1  = (llr_t *)calloc((size_t)n + 16, sizeof(llr_t));
  yp2  = (llr_t *)calloc((size_t)n + 16, sizeof(llr_t));
  ext  = (llr_t *)calloc((size_t)n + 16, sizeof(llr_t));
  ext2 = (llr_t *)calloc((size_t)n + 16, sizeof(llr_t));
  orc_qpp_table(n, pi);

  /* demultiplex into lane layout, reference :1055-1161, tails :1167-1189 */
  for (p = 0; p < n; p++) {
    s0[ST(p)]  = y[3 * p];
    yp1[ST(p)] = y[3 * p + 1];
    yp2[ST(p)] = y[3 * p + 2];
  }
  for (i = 0; i < 3; i++) {
    const int16_t *t = y + 3 * n;
    s0[n + i] = s1[n + i] = s2[n + i] = t[2 * i];
    yp1[n + i] = t[2 * i + 1];
    s0[n + 8 + i] = s1[n + 8 + i] = s2[n + 8 + i] = t[6 + 2 * i];
    yp2[n + i] = t[7 + 2 * i];
  }

  orc_log_map16(s0, yp1, ext, n, 0, NULL, NULL);               /* reference :1199 */
  while (it++ < max_iterations) {                                /* reference :1201 */
    for (i = 0; i < n; i++) s2[ST(i)] = ext[ST(pi[i])];          /* :1209-1231 */
    orc_log_map16(s2, yp2, ext2, n, 1, NULL, NULL);              /* :1236 */
    for (i = 0; i < n; i++) {                                    /* :1241-1265 */
      int j = ST(pi[i]);
      s1[j] = adds(subs(ext2[ST(i)], ext[j]), s0[j]);
    }
    if (it > 1) {                                                /* :1267-1351 */
      uint32_t crc = 0, oldcrc = 0;
      int nb = n >> 3;
      memset(decoded_bytes, 0, (size_t)nb);
      for (i = 0; i < n; i++)
        if (ext2[ST(i)] > 0) decoded_bytes[pi[i] >> 3] |= (uint8_t)(0x80 >> (pi[i] & 7));
      /* oldcrc: little-endian load of the trailing crc_len bytes, :1306-1311 */
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
      case ORC_CRC16:   /* no byte swap in the reference, :1329-1333 */
        crc = orc_crc16(decoded_bytes, n - 16) >> 16;
        break;
      default:
        crc = orc_crc8(decoded_bytes, n - 8) >> 24;
        break;
      }
      if (crc == oldcrc && crc != 0) { ret = it; done = 1; break; }
    }
    if (it < max_iterations) {                                   /* :1354-1375 */
      orc_log_map16(s1, yp1, ext, n, 0, NULL, NULL);
      for (i = 0; i < n; i++) ext[i] = adds(subs(ext[i], s1[i]), s0[i]);
    }
  }
  if (!done) ret = it;                                           /* max+1 (uint8 wrap like :985) */
  free(pi); free(s0); free(s1); free(s2); free(yp1); free(yp2); free(ext); free(ext2);
  return ret;
}

#include <pthread.h>
/* pthread parallel-for over code blocks (the reference's own threading model for this
 * path is an OpenMP `parallel for` over the blocks of one transport block,
 * ulsch_decoding.c:1306-1310; libgomp is not in this image, so plain pthreads). */
typedef struct {
  const int16_t *y; int y_stride; uint8_t *out; int out_stride; uint8_t *ret;
  int nblk; uint16_t n; uint8_t max_it, crc_type; volatile int *next;
} orc_job_t;

static void *orc_worker(void *arg)
{
  orc_job_t *j = (orc_job_t *)arg;
  for (;;) {
    int b = __sync_fetch_and_add(j->next, 1);
    if (b >= j->nblk) break;
    j->ret[b] = orc_turbo_decoder16(j->y + (size_t)b * j->y_stride, j->out + (size_t)b * j->out_stride,
                                    j->n, j->max_it, j->crc_type, 0);
  }
  return NULL;
}

void orc_turbo_decoder16_batch(const int16_t *y, int y_stride, uint8_t *out, int out_stride,
                               uint8_t *ret, int nblk, uint16_t n, uint8_t max_iterations,
                               uint8_t crc_type, int nthreads)
{
  volatile int next = 0;
  orc_job_t job = { y, y_stride, out, out_stride, ret, nblk, n, max_iterations, crc_type, &next };
  pthread_t th[256];
  int t;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  for (t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, orc_worker, &job);
  orc_worker(&job);
  for (t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}
