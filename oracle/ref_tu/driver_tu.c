/* TEST INFRASTRUCTURE: pthread driver that runs the compiled reference decoder over many
 * code blocks ("threaded across code blocks", the reference's own model for this path is
 * the OpenMP parallel-for of ulsch_decoding.c:1306-1310).  Contains no reference code. */
#include "prelude.h"
#include <pthread.h>

unsigned char ref_phy_threegpplte_turbo_decoder16(short *y, unsigned char *decoded_bytes, unsigned short n,
    unsigned short f1, unsigned short f2, unsigned char max_iterations, unsigned char crc_type, unsigned char F,
    void *, void *, void *, void *, void *, void *, void *);
unsigned char ref_phy_threegpplte_turbo_decoder8(short *y, unsigned char *decoded_bytes, unsigned short n,
    unsigned short f1, unsigned short f2, unsigned char max_iterations, unsigned char crc_type, unsigned char F,
    void *, void *, void *, void *, void *, void *, void *);

typedef struct {
  short *y; long y_stride; unsigned char *out; long out_stride; unsigned char *ret;
  int nblk, total, n, max_it, crc, which; volatile int *next;
} job_t;

static void *worker(void *arg)
{
  job_t *j = (job_t *)arg;
  long long stats[7][8];
  memset(stats, 0, sizeof(stats));
  for (;;) {
    int b = __sync_fetch_and_add(j->next, 1), i;
    if (b >= j->total) break;
    i = b % j->nblk;                       /* cycle over the distinct inputs */
    if (b < j->nblk) {
      unsigned char r = (j->which == 8 ? ref_phy_threegpplte_turbo_decoder8 : ref_phy_threegpplte_turbo_decoder16)(
          j->y + i * j->y_stride, j->out + i * j->out_stride, (unsigned short)j->n, 0, 0,
          (unsigned char)j->max_it, (unsigned char)j->crc, 0, stats[0], stats[1], stats[2], stats[3], stats[4], stats[5], stats[6]);
      j->ret[i] = r;
    } else {
      unsigned char scratch[768 + 64] __attribute__((aligned(16)));
      (j->which == 8 ? ref_phy_threegpplte_turbo_decoder8 : ref_phy_threegpplte_turbo_decoder16)(
          j->y + i * j->y_stride, scratch, (unsigned short)j->n, 0, 0,
          (unsigned char)j->max_it, (unsigned char)j->crc, 0, stats[0], stats[1], stats[2], stats[3], stats[4], stats[5], stats[6]);
    }
  }
  return NULL;
}

/* y rows must be 16-byte aligned (y_stride multiple of 8 int16, base aligned); `total` >= nblk
 * decodes are performed, the first nblk of them write out/ret. */
void ref_td_batch(short *y, long y_stride, unsigned char *out, long out_stride, unsigned char *ret,
                  int nblk, int total, int n, int max_it, int crc, int which, int nthreads)
{
  volatile int next = 0;
  job_t job = { y, y_stride, out, out_stride, ret, nblk, total, n, max_it, crc, which, &next };
  pthread_t th[512];
  int t;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 512) nthreads = 512;
  for (t = 1; t < nthreads; t++) {
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, 8u << 20);      /* the decoder keeps ~0.3 MB of VLAs on the stack */
    pthread_create(&th[t], &at, worker, &job);
  }
  worker(&job);
  for (t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
}
