/* Host-side probe: how fast can the box's CPU threads (a) memcpy and (b) range-check + pack int16 LLRs to int8
 * (AVX2 packs) from one page-locked-size buffer to another?  Decides whether compressing host inputs before the PCIe
 * copy can beat sending int16 (tools/README: evidence for DESIGN.md "host-fed path").
 *   gcc -O3 -mavx2 -pthread tools/src/host_pack_probe.c -o tools/bin/host_pack_probe */
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
typedef struct { const int16_t *src; int8_t *dst; size_t n; int mode; int over; } job_t;

static void *work(void *p)
{
  job_t *j = p;
  if (j->mode == 0) { memcpy(j->dst, j->src, j->n * 2); return NULL; }
  const __m256i *s = (const __m256i *)j->src;
  __m256i *d = (__m256i *)j->dst;
  __m256i acc = _mm256_setzero_si256();
  for (size_t i = 0; i < j->n / 32; ++i) {
    __m256i a = _mm256_loadu_si256(s + 2 * i), b = _mm256_loadu_si256(s + 2 * i + 1);
    __m256i pk = _mm256_permute4x64_epi64(_mm256_packs_epi16(a, b), 0xD8);
    /* exact iff unpacking gives the input back: accumulate (a ^ sext(lo)) | (b ^ sext(hi)) */
    __m256i lo = _mm256_cvtepi8_epi16(_mm256_castsi256_si128(pk)), hi = _mm256_cvtepi8_epi16(_mm256_extracti128_si256(pk, 1));
    acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_xor_si256(a, lo), _mm256_xor_si256(b, hi)));
    _mm256_stream_si256(d + i, pk);
  }
  j->over = !_mm256_testz_si256(acc, acc);
  return NULL;
}

int main(void)
{
  const size_t n = (size_t)768 << 20;          /* int16 elements: 1.5 GB, the size of one bench step's input */
  int16_t *src = aligned_alloc(64, n * 2);
  int8_t *dst = aligned_alloc(64, n * 2);
  for (size_t i = 0; i < n; ++i) src[i] = (int16_t)((i * 2654435761u >> 20) % 33) - 16;
  memset(dst, 1, n * 2);
  int nt[] = {1, 2, 4, 8, 12, 16, 24, 32};
  for (int mode = 0; mode < 2; ++mode)
    for (unsigned k = 0; k < sizeof(nt) / sizeof(nt[0]); ++k) {
      int T = nt[k];
      pthread_t th[64]; job_t jb[64];
      double best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        double t0 = now();
        for (int t = 0; t < T; ++t) {
          size_t lo = n / T * t & ~(size_t)31, hi = (t == T - 1) ? n : (n / T * (t + 1) & ~(size_t)31);
          jb[t] = (job_t){src + lo, dst + (mode ? lo : 2 * lo), hi - lo, mode, 0};
          pthread_create(&th[t], NULL, work, &jb[t]);
        }
        for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
        double dt = now() - t0;
        if (dt < best) best = dt;
      }
      printf("%s threads=%2d: %.1f ms for %.2f GB of int16 -> %.1f GB/s of input\n", mode ? "check+pack int16->int8" : "memcpy                ",
             T, best * 1e3, n * 2 / 1e9, n * 2 / 1e9 / best);
    }
  return 0;
}
