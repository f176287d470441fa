/* A C caller written the way the reference's own receivers call the coding library, linked against
 * liboai_turbo_b200.so.  It includes BOTH the reference's PHY/CODING/defs.h and include/oai_turbo_b200.h (so the
 * compiler checks the two prototype sets against each other), declares the `tc` function pointer exactly as
 * dlsch_decoding.c:190-204 does, and runs
 *   (1) the per-code-block loop of dlsch_decoding.c:303-453 (generate_dummy_w -> lte_rate_matching_turbo_rx ->
 *       sub_block_deinterleaving_turbo -> memset c[r] -> tc(), err_flag rule) and the transport-block reassembly of
 *       :486-512, through the library's section-1 entry points, and
 *   (2) the batched equivalent (oai_turbo_submit_tbs: fused front end, OAI_BATCH_DL_STOP_AFTER_FAILURE, transport-block
 *       reassembly and return value on the GPU)
 * on one transport block read from a vector file, comparing c[r], ret and b with the expectations stored in the file
 * (produced by the oracle chain in tests/test_gpu_c_caller.py).  Test infrastructure, not product code.
 *
 * Built by __graft_entry__.build() (needs /root/reference for the headers only):
 *   gcc -std=gnu99 -fcommon -DNO_OPENAIR1 -include tests/c_abi/ref_prelude.h -I/root/reference/openair1 -Iinclude
 *       tests/c_abi/caller.c -o tests/c_abi/_build/caller -Lopenair4g_b200/lib -loai_turbo_b200 -Wl,-rpath,...
 *
 * Vector file (little endian, int32 header):
 *   magic 0x0A1C0DE5, C, Cminus, Kplus, Kminus, F, max_iterations, llr8_flag, G, Qm, Nl, Mdlharq, Kmimo, rvidx, round,
 *   expected_ret, b_bytes
 *   int16 dlsch_llr[G]; per block r: uint8 c_expected[Kr/8]; uint8 ret_expected[C]; uint8 b_expected[b_bytes]
 */
#include "PHY/CODING/defs.h"
#include "oai_turbo_b200.h"

int opp_enabled = 0;                      /* PHY/TOOLS/time_meas.h:38 */

#define NSOFT 1827072                     /* LTE_TRANSPORT/defs.h:62 */
#define MAX_NUM_DLSCH_SEGMENTS 16         /* LTE_TRANSPORT/defs.h:67 */

static void *malloc16(size_t n)
{
  void *p = NULL;
  return posix_memalign(&p, 16, n) ? NULL : p;
}

typedef struct {
  int32_t C, Cminus, Kplus, Kminus, F, max_turbo_iterations, llr8_flag, G, Qm, Nl, Mdlharq, Kmimo, rvidx, round,
          expected_ret, b_bytes;
} hdr_t;

typedef struct {                          /* the members of LTE_DL_UE_HARQ_t the loop touches (LTE_TRANSPORT/defs.h:515-553) */
  int16_t *w[MAX_NUM_DLSCH_SEGMENTS];
  int16_t *d[MAX_NUM_DLSCH_SEGMENTS];
  uint8_t *c[MAX_NUM_DLSCH_SEGMENTS];
  uint8_t *b;
  uint32_t RTC[MAX_NUM_DLSCH_SEGMENTS];
} harq_t;

static int check(const char *what, const hdr_t *h, harq_t *hq, uint32_t ret, uint8_t **c_exp, const uint8_t *b_exp)
{
  int bad = 0;
  for (int r = 0; r < h->C; r++) {
    int Kr = (r < h->Cminus) ? h->Kminus : h->Kplus;
    if (memcmp(hq->c[r], c_exp[r], Kr >> 3)) { printf("%s: c[%d] differs\n", what, r); bad = 1; }
  }
  if ((int)ret != h->expected_ret) { printf("%s: ret %u, expected %d\n", what, ret, h->expected_ret); bad = 1; }
  if (b_exp && h->b_bytes && memcmp(hq->b, b_exp, h->b_bytes)) { printf("%s: transport block b differs\n", what); bad = 1; }
  return bad;
}

int main(int argc, char **argv)
{
  setvbuf(stdout, NULL, _IONBF, 0);
  if (argc < 2) { fprintf(stderr, "usage: caller <vector file>\n"); return 2; }
  FILE *f = fopen(argv[1], "rb");
  int32_t magic = 0;
  hdr_t h;
  if (!f || fread(&magic, 4, 1, f) != 1 || magic != 0x0A1C0DE5 || fread(&h, sizeof(h), 1, f) != 1) {
    fprintf(stderr, "caller: bad vector file\n");
    return 2;
  }
  int16_t *dlsch_llr = malloc16(sizeof(int16_t) * (h.G + 16));
  uint8_t *c_exp[MAX_NUM_DLSCH_SEGMENTS], ret_exp[MAX_NUM_DLSCH_SEGMENTS], *b_exp = malloc(h.b_bytes + 1);
  if (fread(dlsch_llr, 2, h.G, f) != (size_t)h.G) return 2;
  for (int r = 0; r < h.C; r++) {
    int Kr = (r < h.Cminus) ? h.Kminus : h.Kplus;
    c_exp[r] = malloc(Kr >> 3);
    if (fread(c_exp[r], 1, Kr >> 3, f) != (size_t)(Kr >> 3)) return 2;
  }
  if (fread(ret_exp, 1, h.C, f) != (size_t)h.C) return 2;
  if (h.b_bytes && fread(b_exp, 1, h.b_bytes, f) != (size_t)h.b_bytes) return 2;
  fclose(f);

  harq_t hq;
  for (int r = 0; r < h.C; r++) {          /* allocation sizes of new_ue_dlsch, dlsch_decoding.c:129-143 */
    hq.w[r] = malloc16(3 * (6144 + 64) * sizeof(int16_t));
    hq.d[r] = malloc16(((3 * 8 * 6144) + 12 + 96) * sizeof(short));
    hq.c[r] = malloc16(((r == 0) ? 8 : 0) + 3 + 768);
    memset(hq.w[r], 0, 3 * (6144 + 64) * sizeof(int16_t));
    memset(hq.d[r], 0, ((3 * 8 * 6144) + 12 + 96) * sizeof(short));
  }
  hq.b = malloc16(h.b_bytes + 64);

  init_td16();                             /* lte_init.c:894-895 */
  init_td8();

  /* ---- (1) the reference's loop, dlsch_decoding.c:190-229, 303-453 --------------------------------------------- */
  time_stats_t st[7];
  memset(st, 0, sizeof(st));
  uint8_t (*tc)(int16_t *y, uint8_t *, uint16_t, uint16_t, uint16_t, uint8_t, uint8_t, uint8_t, time_stats_t *,
                time_stats_t *, time_stats_t *, time_stats_t *, time_stats_t *, time_stats_t *, time_stats_t *);
  if (h.llr8_flag == 0)
    tc = phy_threegpplte_turbo_decoder16;
  else
    tc = phy_threegpplte_turbo_decoder8;

  static short dummy_w[MAX_NUM_DLSCH_SEGMENTS][3 * (6144 + 64)];
  uint32_t r, r_offset = 0, Kr, Kr_bytes, err_flag = 0, E, ret = 0, offset;
  uint8_t crc_type;
  for (r = 0; r < (uint32_t)h.C; r++) {
    Kr = (r < (uint32_t)h.Cminus) ? h.Kminus : h.Kplus;
    Kr_bytes = Kr >> 3;
    memset(&dummy_w[r][0], 0, 3 * (6144 + 64) * sizeof(short));
    hq.RTC[r] = generate_dummy_w(4 + (Kr_bytes * 8), (uint8_t *)&dummy_w[r][0], (r == 0) ? h.F : 0);
    if (lte_rate_matching_turbo_rx(hq.RTC[r], h.G, hq.w[r], (uint8_t *)&dummy_w[r][0], dlsch_llr + r_offset, h.C,
                                   NSOFT, h.Mdlharq, h.Kmimo, h.rvidx, (h.round == 0) ? 1 : 0, h.Qm, h.Nl, r,
                                   &E) == -1) {
      printf("caller: Problem in rate_matching\n");
      return 1;
    }
    r_offset += E;
    sub_block_deinterleaving_turbo(4 + Kr, &hq.d[r][96], hq.w[r]);
    memset(hq.c[r], 0, Kr_bytes);
    crc_type = (h.C == 1) ? CRC24_A : CRC24_B;
    if (err_flag == 0)
      ret = tc(&hq.d[r][96], hq.c[r], Kr, 0, 0, h.max_turbo_iterations, crc_type, (r == 0) ? h.F : 0, &st[0], &st[1],
               &st[2], &st[3], &st[4], &st[5], &st[6]);
    if ((err_flag == 0) && (ret >= (1 + (uint32_t)h.max_turbo_iterations)))
      err_flag = 1;
  }
  if (err_flag == 1)
    ret = 1 + h.max_turbo_iterations;
  else {                                   /* reassembly, dlsch_decoding.c:486-512 */
    offset = 0;
    for (r = 0; r < (uint32_t)h.C; r++) {
      Kr = (r < (uint32_t)h.Cminus) ? h.Kminus : h.Kplus;
      Kr_bytes = Kr >> 3;
      if (r == 0) {
        memcpy(hq.b, &hq.c[0][(h.F >> 3)], Kr_bytes - (h.F >> 3) - ((h.C > 1) ? 3 : 0));
        offset = Kr_bytes - (h.F >> 3) - ((h.C > 1) ? 3 : 0);
      } else {
        memcpy(hq.b + offset, hq.c[r], Kr_bytes - ((h.C > 1) ? 3 : 0));
        offset += (Kr_bytes - ((h.C > 1) ? 3 : 0));
      }
    }
  }
  int bad = check("per-block loop", &h, &hq, ret, c_exp, err_flag ? NULL : b_exp);

  /* ---- (2) the batched equivalent: one submit for the whole transport block ------------------------------------ */
  oai_cb_desc_t cb[MAX_NUM_DLSCH_SEGMENTS];
  uint8_t status[MAX_NUM_DLSCH_SEGMENTS];
  r_offset = 0;
  for (r = 0; r < (uint32_t)h.C; r++) {
    Kr = (r < (uint32_t)h.Cminus) ? h.Kminus : h.Kplus;
    memset(hq.c[r], 0xA5, Kr >> 3);        /* must be overwritten or zeroed by the library */
    memset(hq.w[r], 0, 3 * (6144 + 64) * sizeof(int16_t));
    memset(&cb[r], 0, sizeof(cb[r]));
    cb[r].in = dlsch_llr + r_offset;
    cb[r].decoded_bytes = hq.c[r];
    cb[r].status = &status[r];
    cb[r].K = Kr;
    cb[r].max_iterations = h.max_turbo_iterations;
    cb[r].crc_type = (h.C == 1) ? CRC24_A : CRC24_B;
    cb[r].F = (r == 0) ? h.F : 0;
    cb[r].llr8 = h.llr8_flag;
    cb[r].decode_enable = 1;
    cb[r].dematch_enable = 1;
    cb[r].w = hq.w[r];
    cb[r].G = h.G; cb[r].Nsoft = NSOFT; cb[r].C = h.C; cb[r].r = r; cb[r].rvidx = h.rvidx;
    cb[r].clear = (h.round == 0) ? 1 : 0; cb[r].Qm = h.Qm; cb[r].Nl = h.Nl; cb[r].Mdlharq = h.Mdlharq; cb[r].Kmimo = h.Kmimo;
    cb[r].tb_id = 7;
    {                                      /* E of this block: lte_rate_matching.c:719-731 */
      uint32_t Gp = h.G / (h.Nl * h.Qm), q = Gp / h.C, m = Gp % h.C;
      r_offset += h.Nl * h.Qm * (q + ((r >= h.C - m) ? 1 : 0));
    }
  }
  oai_turbo_batch_t *hb;
  oai_tb_desc_t tbd;
  uint8_t tb_ret = 0xEE;
  uint32_t tb_valid = 0;
  memset(&tbd, 0, sizeof(tbd));
  memset(hq.b, 0x5A, h.b_bytes + 64);
  tbd.first_cb = 0; tbd.C = h.C; tbd.b = hq.b; tbd.b_capacity = h.b_bytes + 64; tbd.ret = &tb_ret; tbd.valid_bytes = &tb_valid;
  tbd.uplink = 0;
  if (oai_turbo_submit_tbs(cb, h.C, &tbd, 1, OAI_BATCH_DL_STOP_AFTER_FAILURE, -1, &hb) || oai_turbo_wait(hb)) {
    printf("caller: batched call failed: %s\n", oai_turbo_b200_last_error());
    return 1;
  }
  err_flag = (tb_ret >= 1 + h.max_turbo_iterations);
  for (r = 0; r < (uint32_t)h.C; r++) {
    uint8_t want = ret_exp[r];                 /* 0xFE in the vector = not decoded (after the first failing block) */
    if (status[r] != want) { printf("batched: status[%u] = %u, expected %u\n", r, status[r], want); bad = 1; }
  }
  if ((int)tb_valid != (err_flag ? 0 : h.b_bytes)) { printf("batched: %u bytes of b, expected %d\n", tb_valid, err_flag ? 0 : h.b_bytes); bad = 1; }
  if (err_flag && hq.b[0] != 0x5A) { printf("batched: b written on a NACK\n"); bad = 1; }
  bad |= check("batched submit (b and ret from the GPU)", &h, &hq, tb_ret, c_exp, err_flag ? NULL : b_exp);

  free_td16();
  free_td8();
  printf(bad ? "caller: FAILED\n" : "caller: OK (C=%d, K+=%d, per-block loop and batched submit agree with the reference chain)\n",
         h.C, h.Kplus);
  return bad;
}
