/* TEST INFRASTRUCTURE: compiles the reference's LTE_TRANSPORT/ulsch_decoding.c IN PLACE (shim4 shadows the headers that
 * need ASN.1-generated files) and drives the unmodified ulsch_decoding() from flat parameters, so that the port of its
 * demultiplexing / descrambling / e-fill / ACK-RI part (oracle/port/ulfront_port.c) and the whole uplink chain can be
 * pinned on it.  Contains no reference code: only the 36.212 5.2.2.8 column sets and the 36.213 bundling masks as data
 * (the reference instantiates them in LTE_TRANSPORT/vars.h together with a hundred unrelated globals). */
#include "PHY/defs.h"
#include "MAC_INTERFACE/defs.h"
#include "PHY/CODING/lte_interleaver2.h"          /* f1f2mat_old: passed to tc(), which ignores it */

unsigned char cs_ri_normal[4] = {1, 4, 7, 10}, cs_ri_extended[4] = {0, 3, 5, 8};
unsigned char cs_ack_normal[4] = {2, 3, 8, 9}, cs_ack_extended[4] = {1, 2, 6, 7};
int8_t wACK_RX[5][4] = {{-1, -1, -1, -1}, {-1, 1, -1, 1}, {-1, -1, 1, 1}, {-1, 1, 1, -1}, {1, 1, 1, 1}};
static void exit_stub(const char *s) { (void)s; }
static MAC_xface xface = {exit_stub};
MAC_xface *mac_xface = &xface;
uint8_t get_Qm_ul(uint8_t I_MCS) { return I_MCS < 11 ? 2 : (I_MCS < 21 ? 4 : 6); }      /* lte_mcs.c:57-67 */
uint8_t subframe2harq_pid(LTE_DL_FRAME_PARMS *fp, frame_t frame, uint8_t subframe) { (void)fp; (void)frame; (void)subframe; return 0; }
/* the CQI payload decoder is not on the pinned path: q[] (its soft input) is what the driver returns */
void phy_viterbi_lte_sse2(int8_t *y, uint8_t *decoded_bytes, uint16_t n) { (void)y; (void)decoded_bytes; (void)n; }

#include "PHY/LTE_TRANSPORT/ulsch_decoding.c"

typedef struct {
  uint32_t TBS, nb_rb, Nsymb_pusch, Nsymb_initial, Msc_initial, mcs, rvidx, round, O_ACK, O_RI, Or1, bundling, Nbundled, Ncp,
           beta_cqi_x8, beta_ri_x8, beta_ack_x8, rnti, subframe, Nid_cell, max_iter, Mdlharq, llr8;
} ref_ul_params_t;

/* Runs ulsch_decoding() once.  hq_state: opaque HARQ state kept by the caller across rounds (sizeof via ref_ul_state_size).
 * Outputs: e (G data soft bits), q_ACK[18], q_RI[6], q_cqi (Q_CQI int8), o_ACK[2], o_RI[1], c (16 x 768 bytes), b, status
 * of the call (return value). */
size_t ref_ul_state_size(void) { return sizeof(LTE_UL_eNB_HARQ_t); }
unsigned int ref_ulsch_decoding_run(const ref_ul_params_t *p, void *hq_state, int16_t *llr, int16_t *e_out, int e_cap,
                                    int16_t *q_ack, int16_t *q_ri, int8_t *q_cqi, int q_cap, uint8_t *o_ack, uint8_t *o_ri,
                                    uint8_t *c_out, uint8_t *b_out, int b_cap, int16_t *w_out)
{
  static __thread PHY_VARS_eNB enb;
  static __thread LTE_eNB_ULSCH_t ulsch;
  static __thread LTE_eNB_PUSCH pusch;
  LTE_UL_eNB_HARQ_t *h = hq_state;
  memset(&enb, 0, sizeof(enb));
  memset(&ulsch, 0, sizeof(ulsch));
  enb.lte_frame_parms.Ncp = p->Ncp; enb.lte_frame_parms.Nid_cell = p->Nid_cell;
  enb.lte_eNB_pusch_vars[0] = &pusch; pusch.llr = llr;
  enb.ulsch_eNB[0] = &ulsch;
  enb.proc[0].subframe_rx = p->subframe;
  ulsch.harq_processes[0] = h;
  ulsch.Mdlharq = p->Mdlharq; ulsch.max_turbo_iterations = p->max_iter; ulsch.bundling = p->bundling; ulsch.rnti = p->rnti;
  ulsch.beta_offset_cqi_times8 = p->beta_cqi_x8; ulsch.beta_offset_ri_times8 = p->beta_ri_x8; ulsch.beta_offset_harqack_times8 = p->beta_ack_x8;
  if (!h->b) {
    h->b = malloc16(MAX_ULSCH_PAYLOAD_BYTES);
    for (int r = 0; r < MAX_NUM_ULSCH_SEGMENTS; r++) {
      h->c[r] = malloc16(8 + 3 + 768);
      h->d[r] = malloc16(((3 * 8 * 6144) + 12 + 96) * sizeof(short));
      memset(h->c[r], 0, 8 + 3 + 768);
      memset(h->d[r], 0, ((3 * 8 * 6144) + 12 + 96) * sizeof(short));
    }
    memset(h->b, 0, MAX_ULSCH_PAYLOAD_BYTES);
  }
  h->TBS = p->TBS; h->nb_rb = p->nb_rb; h->Nsymb_pusch = p->Nsymb_pusch; h->Nsymb_initial = p->Nsymb_initial;
  h->Msc_initial = p->Msc_initial; h->mcs = p->mcs; h->rvidx = p->rvidx; h->round = p->round; h->O_ACK = p->O_ACK;
  h->O_RI = p->O_RI; h->Or1 = p->Or1; h->Nl = 1;
  unsigned int ret = ulsch_decoding(&enb, 0, 0, 0, p->Nbundled, p->llr8);
  memcpy(e_out, h->e, sizeof(int16_t) * (e_cap < MAX_NUM_CHANNEL_BITS ? e_cap : MAX_NUM_CHANNEL_BITS));
  memcpy(q_ack, h->q_ACK, sizeof(h->q_ACK));
  memcpy(q_ri, h->q_RI, sizeof(h->q_RI));
  memcpy(q_cqi, h->q, q_cap < (int)sizeof(h->q) ? q_cap : (int)sizeof(h->q));
  memcpy(o_ack, h->o_ACK, 2);
  o_ri[0] = h->o_RI[0];
  for (int r = 0; r < MAX_NUM_ULSCH_SEGMENTS; r++) memcpy(c_out + 768 * r, h->c[r], 768);
  memcpy(b_out, h->b, b_cap < MAX_ULSCH_PAYLOAD_BYTES ? b_cap : MAX_ULSCH_PAYLOAD_BYTES);
  if (w_out) memcpy(w_out, h->w, sizeof(h->w));
  return ret;
}
void ref_ul_state_free(void *hq_state)
{
  LTE_UL_eNB_HARQ_t *h = hq_state;
  if (h->b) free(h->b);
  for (int r = 0; r < MAX_NUM_ULSCH_SEGMENTS; r++) { if (h->c[r]) free(h->c[r]); if (h->d[r]) free(h->d[r]); }
  memset(h, 0, sizeof(*h));
}
