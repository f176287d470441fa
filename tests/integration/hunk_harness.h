/* Test harness for tests/test_integration.py: minimal stand-ins for the reference types that the hunks added by
 * integration/0002 and 0003 touch (member names as in openair1/PHY/LTE_TRANSPORT/defs.h:408-440,515-553,
 * openair1/PHY/defs.h), so that the ADDED code can be compiled without the reference's ASN.1-generated headers.
 * The real prototypes come from the reference's PHY/CODING/defs.h when /root/reference is present. */
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#define NSOFT 1827072
#define MAX_NUM_DLSCH_SEGMENTS 16
#define MAX_NUM_ULSCH_SEGMENTS 16
#define PHY 0
#define LOG_E(c, ...) fprintf(stderr, __VA_ARGS__)
#define start_meas(x) (void)(x)
#define stop_meas(x) (void)(x)
#define VCD_SIGNAL_DUMPER_DUMP_FUNCTION_BY_NAME(a, b) (void)0
#define VCD_SIGNAL_DUMPER_FUNCTIONS_PHY_ENB_ULSCH_DECODING 0
typedef struct { uint32_t C, Cminus, Cplus, Kminus, Kplus, F, TBS; uint8_t rvidx, round, Qm, Nl; uint8_t *c[16], *b; int16_t *w[16], *e; } harq_t;
typedef struct { uint8_t max_turbo_iterations, Mdlharq, Kmimo; } sch_t;
typedef struct { time_stats_t ulsch_turbo_decoding_stats; } enb_t;
