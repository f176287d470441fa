/* include-order check: oai_turbo_b200.h first (with -DOAI_TURBO_B200_WITH_REFERENCE_HEADERS it includes the
 * reference's PHY/CODING/defs.h itself), the reference header again afterwards; the tc pointer of
 * ulsch_decoding.c:275-312 must accept both decoders without a cast. */
#include "oai_turbo_b200.h"
#include "PHY/CODING/defs.h"

uint8_t (*tc_check[2])(int16_t *y, uint8_t *, uint16_t, uint16_t, uint16_t, uint8_t, uint8_t, uint8_t, time_stats_t *,
                       time_stats_t *, time_stats_t *, time_stats_t *, time_stats_t *, time_stats_t *,
                       time_stats_t *) = {phy_threegpplte_turbo_decoder16, phy_threegpplte_turbo_decoder8};
