/* TEST INFRASTRUCTURE (see oracle_port.h).  8-bit decoder restatement: placeholder
 * until the 8-bit row is built; returns 254 = "not implemented". */
#include "oracle_port.h"
uint8_t orc_turbo_decoder8(const int16_t *y, uint8_t *decoded_bytes, uint16_t n,
                           uint8_t max_iterations, uint8_t crc_type, uint8_t F)
{
  (void)y; (void)decoded_bytes; (void)n; (void)max_iterations; (void)crc_type; (void)F;
  return 254;
}
