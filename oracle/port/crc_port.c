/* TEST INFRASTRUCTURE (see oracle_port.h).
 * CRC24A/24B/16/8 as the reference computes them: MSB-first, zero initial value,
 * register kept left-aligned in 32 bits (reference: openair1/PHY/CODING/crc_byte.c:53-207;
 * polynomials :53-57, bitwise generator :66-84, byte-table walk :116-207).
 * Restated as a bitwise shift register; a byte-table walk and a bitwise walk of
 * the same polynomial are the same linear map, including the reference's
 * "residual bits" step (crc_byte.c:130-131), which consumes the top `resbit`
 * bits of the next byte. */
#include "oracle_port.h"

static uint32_t crc_bits(const uint8_t *in, int bitlen, uint32_t poly)
{
  uint32_t crc = 0;
  int i;
  for (i = 0; i < bitlen; i++) {
    uint32_t bit = (in[i >> 3] >> (7 - (i & 7))) & 1u;
    uint32_t top = (crc >> 31) ^ bit;
    crc <<= 1;
    if (top) crc ^= poly;
  }
  return crc;
}

uint32_t orc_crc24a(const uint8_t *in, int bitlen) { return crc_bits(in, bitlen, 0x864cfb00u); }
uint32_t orc_crc24b(const uint8_t *in, int bitlen) { return crc_bits(in, bitlen, 0x80006300u); }
uint32_t orc_crc16(const uint8_t *in, int bitlen)  { return crc_bits(in, bitlen, 0x10210000u); }

/* crc8 in the reference (crc_byte.c:193-207) drops the running value's low bits each
 * byte (`crc = table[..] << 24`), which for an 8-bit register is the plain CRC; the
 * residual step is the generic one. */
uint32_t orc_crc8(const uint8_t *in, int bitlen)   { return crc_bits(in, bitlen, 0x9B000000u); }
