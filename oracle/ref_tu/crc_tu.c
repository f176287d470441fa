/* TEST INFRASTRUCTURE: reference CRC routines. */
#include "prelude.h"
#include "PHY/CODING/crc_byte.c"
