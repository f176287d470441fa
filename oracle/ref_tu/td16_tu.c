/* TEST INFRASTRUCTURE: compiles the reference 16-bit decoder from where it lies. */
#include "prelude.h"
#define TEST_DEBUG
#include "extern_3GPPinterleaver.h"
#include "PHY/CODING/3gpplte_turbo_decoder_sse_16bit.c"
