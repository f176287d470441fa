/* TEST INFRASTRUCTURE: compiles the reference 8-bit decoder from a build-time
 * copy (oracle/_ref/td8_patched.c) whose seven [n+16] stack arrays are enlarged
 * to [n+48] to remove the tail-write overflow. */
#include "prelude.h"
#define TEST_DEBUG
#include "extern_3GPPinterleaver.h"
#include "td8_patched.c"
