/* TEST INFRASTRUCTURE: instantiates the regenerated interleaver tables. */
#include "prelude.h"
#define INCL_INTERLEAVE
#include "extern_3GPPinterleaver.h"
#include "lte_interleaver.h"   /* generated into oracle/_ref/shim by gen_interleaver.py */
int opp_enabled = 0;
