/* Prelude for compiling a caller against the reference's PHY/CODING/defs.h with -DNO_OPENAIR1:
 * in that mode defs.h pulls in PHY/TOOLS/time_meas.h only, and the one type it still needs from PHY/impl_defs_lte.h
 * has to be supplied.  In the full tree (no NO_OPENAIR1) PHY/defs.h provides it and this file is not used. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
typedef int lte_prefix_type_t;
#define msg printf
