/* TEST INFRASTRUCTURE: shadows openair1/PHY/defs.h for lte_rate_matching.c and
 * lte_segmentation.c.  Contains no reference code. */
#ifndef ORACLE_SHIM_PHY_DEFS_H
#define ORACLE_SHIM_PHY_DEFS_H
#include "../../prelude.h"
#define cmin(a,b) ((a)<(b) ? (a) : (b))
#define cmax(a,b) ((a)>(b) ? (a) : (b))
#define MAX_NUM_DLSCH_SEGMENTS 16
#ifndef NO_OPENAIR1
#define NO_OPENAIR1
#endif
#include "PHY/CODING/defs.h"
#endif
