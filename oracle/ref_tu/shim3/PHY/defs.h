/* TEST INFRASTRUCTURE: shadows openair1/PHY/defs.h for LTE_REFSIG/lte_gold.c (only the
 * frame-parameter field its table builders read).  Contains no reference code. */
#ifndef ORACLE_SHIM3_PHY_DEFS_H
#define ORACLE_SHIM3_PHY_DEFS_H
#include "../../prelude.h"
typedef struct { unsigned char Ncp; } LTE_DL_FRAME_PARMS;
typedef struct oracle_opaque_enb PHY_VARS_eNB;
typedef struct oracle_opaque_ue PHY_VARS_UE;
typedef struct oracle_opaque_rn PHY_VARS_RN;
typedef int mod_sym_t;
#endif
