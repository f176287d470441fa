/* TEST INFRASTRUCTURE: reference turbo encoder (vector generation only). */
#include "prelude.h"
#include "PHY/CODING/3gpplte_sse.c"
