/* TEST INFRASTRUCTURE: prelude for compiling the reference sources in place
 * (SURVEY.md Appendix B).  Contains no reference code. */
#ifndef ORACLE_REF_PRELUDE_H
#define ORACLE_REF_PRELUDE_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
typedef int lte_prefix_type_t;
#define msg printf
#endif
