/* TEST INFRASTRUCTURE: shadow of MAC_INTERFACE/defs.h -- the one member ulsch_decoding.c calls */
#ifndef ORACLE_SHIM4_MAC_DEFS_H
#define ORACLE_SHIM4_MAC_DEFS_H
typedef struct { void (*macphy_exit)(const char *); } MAC_xface;
#endif
