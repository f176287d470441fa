/* TEST INFRASTRUCTURE: empty shadow (see shim4/PHY/defs.h) */
