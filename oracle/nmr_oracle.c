/*
 * oracle/nmr_oracle.c -- TEST INFRASTRUCTURE ONLY.
 * Instantiates the CPU restatement of the Neural-Mesh-Renderer rasterizer (see
 * nmr_oracle_impl.h for provenance and the "parity unpinned" note) for float and double.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load the resulting library.  Build: see oracle/Makefile (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORA_IMIN(a, b) ((a) < (b) ? (a) : (b))
#define ORA_IMAX(a, b) ((a) > (b) ? (a) : (b))

#define REAL float
#define FN(name) nmr_##name##_f32
#include "nmr_oracle_impl.h"
#undef REAL
#undef FN

#define REAL double
#define FN(name) nmr_##name##_f64
#include "nmr_oracle_impl.h"
#undef REAL
#undef FN

int nmr_oracle_abi_version(void) { return 1; }
