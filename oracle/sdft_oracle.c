/*
 * oracle/sdft_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Instantiates the CPU restatement (sdft_oracle_impl.h) for the four supported type pairs
 * (c/src/sdft/sdft.h:101-125; long double is out of scope).  Exported symbols are
 * oracle_<td><fd>_{alloc,free,reset,sdft_n,isdft_n,size,cursor,get_twiddles,get_state}.
 *
 * Build (see oracle/Makefile): gcc -std=gnu99 -O2 -ffp-contract=off -fPIC -shared ... -lm
 * -ffp-contract=off keeps the rounding points of a plain x86-64 build of the reference.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define OR_TD float
#define OR_FD float
#define OR_SUFFIX f32f32
#define OR_COS cosf
#define OR_SIN sinf
#define OR_ACOS acosf
#include "sdft_oracle_impl.h"
#undef OR_TD
#undef OR_FD
#undef OR_SUFFIX
#undef OR_COS
#undef OR_SIN
#undef OR_ACOS

#define OR_TD float
#define OR_FD double
#define OR_SUFFIX f32f64
#define OR_COS cos
#define OR_SIN sin
#define OR_ACOS acos
#include "sdft_oracle_impl.h"
#undef OR_TD
#undef OR_FD
#undef OR_SUFFIX
#undef OR_COS
#undef OR_SIN
#undef OR_ACOS

#define OR_TD double
#define OR_FD float
#define OR_SUFFIX f64f32
#define OR_COS cosf
#define OR_SIN sinf
#define OR_ACOS acosf
#include "sdft_oracle_impl.h"
#undef OR_TD
#undef OR_FD
#undef OR_SUFFIX
#undef OR_COS
#undef OR_SIN
#undef OR_ACOS

#define OR_TD double
#define OR_FD double
#define OR_SUFFIX f64f64
#define OR_COS cos
#define OR_SIN sin
#define OR_ACOS acos
#include "sdft_oracle_impl.h"
