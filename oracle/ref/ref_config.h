/* Force-included (-include) ahead of every reference TU when building oracle/_ref.
 *
 * TEST INFRASTRUCTURE ONLY. The reference picks its NN backend at compile time in
 * /root/reference/config.h:21-28 (USE_OPENCL on, USE_BLAS off). That header cannot be
 * shadowed with -I because `#include "config.h"` resolves next to the including source,
 * so we pre-define its include guard and state the switches we want instead:
 * BLAS + OpenBLAS on (the im2col + cblas_sgemm path, Network.cpp:344-447), OpenCL off.
 * Nothing from the reference is copied; the sources are compiled where they lie.
 */
#ifndef LB2_REF_CONFIG_H
#define LB2_REF_CONFIG_H
#define CONFIG_INCLUDED

#define HAVE_SELECT
#define GETTIMEOFDAY
#define USE_OPTIONS
#define USE_BLAS
#define USE_OPENBLAS
#define USE_SEARCH
#define PROGRAM_NAME "Leela"
#define PROGRAM_VERSION "0.11.0"
#define MAX_CPUS 64

#include <sys/time.h>
#include <time.h>
typedef int int32;       typedef unsigned int uint32;
typedef short int16;     typedef unsigned short uint16;
typedef signed char int8; typedef unsigned char uint8;
typedef long long int int64; typedef unsigned long long int uint64;
typedef struct timeval rtime_t;

#ifdef __cplusplus
#include <string>   /* MCPolicy.h uses std::string without including <string> */
#endif
#endif
