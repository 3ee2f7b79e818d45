/* oracle/standin/autoconfig.h -- TEST INFRASTRUCTURE.
 * What the reference's ./configure would emit on x86-64 Linux with gcc (the
 * feature macros src/config.h and src/Chromosome.cpp test for).  Written for the
 * out-of-tree oracle build only; nothing in the product includes it. */
#ifndef GASELECT_B200_ORACLE_AUTOCONFIG_H
#define GASELECT_B200_ORACLE_AUTOCONFIG_H
#define HAVE_INTTYPES_H 1
#define HAVE_STDINT_H 1
#define HAVE_STRINGS_H 1
#define HAVE_LIMITS_H 1
#define HAVE_CLIMITS 1
#define HAVE_UNSIGNED_LONG_LONG 1
#define HAVE_UINT8_16_MAX 1
#define HAVE_GCC_CTZL 1
#define HAVE_GCC_CTZLL 1
#define HAVE_FFSL 1
#define HAVE_BUILTIN_POPCOUNTL 1
#define HAVE_BUILTIN_POPCOUNTLL 1
#define HAVE_PTHREAD_H 1
#endif
