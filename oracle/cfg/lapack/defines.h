/* oracle/cfg/lapack/defines.h -- TEST INFRASTRUCTURE ONLY.
 * Stands in for lapackpp's generated header (hook: lapackpp/include/lapack/mangling.h). */
#ifndef LAPACK_DEFINES_H
#define LAPACK_DEFINES_H
#define LAPACK_GLOBAL( lower, UPPER ) scipy_##lower##_
#define LAPACK_VERSION 31100
#endif
