/* oracle/cfg/blas/defines.h -- TEST INFRASTRUCTURE ONLY.
 * Stands in for the header blaspp's configure step would generate
 * (hook: blaspp/include/blas/mangling.h:17-25).  Host BLAS = the OpenBLAS that
 * ships inside the scipy wheel, whose Fortran symbols carry a scipy_ prefix. */
#ifndef BLAS_DEFINES_H
#define BLAS_DEFINES_H
#define BLAS_FORTRAN_NAME( lower, UPPER ) scipy_##lower##_
#define BLAS_HAVE_OPENBLAS 1
#endif
