// getrf.cu -- LU with partial pivoting (placeholder until the GPU panel lands).
#include "runtime.hh"
extern "C" int sb200_getrf_d(sb200_matrix_t A, int64_t* pivots, const sb200_options_t* opts, int64_t* info)
{
    (void) A; (void) pivots; (void) opts; (void) info;
    return SB200_ENOTSUP;
}
