// spmv_args.h -- what a block-row SpMV launch is told (shared by the kernels and by every caller of launch_spmv)
#pragma once
#include "common.cuh"
#include "krylov_scalars.cuh"

enum { DOT_NONE = 0, DOT_YX = 1, DOT_YY = 2, DOT_YW = 3, DOT_OMEGA = 4 } ;

struct SpmvArgs
{
    const uint32_t * rowptr ;
    const uint32_t * col ;
    const double * vals ;
    const double * x ;
    const double * b ;          // MINUS_B: y = sign*(A x - b)
    double * y ;
    const double * w ;          // DOT_YW: sum y.w ; DOT_OMEGA: s
    const double * d ;          // DOT_OMEGA: inverse diagonal (NULL = identity)
    uint32_t row0 ;             // first block row computed (rowstart / S)
    uint32_t nrows ;            // block rows computed
    uint32_t colstart_blk ;     // block columns < this are skipped (colstart / S)
    double sign ;
    KrylovState * st ;          // may be NULL (plain SpMV)
    double * partials ;
    int finalize ;              // FIN_* (krylov_scalars.cuh)
    int check_stop ;
} ;
