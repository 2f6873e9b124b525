// launch.cuh -- grid sizing and launch helpers (host side, included by the .cu files).
#pragma once
#include "context.h"
#include "dist.h"
#include "spmv_args.h"
#include "kernels_vec.cuh"
#include "kernels_setup.cuh"

struct SpmvCall
{
    const double * x = nullptr ;
    const double * b = nullptr ;
    double * y = nullptr ;
    const double * w = nullptr ;
    const double * d = nullptr ;
    int dot = DOT_NONE ;
    bool minus_b = false ;
    double sign = 1. ;
    uint64_t rowstart = 0 ;     // DOF units, multiple of S
    uint64_t colstart = 0 ;
    int finalize = FIN_STORE ;
    int check_stop = 0 ;
    bool smoothing = false ;    // stats only
} ;

// spmv_launch.cu: the ONE translation unit that instantiates the block-row SpMV kernels (each of them is a few hundred
// KB of SASS; included from every caller they were compiled five times over)
int launch_spmv(amie_b200_ctx * ctx, const SpmvCall & c) ;
// block rows [blk_row0, blk_row0+blk_nrows) of the local matrix; `finalize` overrides c.finalize
int launch_spmv_range(amie_b200_ctx * ctx, const SpmvCall & c, uint32_t blk_row0, uint32_t blk_nrows, int finalize) ;
// amie_b200_spmv_resident: one launch of kernel selection `variant` (tuning configurations included), y = A x on resident vectors
int launch_spmv_variant(amie_b200_ctx * ctx, const SpmvCall & c, int variant, bool insolve) ;

// fused reductions of the vector kernels: on one device the last block runs the scalar step itself;
// on a row-partitioned context it only stores the rank's partial sums and dist_finalize() follows
static inline int fin_kind(const amie_b200_ctx * ctx, int kind) { return ctx->dist ? FIN_DEFER_SET : kind ; }
static inline int after_reduce(amie_b200_ctx * ctx, int kind) { return ctx->dist ? dist_finalize(ctx, kind) : AMIE_B200_OK ; }

static inline VecArgs vec_args(amie_b200_ctx * ctx, uint64_t begin, int finalize, int check_stop)
{
    VecArgs a ;
    a.x = ctx->x ; a.r = ctx->r ; a.z = ctx->z ; a.p = ctx->p ; a.q = ctx->q ; a.xc = ctx->xc ; a.rc = ctx->rc ;
    a.d = ctx->dinv ;
    a.r_ = ctx->w[0] ; a.p_ = ctx->w[1] ; a.v = ctx->w[2] ; a.s = ctx->w[3] ; a.s_ = ctx->w[4] ; a.t = ctx->w[5] ;
    a.begin = begin ; a.end = ctx->N ;
    a.st = ctx->st ; a.partials = ctx->partials+AMIE_MAX_PARTIALS*2 ;
    a.finalize = finalize ; a.check_stop = check_stop ;
    return a ;
}
