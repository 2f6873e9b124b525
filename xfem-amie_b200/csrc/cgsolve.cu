// cgsolve.cu -- Assembly::cgsolve's solver part with the displacement history kept in HBM.
//
// Reference (solvers/assembly.cpp): cgsolve (:1829-1911) solves from x0 = extrapolate() (:1772-1814) -- the last
// solution plus the last increment -- and then shifts displacementHistory (:1859-1868).  With the drop-in shim the
// host computes x0 and ships it to the device for every solve, and ships x back.  For callers that keep the loop on
// the device (values assembled there, fields recovered there) this entry point does the three steps in HBM:
// no x0 upload, no x download.
//
// Arithmetic: the reference's expression, per entry and in its order (explicit _rn, no contraction), including the
// term in dxxddb that is identically zero because the history never holds more than two vectors (:1859-1868 keeps
// size 2), and the NaN scrub of the newest vector (:1793-1794).
#include "context.h"
#include "group.h"
#include "dist.h"
#include "kernels_history.cuh"
#include <algorithm>

void history_destroy(amie_b200_ctx * ctx)
{
    for(int i = 0 ; i < 2 ; i++) { if(ctx->hist[i]) cudaFree(ctx->hist[i]) ; ctx->hist[i] = nullptr ; }
    ctx->hist_count = 0 ;
    ctx->hist_n = 0 ;
}

// x <- Assembly::extrapolate(factor).  Returns through *used what the reference's three cases amount to:
// 0 = fewer than two vectors: x0 = displacements, i.e. the resident x is left as it is (:1774-1779);
// 1 = extrapolated;  2 = size mismatch: history cleared and x0 = Vector(0), i.e. x zeroed (:1781-1785).
static int extrapolate_into_x(amie_b200_ctx * ctx, double factor, int * used)
{
    *used = 0 ;
    if(ctx->hist_count < 2) return AMIE_B200_OK ;
    if(ctx->hist_n != ctx->N)
    {
        ctx->hist_count = 0 ;
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->x, 0, ctx->vec_len*sizeof(double), ctx->stream)) ;
        *used = 2 ;
        return AMIE_B200_OK ;
    }
    if(ctx->N)
        k_extrapolate<<<vec_grid(ctx, ctx->N), AMIE_VEC_THREADS, 0, ctx->stream>>>(ctx->hist[0], ctx->hist[1], ctx->x, ctx->N, factor) ;
    CUDA_TRY(ctx, cudaGetLastError()) ;
    *used = 1 ;
    return AMIE_B200_OK ;
}

// displacementHistory update after a solve (:1859-1868)
static int history_push(amie_b200_ctx * ctx)
{
    const uint64_t n = ctx->N ;
    if(ctx->hist_count == 2 && ctx->hist_n == n)
    {
        std::swap(ctx->hist[0], ctx->hist[1]) ;                   // [0] = [1]
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hist[1], ctx->x, n*sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream)) ;
        return AMIE_B200_OK ;
    }
    if(ctx->hist_n != n || !ctx->hist[0] || !ctx->hist[1])
    {
        history_destroy(ctx) ;
        for(int i = 0 ; i < 2 ; i++) CUDA_TRY(ctx, cudaMalloc(&ctx->hist[i], (n ? n : 1)*sizeof(double))) ;
        ctx->hist_n = n ;
    }
    if(n) k_times_zero<<<vec_grid(ctx, n), AMIE_VEC_THREADS, 0, ctx->stream>>>(ctx->x, ctx->hist[0], n) ;
    CUDA_TRY(ctx, cudaGetLastError()) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hist[1], ctx->x, n*sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream)) ;
    ctx->hist_count = 2 ;
    return AMIE_B200_OK ;
}

extern "C" {

int amie_b200_cgsolve_resident(amie_b200_ctx * ctx, int precond_kind, double eps, uint64_t nssor,
                               uint64_t rowstart, uint64_t colstart, double factor,
                               uint64_t * nit_out, double * err_out, double * rho_out)
{
    if(!ctx) return AMIE_B200_ERR_ARG ;
    if(ctx->group)
    {
        // extrapolation and history shift are per entry: every device does them on its own rows
        uint64_t nit[GROUP_MAX] = {} ;
        double err[GROUP_MAX] = {}, rho[GROUP_MAX] = {} ;
        int rets[GROUP_MAX] = {} ;
        int rc = group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t, uint64_t)
        {
            const int r = dist_rank(c) ;
            rets[r] = amie_b200_cgsolve_resident(c, precond_kind, eps, nssor, rowstart, colstart, factor, nit+r, err+r, rho+r) ;
            return rets[r] < 0 ? rets[r] : 0 ;
        }) ;
        if(rc) return rc ;
        if(nit_out) *nit_out = nit[0] ;
        if(err_out) *err_out = err[0] ;
        if(rho_out) *rho_out = rho[0] ;
        return rets[0] ;
    }
    if(!ctx->have_structure || !ctx->have_values || !ctx->have_rhs)
    { ctx->set_error("cgsolve_resident needs the matrix (set_values / assemble) and the forces (upload_rhs) on the device") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    int used = 0 ;
    int rc = extrapolate_into_x(ctx, factor, &used) ;
    if(rc) return rc ;
    const int ret = solve_cg_resident(ctx, precond_kind, eps, -1, nssor, rowstart, colstart, nit_out, err_out, rho_out) ;
    if(ret < 0) return ret ;
    rc = history_push(ctx) ;
    if(rc) return rc ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return ret ;
}

int amie_b200_reset_history(amie_b200_ctx * ctx)
{
    if(!ctx) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t, uint64_t) { return amie_b200_reset_history(c) ; }) ;
    ctx->hist_count = 0 ;
    return AMIE_B200_OK ;
}

int amie_b200_extrapolate(amie_b200_ctx * ctx, double factor, double * x0_out, int * case_out)
{
    if(!ctx) return AMIE_B200_ERR_ARG ;
    if(ctx->group)
    {
        int cases[GROUP_MAX] = {} ;
        int rc = group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t) { return amie_b200_extrapolate(c, factor, x0_out ? x0_out+d0 : nullptr, cases+dist_rank(c)) ; }) ;
        if(!rc && case_out) *case_out = cases[0] ;
        return rc ;
    }
    if(!ctx->have_structure) { ctx->set_error("extrapolate before set_structure") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    int used = 0 ;
    int rc = extrapolate_into_x(ctx, factor, &used) ;
    if(rc) return rc ;
    if(x0_out) CUDA_TRY(ctx, cudaMemcpyAsync(x0_out, ctx->x, ctx->N*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    if(case_out) *case_out = used ;
    return AMIE_B200_OK ;
}

int amie_b200_push_history(amie_b200_ctx * ctx)
{
    if(!ctx) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t, uint64_t) { return amie_b200_push_history(c) ; }) ;
    if(!ctx->have_structure) { ctx->set_error("push_history before set_structure") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    int rc = history_push(ctx) ;
    if(rc) return rc ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return AMIE_B200_OK ;
}

}
