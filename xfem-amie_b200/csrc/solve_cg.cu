// solve_cg.cu -- ConjugateGradient::solve on the device.
//
// Control flow restated from solvers/conjugategradient.cpp:69-318 (line numbers in comments):
// homogeneous test, Jacobi-Richardson pre-smoothing, (re)start, the PCG loop, final residual,
// post-smoothing with xmin bookkeeping, acceptance at sqrt(realeps).  What is different is WHERE
// things run: the inner loop (:218-257) is three fused kernels per iteration
//     k_cg_dir  ->  k_spmv (q = A p, fused p.q)  ->  k_cg_update (x, r Kahan; fused r.D^-1 r)
// whose scalars (rho, beta, pq, alpha, nit, the while() test) are updated on the device, so the
// host only queues batches of iterations and polls a pinned copy of the state.
#include "batch_loop.cuh"
#include "kernels_vec_block.cuh"
#include <cmath>
#include <algorithm>

namespace {

struct CgRun
{
    amie_b200_ctx * ctx ;
    int precond ;
    uint64_t rowstart, colstart ;
} ;

int launch_dir(const CgRun & R, bool first)
{
    amie_b200_ctx * ctx = R.ctx ;
    VecArgs a = vec_args(ctx, first ? 0 : R.rowstart, first ? fin_kind(ctx, FIN_CG_RHO0) : FIN_STORE, first ? 0 : 1) ;
    int grid = vec_grid(ctx, a.end-a.begin) ;
    if(R.precond == PRECOND_BLOCK)
    {
        // one thread per node (kernels_vec_block.cuh)
        grid = vec_grid(ctx, (a.end-a.begin)/ctx->S) ;
        if(ctx->S == 2) { if(first) k_cg_dir_blk<2, true><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ; else k_cg_dir_blk<2, false><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ; }
        else            { if(first) k_cg_dir_blk<3, true><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ; else k_cg_dir_blk<3, false><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ; }
        ctx->stats.kernel_launches++ ;
        return first ? after_reduce(ctx, FIN_CG_RHO0) : AMIE_B200_OK ;
    }
    if(first)
    {
        // z = r ; P->precondition(r, z) ; p = z  over the whole vector (:183-186) ; last_rho over rows >= rowstart (:189).
        // r is 0 on [0, rowstart) so those entries add nothing to the sum.
        if(R.precond == PRECOND_JACOBI) k_cg_dir<PRECOND_JACOBI, true><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ;
        else                            k_cg_dir<PRECOND_NULL, true><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ;
        ctx->stats.kernel_launches++ ;
        return after_reduce(ctx, FIN_CG_RHO0) ;
    }
    else
    {
        if(R.precond == PRECOND_JACOBI) k_cg_dir<PRECOND_JACOBI, false><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ;
        else                            k_cg_dir<PRECOND_NULL, false><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ;
    }
    ctx->stats.kernel_launches++ ;
    return AMIE_B200_OK ;
}

int launch_update(const CgRun & R, bool first)
{
    amie_b200_ctx * ctx = R.ctx ;
    const int kind = first ? FIN_CG_RHO_FIRST : FIN_CG_RHO ;
    VecArgs a = vec_args(ctx, R.rowstart, fin_kind(ctx, kind), 1) ;
    int grid = vec_grid(ctx, a.end-a.begin) ;
    if(R.precond == PRECOND_BLOCK)
    {
        grid = vec_grid(ctx, (a.end-a.begin)/ctx->S) ;
        if(ctx->S == 2) k_cg_update_blk<2><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ;
        else            k_cg_update_blk<3><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ;
    }
    else if(R.precond == PRECOND_JACOBI) k_cg_update<PRECOND_JACOBI><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ;
    else                            k_cg_update<PRECOND_NULL><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ;
    ctx->stats.kernel_launches++ ;
    return after_reduce(ctx, kind) ;
}

int launch_smooth(const CgRun & R)
{
    amie_b200_ctx * ctx = R.ctx ;
    VecArgs a = vec_args(ctx, R.rowstart, fin_kind(ctx, FIN_STORE), 0) ;
    int grid = vec_grid(ctx, a.end-a.begin) ;
    if(R.precond == PRECOND_BLOCK)
    {
        grid = vec_grid(ctx, (a.end-a.begin)/ctx->S) ;
        if(ctx->S == 2) k_smooth_blk<2><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ;
        else            k_smooth_blk<3><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ;
    }
    else if(R.precond == PRECOND_JACOBI) k_smooth<PRECOND_JACOBI><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ;
    else                            k_smooth<PRECOND_NULL><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(a) ;
    ctx->stats.kernel_launches++ ;
    return after_reduce(ctx, FIN_STORE) ;
}

// r = sign (A x - b) on rows >= rowstart, fused |r|^2 -> *norm
int residual(const CgRun & R, double sign, uint64_t colstart, bool smoothing, double * norm)
{
    amie_b200_ctx * ctx = R.ctx ;
    SpmvCall c ;
    c.x = ctx->x ; c.b = ctx->b ; c.y = ctx->r ; c.minus_b = true ; c.sign = sign ;
    c.dot = DOT_YY ; c.finalize = FIN_STORE ;
    c.rowstart = R.rowstart ; c.colstart = colstart ; c.smoothing = smoothing ;
    int rc = launch_spmv(ctx, c) ;
    if(rc) return rc ;
    rc = ctx_sync_state(ctx, 2) ;
    if(rc) return rc ;
    *norm = std::sqrt(ctx->st_host[2].dot[0]) ;
    return AMIE_B200_OK ;
}

// one PCG iteration (:220-256), all decisions on the device
int queue_iteration(const CgRun & R)
{
    amie_b200_ctx * ctx = R.ctx ;
    int rc ;
    if((rc = launch_dir(R, false))) return rc ;
    SpmvCall c ;
    c.x = ctx->p ; c.y = ctx->q ; c.dot = DOT_YX ; c.finalize = FIN_CG_PQ ; c.check_stop = 1 ;
    c.rowstart = R.rowstart ; c.colstart = R.colstart ;
    if((rc = launch_spmv(ctx, c))) return rc ;
    return launch_update(R, false) ;
}

}

int solve_cg_resident(amie_b200_ctx * ctx, int precond_kind, double eps, int maxit, uint64_t nssor,
                      uint64_t rowstart, uint64_t colstart, uint64_t * nit_out, double * err_out, double * rho_out)
{
    if(precond_kind < AMIE_B200_PRECOND_JACOBI || precond_kind > AMIE_B200_PRECOND_BLOCK3X3)
    {
        ctx->set_error("pcg: preconditioner kind not on the device path (diagonal and node-block preconditioners and NullPreconditionner are)") ;
        return AMIE_B200_ERR_UNSUPPORTED ;
    }
    const int S = ctx->S ;
    const uint64_t N = ctx->N ;
    if(rowstart%S || colstart%S || rowstart > ctx->nb_global*(uint64_t)S)
    {
        ctx->set_error("pcg: rowstart/colstart must be multiples of the stride") ;
        return AMIE_B200_ERR_ARG ;
    }
    // row-partitioned context: rows are addressed locally from here on, columns stay global (dist_spmv)
    const uint64_t rowstart_global = rowstart ;
    if(ctx->dist) rowstart = dist_local_rowstart(ctx, rowstart) ;
    ctx_reset_solve_stats(ctx) ;
    cudaEvent_t ev0 = ctx->ev_a, ev1 = ctx->ev_b ;
    cudaEventRecord(ev0, ctx->stream) ;
    const bool block_precond = precond_kind == AMIE_B200_PRECOND_BLOCK2X2 || precond_kind == AMIE_B200_PRECOND_BLOCK3X3 ;
    CgRun R { ctx, precond_kind == AMIE_B200_PRECOND_NULL ? PRECOND_NULL : (block_precond ? PRECOND_BLOCK : PRECOND_JACOBI), rowstart, colstart } ;
    const size_t vbytes = N*sizeof(double) ;
    int rc ;
    uint64_t nit = 0 ;
    double err_final = 0., rho_final = 0. ;

    auto finish = [&](int r) -> int
    {
        cudaEventRecord(ev1, ctx->stream) ;
        cudaError_t e = cudaStreamSynchronize(ctx->stream) ;
        if(e == cudaSuccess) e = cudaGetLastError() ;
        if(e != cudaSuccess) { ctx->set_error(std::string("pcg: ")+cudaGetErrorString(e)) ; return AMIE_B200_ERR_CUDA ; }
        float ms = 0.f ;
        cudaEventElapsedTime(&ms, ev0, ev1) ;
        ctx->stats.solve_ms = ms ;
        ctx->stats.iterations = nit ;
        ctx_collect_spmv_times(ctx) ;
        if(nit_out) *nit_out = nit ;
        if(err_out) *err_out = err_final ;
        if(rho_out) *rho_out = rho_final ;
        return r ;
    } ;

    // :74-78  homogeneous right-hand side: the solver's x stays zero-initialised
    double bmax = 0. ;
    if((rc = ctx_max(ctx, ctx->b, N, 0, &bmax))) return rc ;
    if(getenv("AMIE_B200_TRACE")) { fprintf(stderr, "[amie_b200 dev %d] pcg: |b|max %g, N %llu (global %llu)\n", ctx->device, bmax, (unsigned long long)N, (unsigned long long)(ctx->nb_global*S)) ; fflush(stderr) ; }
    if(bmax < eps*eps)
    {
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->x, 0, vbytes, ctx->stream)) ;
        if(ctx->opt_verbose) fprintf(stderr, "\n CG %llu homogeneous. %g\n", (unsigned long long)N, bmax) ;
        ctx->stats.early_return = 1 ;
        return finish(1) ;
    }
    // :80-90
    // every diagonal preconditioner (t = v .* d) runs the Jacobi kernels on its own d
    if(R.precond != PRECOND_NULL && (rc = ctx_ensure_dinv(ctx, precond_kind))) return rc ;
    if(ctx->dist && (rc = dist_host_barrier(ctx))) return rc ;          // allocations done everywhere before the parts start waiting for one another

    const double realeps = std::max(1e-12, eps) ;                                   // :92
    // getForces().size() is the GLOBAL system size on a row-partitioned context
    const uint64_t Nglob = ctx->nb_global*(uint64_t)S ;
    const uint64_t Maxit = (maxit != -1) ? (uint64_t)(int64_t)maxit : Nglob/2 ;     // :93
    // :95-104 x = x0 was done by the caller (upload) ; :106-111
    if(rowstart) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->x, ctx->b, rowstart*sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream)) ;
    // assign() writes 0 to rows < rowstart on every call; the vectors keep those zeros because no
    // kernel below touches entries < rowstart
    for(double * v : { ctx->r, ctx->z, ctx->p, ctx->q })
        CUDA_TRY(ctx, cudaMemsetAsync(v, 0, ctx->vec_len*sizeof(double), ctx->stream)) ;
    double errmin = 1e9 ;                                                           // conjugategradient.h:29
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->xmin, ctx->x, vbytes, cudaMemcpyDeviceToDevice, ctx->stream)) ;   // :120

    // iterations queued per poll: enough work to hide the poll, few enough wasted no-op launches
    const double iter_bytes = (double)ctx->nnzb*(8*S*S+4)+(double)N*8*16 ;
    int batch = ctx->opt_batch > 0 ? ctx->opt_batch : (int)std::min(64., std::max(2., 300e-6/(iter_bytes/5e12))) ;

    for( ; nit < Maxit ; )                                                          // :121
    {
        ctx->stats.restarts++ ;
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->rc, 0, vbytes, ctx->stream)) ;           // :124-125
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->xc, 0, vbytes, ctx->stream)) ;
        if(nssor)                                                                   // :128-152
        {
            double err = 2, perr = 0 ;
            uint64_t iter = 0 ;
            while(iter++ < nssor && err > realeps)
            {
                perr = err ;
                if((rc = residual(R, 1., colstart, true, &err))) return rc ;        // :136-138
                if(err > perr) break ;                                              // :139 (NaN compares false, like the reference)
                if((rc = launch_smooth(R))) return rc ;                             // :141-149
            }
            CUDA_TRY(ctx, cudaMemsetAsync(ctx->xc, 0, vbytes, ctx->stream)) ;       // :151
        }
        double err0 = 0. ;
        if((rc = residual(R, -1., colstart, true, &err0))) return rc ;              // :153-159
        if(std::isnan(err0))                                                        // :160-165 (the reference prints the assembly and exit(0)s)
        {
            ctx->set_error("pcg: NaN initial residual") ;
            finish(0) ;
            return AMIE_B200_ERR_NAN ;
        }
        if(getenv("AMIE_B200_TRACE")) { fprintf(stderr, "[amie_b200 dev %d] pcg: restart at nit %llu, err0 %g\n", ctx->device, (unsigned long long)nit, err0) ; fflush(stderr) ; }
        if(nit == 0) errmin = err0 ;                                                // :167
        if(err0 < realeps)                                                          // :170-175
        {
            err_final = err0 ; rho_final = 0. ;
            if(ctx->opt_verbose) fprintf(stderr, "\n CG %llu converged after %llu iterations. Error : %g, last rho = 0\n", (unsigned long long)N, (unsigned long long)nit, err0) ;
            return finish(1) ;
        }
        if(ctx->opt_verbose) fprintf(stderr, "p\t%g\n", err0) ;                     // :177

        KrylovState s0 ;
        memset(&s0, 0, sizeof(s0)) ;
        s0.realeps = realeps ;
        s0.nit = nit ;
        s0.localnit = 0 ;
        s0.n_limit = Nglob ;                                                        // localnit < getForces().size()  (:218)
        if((rc = ctx_push_state(ctx, s0))) return rc ;

        if((rc = launch_dir(R, true))) return rc ;                                  // :183-186, :189
        {
            // :187  q = A*p through operator Vector(): every column (no colstart)
            SpmvCall c ;
            c.x = ctx->p ; c.y = ctx->q ; c.dot = DOT_YX ; c.finalize = FIN_CG_PQ_INIT ; c.check_stop = 0 ;
            c.rowstart = rowstart ; c.colstart = 0 ;
            if((rc = launch_spmv(ctx, c))) return rc ;                              // :190-197
        }
        if((rc = launch_update(R, true))) return rc ;                               // :199-210 (not counted in nit)

        // :218-257, queued speculatively (batch_loop.cuh)
        {
            const bool graph = want_graph(ctx, iter_bytes) ;
            const int nb_iter = graph ? (ctx->opt_batch > 0 ? ctx->opt_batch : 32) : batch ;
            if((rc = run_iteration_batches(ctx, ctx->graph_cg, graph, nb_iter, R.precond, rowstart, colstart, 3, 1,
                                           [&]() { return queue_iteration(R) ; }))) return rc ;
        }
        if((rc = ctx_sync_state(ctx, 2))) return rc ;
        CUDA_TRY(ctx, cudaGetLastError()) ;
        const KrylovState fin = ctx->st_host[2] ;
        nit = fin.nit ;
        const double last_rho = fin.last_rho ;
        const double rho = fin.rho ;
        if(fin.stop == STOP_PQ_INIT)                                                // :191-196
        {
            err_final = err0 ; rho_final = last_rho ;
            return finish(1) ;
        }

        double err = 0. ;
        if((rc = residual(R, 1., rowstart_global, true, &err))) return rc ;         // :266-267 (rowstart passed as colstart)
        if(err < errmin)                                                            // :268-272
        {
            errmin = std::sqrt(std::fabs(rho)) ;
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->xmin, ctx->x, vbytes, cudaMemcpyDeviceToDevice, ctx->stream)) ;
        }
        if(nssor)                                                                   // :274-302
        {
            uint64_t iters = 0 ;
            while(err > realeps && iters++ < Nglob)
            {
                double dummy ;
                if((rc = residual(R, 1., colstart, true, &dummy))) return rc ;      // :279
                if((rc = launch_smooth(R))) return rc ;                             // :280-288
                if((rc = ctx_sync_state(ctx, 2))) return rc ;
                const double perr = err ;
                err = std::sqrt(ctx->st_host[2].dot[0]) ;                           // :291  |D^-1 r|
                if(err > perr) break ;                                              // :294
                CUDA_TRY(ctx, cudaMemcpyAsync(ctx->xmin, ctx->x, vbytes, cudaMemcpyDeviceToDevice, ctx->stream)) ;  // :297
            }
            if(iters > 2)                                                           // :299-300
                CUDA_TRY(ctx, cudaMemcpyAsync(ctx->x, ctx->xmin, vbytes, cudaMemcpyDeviceToDevice, ctx->stream)) ;
        }
        if(std::min(err, std::sqrt(std::fabs(last_rho))) < std::sqrt(realeps))      // :305-309
        {
            err_final = err ; rho_final = last_rho ;
            if(ctx->opt_verbose) fprintf(stderr, "\n CG %llu converged after %llu iterations. Error : %g, last rho = %g\n", (unsigned long long)N, (unsigned long long)nit, err, last_rho) ;
            return finish(1) ;
        }
    }
    // :314-317
    {
        k_dot2<<<vec_grid(ctx, N-rowstart), AMIE_VEC_THREADS, 0, ctx->stream>>>(ctx->r, ctx->r, nullptr, rowstart, N, ctx->st, ctx->partials+AMIE_MAX_PARTIALS*2, fin_kind(ctx, FIN_STORE)) ;
        ctx->stats.kernel_launches++ ;
        if((rc = after_reduce(ctx, FIN_STORE))) return rc ;
        if((rc = ctx_sync_state(ctx, 2))) return rc ;
        err_final = std::sqrt(ctx->st_host[2].dot[0]) ;
        rho_final = err_final ;
        if(ctx->opt_verbose) fprintf(stderr, "\n CG %llu did not converge after %llu iterations. Error : %g\n", (unsigned long long)N, (unsigned long long)nit, err_final) ;
    }
    return finish(0) ;
}
