// krylov_scalars.cuh -- the scalar recurrences and loop-control tests of the reference solvers,
// executed on the device by ONE thread right after a fused reduction completes.
//
// CG      : solvers/conjugategradient.cpp:189-197 (restart), :218-257 (loop)
// BiCGStab: solvers/biconjugategradientstabilized.cpp:88-128
#pragma once
#include "common.cuh"

enum
{
    FIN_STORE = 0,           // dot[0] = a, dot[1] = b
    FIN_CG_RHO0,             // last_rho = r.z at (re)start                         (:189)
    FIN_CG_PQ_INIT,          // pq = q.p ; |pq| < 1e-12 last_rho -> stop ; alpha    (:190-197)
    FIN_CG_RHO_FIRST,        // rho of the first loop iteration (after the uncounted update :199-210)
    FIN_CG_PQ,               // pq = q.p ; |pq| < 1e-24 rho -> break ; alpha        (:234-240)
    FIN_CG_RHO,              // last_rho = rho ; nit++ ; while() test ; rho, beta   (:255-256, :218-225)
    FIN_BICG_RV,             // alpha = rho / (r_.v)                                (:100)
    FIN_BICG_OMEGA,          // omega = (t''.s'')/(t''.t'')                          (:115)
    FIN_BICG_RHO,            // rho_ = rho ; while() test ; nit++ ; rho ; beta       (:120, :88-94)
    FIN_BICG_RHO_INIT,       // same after the start-up half step, plus err0 and the threshold (:77-79)
    FIN_DEFER_SET,           // multi-GPU: red_local = (a, b); the scalar step runs after the all-reduce (dist.cu)
    FIN_DEFER_ADD,           // multi-GPU: red_local += (a, b)   (second / third row range of one SpMV)
} ;

#if defined(__CUDACC__)

// while(sqrt(|last_rho|) > realeps && localnit < N) { localnit++ ; rho = r.z ; beta = rho/last_rho ; ...
__device__ __forceinline__ void cg_loop_head(KrylovState * st, double rho_next)
{
    if(sqrt(fabs(st->last_rho)) > st->realeps && st->localnit < st->n_limit)
    {
        st->localnit++ ;
        st->rho = rho_next ;
        st->beta = rho_next/st->last_rho ;
    }
    else
        st->stop = STOP_LOOP_END ;
}

__device__ __forceinline__ void krylov_finalize(KrylovState * st, int kind, double a, double b)
{
    switch(kind)
    {
    case FIN_DEFER_SET :
        st->red_local[0] = a ;
        st->red_local[1] = b ;
        break ;
    case FIN_DEFER_ADD :
        st->red_local[0] += a ;
        st->red_local[1] += b ;
        break ;
    case FIN_STORE :
        st->dot[0] = a ;
        st->dot[1] = b ;
        break ;
    case FIN_CG_RHO0 :
        st->last_rho = a ;
        break ;
    case FIN_CG_PQ_INIT :
        st->pq = a ;
        if(fabs(a) < 1e-12*st->last_rho)
            st->stop = STOP_PQ_INIT ;
        else
            st->alpha = st->last_rho/a ;
        break ;
    case FIN_CG_RHO_FIRST :
        cg_loop_head(st, a) ;
        break ;
    case FIN_CG_PQ :
        st->pq = a ;
        if(fabs(a) < 1e-24*st->rho)
        {
            st->last_rho = st->rho ;
            st->stop = STOP_PQ_BREAK ;
        }
        else
            st->alpha = st->rho/a ;
        break ;
    case FIN_CG_RHO :
        st->last_rho = st->rho ;
        st->nit++ ;
        cg_loop_head(st, a) ;
        break ;
    case FIN_BICG_RV :
        st->rv = a ;
        st->alpha = st->rho/a ;
        break ;
    case FIN_BICG_OMEGA :
        st->ts = a ;
        st->tt = b ;
        st->omega = a/b ;
        break ;
    case FIN_BICG_RHO_INIT :
    {
        // err0 = sqrt(|r.r|) (:79) ; thr = max(|err0| veps veps, veps veps) (:88) ; realeps holds veps
        const double err0 = sqrt(fabs(b)) ;
        const double veps = st->realeps ;
        const double t0 = fabs(err0)*veps*veps, t1 = veps*veps ;
        st->dot[2] = err0 ;
        st->thr = t0 > t1 ? t0 : t1 ;
    }
    // fall through
    case FIN_BICG_RHO :
        // end of an iteration: rho_ = rho (:120); then the while() test of :88 with the rho just used
        st->rho_prev = st->rho ;
        if(st->nit < st->n_limit && fabs(st->rho)*st->nsq*st->nsq > st->thr)
        {
            st->nit++ ;
            st->rho = a ;                                               // :92
            st->beta = (a/st->rho_prev)*(st->alpha/st->omega) ;         // :94
        }
        else
            st->stop = STOP_LOOP_END ;
        break ;
    }
}

#endif
