// common.cuh -- shared device/host definitions of the B200 Krylov path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#define AMIE_MAX_PARTIALS 8192          // max thread blocks taking part in one fused reduction
#define AMIE_VEC_THREADS 256

// Scalars of the running Krylov iteration.  They live in HBM so that the loop-control
// decisions of the reference (solvers/conjugategradient.cpp:218-257,
// solvers/biconjugategradientstabilized.cpp:88-128) are taken ON THE DEVICE by the last
// thread block of each fused reduction: the host queues iterations speculatively and every
// kernel returns at once when `stop` is set, so there is no host round trip per iteration.
struct KrylovState
{
    // CG
    double rho ;            // r.z of the running iteration
    double last_rho ;
    double pq ;
    double alpha ;
    double beta ;
    // BiCGStab
    double omega ;
    double rho_prev ;       // rho_
    double rv ;             // r_.v
    double ts, tt ;         // t''.s'' , t''.t''
    double thr ;            // max(|err0| veps^2, veps^2)
    double nsq ;            // vsize (multiplied twice, as the reference does)
    // generic results of fused reductions (residual norms etc.)
    double dot[4] ;
    // multi-GPU: per-rank partial sums of the running fused reduction, and their all-reduced values
    double red_local[2] ;
    double red_global[2] ;
    // control
    double realeps ;
    unsigned long long nit ;
    unsigned long long localnit ;
    unsigned long long n_limit ;     // localnit < N   /  lastit
    int stop ;                       // 0 = keep iterating
    int pad ;
    unsigned int ticket[8] ;         // last-block tickets, one per kernel family
} ;

enum
{
    STOP_NONE = 0,
    STOP_LOOP_END = 1,        // while() condition false (converged or localnit limit)
    STOP_PQ_BREAK = 2,        // |pq| < 1e-24 rho  (conjugategradient.cpp:235-239)
    STOP_PQ_INIT = 3,         // |pq| < 1e-12 last_rho at (re)start (:191-196)
} ;

enum { TICKET_SPMV = 0, TICKET_UPDATE = 1, TICKET_DIR = 2, TICKET_MISC = 3 } ;

#define CUDA_TRY(ctx, expr) do { cudaError_t _e = (expr) ; if(_e != cudaSuccess) { \
        (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)) ; return AMIE_B200_ERR_CUDA ; } } while(0)

#if defined(__CUDACC__)

__device__ __forceinline__ double warp_sum(double v)
{
    #pragma unroll
    for(int o = 16 ; o > 0 ; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o) ;
    return v ;
}

// Deterministic grid-wide sum of NV per-thread values.
// Every block writes its partial to partials[k*AMIE_MAX_PARTIALS + blockIdx.x]; the block that
// draws the last ticket re-reads all partials in a fixed order and returns true with the totals in
// out[] (valid in thread 0).  No floating-point atomics: same grid -> same bits, run to run.
template<int NV, int THREADS>
__device__ __forceinline__ bool grid_sum(double (&v)[NV], double * partials, unsigned int * ticket, double (&out)[NV])
{
    __shared__ double sm[NV][THREADS/32] ;
    __shared__ bool is_last ;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5 ;
    #pragma unroll
    for(int k = 0 ; k < NV ; k++)
    {
        double s = warp_sum(v[k]) ;
        if(lane == 0) sm[k][wid] = s ;
    }
    __syncthreads() ;
    if(wid == 0)
    {
        #pragma unroll
        for(int k = 0 ; k < NV ; k++)
        {
            double s = lane < THREADS/32 ? sm[k][lane] : 0. ;
            s = warp_sum(s) ;
            if(lane == 0) partials[k*AMIE_MAX_PARTIALS+blockIdx.x] = s ;
        }
        if(lane == 0)
        {
            __threadfence() ;
            unsigned int t = atomicAdd(ticket, 1u) ;
            is_last = (t == gridDim.x-1) ;
        }
    }
    __syncthreads() ;
    if(!is_last) return false ;
    __threadfence() ;
    #pragma unroll
    for(int k = 0 ; k < NV ; k++)
    {
        double s = 0. ;
        for(unsigned int i = threadIdx.x ; i < gridDim.x ; i += THREADS)
            s += __ldcg(partials+k*AMIE_MAX_PARTIALS+i) ;
        s = warp_sum(s) ;
        __syncthreads() ;
        if(lane == 0) sm[k][wid] = s ;
    }
    __syncthreads() ;
    if(wid == 0)
    {
        #pragma unroll
        for(int k = 0 ; k < NV ; k++)
        {
            double s = lane < THREADS/32 ? sm[k][lane] : 0. ;
            out[k] = warp_sum(s) ;
        }
        if(lane == 0) *ticket = 0u ;
    }
    return true ;
}

#endif
