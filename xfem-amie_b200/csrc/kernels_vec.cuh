// kernels_vec.cuh -- fused streaming kernels of the Krylov loops (HBM-bound, one pass each).
//
// Every kernel fuses what the reference does in separate OpenMP sweeps, keeps the reference's
// operation ORDER per entry (explicit __dmul_rn/__dadd_rn: no FMA contraction, so each entry is
// rounded exactly like the CPU code), and ends with a deterministic grid reduction whose last
// block updates the loop-control scalars (krylov_scalars.cuh).
//
//   k_cg_dir     : z = D^-1 r ; p = beta p + z        conjugategradient.cpp:222, :227-231  (3R+1W)
//                  restart form: z = D^-1 r ; p = z ; rho0 = r.z            :183-189
//   k_cg_update  : x += alpha p ; r -= alpha q (Kahan) ; rho' = r.(D^-1 r)  :242-253, :222-223 (7R+4W)
//   k_smooth     : r = D^-1 r ; x -= r (Kahan) ; |r|^2                      :141-149, :280-291
//   k_bicg_p     : p = r + (p - v w) beta ; p^ = D^-1 p                     biconjugategradientstabilized.cpp:95-97
//   k_bicg_s     : s = r - v alpha ; s^ = D^-1 s                            :101-104
//   k_bicg_xr    : x += p^ alpha + s^ omega ; r = s - t omega ; r.r_ ; r.r  :118-119, :92
#pragma once
#include "common.cuh"
#include "krylov_scalars.cuh"

enum { PRECOND_JACOBI = 0, PRECOND_NULL = 1 } ;

struct VecArgs
{
    double * x ; double * r ; double * z ; double * p ; double * q ;
    double * xc ; double * rc ;
    const double * d ;
    // BiCGStab extras
    double * r_ ; double * p_ ; double * v ; double * s ; double * s_ ; double * t ;
    uint64_t begin ;        // rowstart
    uint64_t end ;          // N
    KrylovState * st ;
    double * partials ;
    int finalize ;
    int check_stop ;
} ;

#define VEC_LOOP(i, a) for(uint64_t i = (a).begin+(uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < (a).end ; i += (uint64_t)gridDim.x*blockDim.x)

template<int PRECOND, bool FIRST>
__global__ void __launch_bounds__(AMIE_VEC_THREADS) k_cg_dir(VecArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    double sum[1] = {0.} ;
    const double beta = FIRST ? 0. : a.st->beta ;
    VEC_LOOP(i, a)
    {
        const double ri = a.r[i] ;
        double zi ;
        if(PRECOND == PRECOND_JACOBI) zi = __dmul_rn(ri, a.d[i]) ;
        else zi = FIRST ? ri : a.z[i] ;               // NullPreconditionner: z keeps its restart value (z = r, :183)
        if(FIRST)
        {
            a.z[i] = zi ;
            a.p[i] = zi ;
            sum[0] = fma(ri, zi, sum[0]) ;
        }
        else
            a.p[i] = __dadd_rn(__dmul_rn(a.p[i], beta), zi) ;
    }
    if(FIRST)
    {
        double tot[1] ;
        if(grid_sum<1, AMIE_VEC_THREADS>(sum, a.partials, a.st->ticket+TICKET_DIR, tot) && threadIdx.x == 0)
            krylov_finalize(a.st, a.finalize, tot[0], 0.) ;
    }
}

template<int PRECOND>
__global__ void __launch_bounds__(AMIE_VEC_THREADS) k_cg_update(VecArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    const double alpha = a.st->alpha ;
    double sum[1] = {0.} ;
    VEC_LOOP(i, a)
    {
        const double ri = a.r[i], xi = a.x[i] ;
        const double yr = __dsub_rn(__dmul_rn(-a.q[i], alpha), a.rc[i]) ;
        const double yx = __dsub_rn(__dmul_rn( a.p[i], alpha), a.xc[i]) ;
        const double rtot = __dadd_rn(ri, yr) ;
        const double xtot = __dadd_rn(xi, yx) ;
        a.rc[i] = __dsub_rn(__dsub_rn(rtot, ri), yr) ;
        a.xc[i] = __dsub_rn(__dsub_rn(xtot, xi), yx) ;
        a.r[i] = rtot ;
        a.x[i] = xtot ;
        const double zi = (PRECOND == PRECOND_JACOBI) ? __dmul_rn(rtot, a.d[i]) : a.z[i] ;
        sum[0] = fma(rtot, zi, sum[0]) ;
    }
    double tot[1] ;
    if(grid_sum<1, AMIE_VEC_THREADS>(sum, a.partials, a.st->ticket+TICKET_UPDATE, tot) && threadIdx.x == 0)
        krylov_finalize(a.st, a.finalize, tot[0], 0.) ;
}

// Jacobi-Richardson smoothing sweep given r = A x - b
template<int PRECOND>
__global__ void __launch_bounds__(AMIE_VEC_THREADS) k_smooth(VecArgs a)
{
    double sum[1] = {0.} ;
    VEC_LOOP(i, a)
    {
        double ri = a.r[i] ;
        if(PRECOND == PRECOND_JACOBI)
        {
            ri = __dmul_rn(ri, a.d[i]) ;
            a.r[i] = ri ;
        }
        const double xi = a.x[i] ;
        const double yx = __dsub_rn(-ri, a.xc[i]) ;
        const double xtot = __dadd_rn(xi, yx) ;
        a.xc[i] = __dsub_rn(__dsub_rn(xtot, xi), yx) ;
        a.x[i] = xtot ;
        sum[0] = fma(ri, ri, sum[0]) ;
    }
    double tot[1] ;
    if(grid_sum<1, AMIE_VEC_THREADS>(sum, a.partials, a.st->ticket+TICKET_MISC, tot) && threadIdx.x == 0)
        krylov_finalize(a.st, a.finalize, tot[0], 0.) ;
}

template<int PRECOND>
__global__ void __launch_bounds__(AMIE_VEC_THREADS) k_bicg_p(VecArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    const double beta = a.st->beta, omega = a.st->omega ;
    VEC_LOOP(i, a)
    {
        const double pi = __dadd_rn(a.r[i], __dmul_rn(__dsub_rn(a.p[i], __dmul_rn(a.v[i], omega)), beta)) ;
        a.p[i] = pi ;
        if(PRECOND == PRECOND_JACOBI) a.p_[i] = __dmul_rn(pi, a.d[i]) ;
    }
}

template<int PRECOND>
__global__ void __launch_bounds__(AMIE_VEC_THREADS) k_bicg_s(VecArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    const double alpha = a.st->alpha ;
    VEC_LOOP(i, a)
    {
        const double si = __dsub_rn(a.r[i], __dmul_rn(a.v[i], alpha)) ;
        a.s[i] = si ;
        if(PRECOND == PRECOND_JACOBI) a.s_[i] = __dmul_rn(si, a.d[i]) ;
    }
}

static __global__ void __launch_bounds__(AMIE_VEC_THREADS) k_bicg_xr(VecArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    const double alpha = a.st->alpha, omega = a.st->omega ;
    double sum[2] = {0., 0.} ;
    VEC_LOOP(i, a)
    {
        a.x[i] = __dadd_rn(a.x[i], __dadd_rn(__dmul_rn(a.p_[i], alpha), __dmul_rn(a.s_[i], omega))) ;
        const double ri = __dsub_rn(a.s[i], __dmul_rn(a.t[i], omega)) ;
        a.r[i] = ri ;
        sum[0] = fma(ri, a.r_[i], sum[0]) ;
        sum[1] = fma(ri, ri, sum[1]) ;
    }
    double tot[2] ;
    if(grid_sum<2, AMIE_VEC_THREADS>(sum, a.partials, a.st->ticket+TICKET_UPDATE, tot) && threadIdx.x == 0)
    {
        if(a.finalize == FIN_STORE || a.finalize == FIN_DEFER_SET)
            krylov_finalize(a.st, a.finalize, tot[0], tot[1]) ;
        else
        {
            a.st->dot[1] = tot[1] ;
            krylov_finalize(a.st, a.finalize, tot[0], tot[1]) ;
        }
    }
}

// generic fused dot products on [begin,end): sums u.v and (optionally) u.w
static __global__ void __launch_bounds__(AMIE_VEC_THREADS) k_dot2(const double * u, const double * v, const double * w,
                                                          uint64_t begin, uint64_t end, KrylovState * st, double * partials, int finalize)
{
    double sum[2] = {0., 0.} ;
    for(uint64_t i = begin+(uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < end ; i += (uint64_t)gridDim.x*blockDim.x)
    {
        const double ui = u[i] ;
        sum[0] = fma(ui, v[i], sum[0]) ;
        if(w) sum[1] = fma(ui, w[i], sum[1]) ;
    }
    double tot[2] ;
    if(grid_sum<2, AMIE_VEC_THREADS>(sum, partials, st->ticket+TICKET_MISC, tot) && threadIdx.x == 0)
        krylov_finalize(st, finalize, tot[0], tot[1]) ;
}


// out = D^-1 in  (InverseDiagonal::precondition, solvers/inversediagonal.cpp:62-67)
static __global__ void __launch_bounds__(AMIE_VEC_THREADS) k_precond(const double * in, const double * d, double * out, uint64_t n)
{
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < n ; i += (uint64_t)gridDim.x*blockDim.x)
        out[i] = __dmul_rn(in[i], d[i]) ;
}

// per-block max of |v| (mode 0) or of v (mode 1) -> partials[blockIdx.x]; host finishes (<= a few thousand values)
static __global__ void __launch_bounds__(AMIE_VEC_THREADS) k_max(const double * v, uint64_t n, int mode, double * partials)
{
    __shared__ double sm[AMIE_VEC_THREADS/32] ;
    double m = -INFINITY ;
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < n ; i += (uint64_t)gridDim.x*blockDim.x)
    {
        const double x = mode == 0 ? fabs(v[i]) : v[i] ;
        m = x > m ? x : m ;
        if(x != x) m = x ;        // keep NaN visible
    }
    #pragma unroll
    for(int o = 16 ; o > 0 ; o >>= 1)
    {
        const double y = __shfl_xor_sync(0xffffffffu, m, o) ;
        m = (y > m || y != y) ? y : m ;
    }
    if((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m ;
    __syncthreads() ;
    if(threadIdx.x == 0)
    {
        for(int k = 1 ; k < AMIE_VEC_THREADS/32 ; k++) m = (sm[k] > m || sm[k] != sm[k]) ? sm[k] : m ;
        partials[blockIdx.x] = m ;
    }
}
