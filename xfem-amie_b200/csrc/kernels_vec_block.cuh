// kernels_vec_block.cuh -- the fused CG kernels of kernels_vec.cuh for BLOCK preconditioners.
//
//   precond_kind 5 : Inverse2x2Diagonal (solvers/inversediagonal.cpp:84-133) on stride-2 systems -- the reference's own
//                    class: blocks built with det() / invert2x2Matrix (utilities/matrixops.cpp:834-837, :536-561),
//                    applied as t = B v with every product and sum rounded separately;
//   precond_kind 6 : the same construction on the 3x3 node blocks of stride-3 systems, with the reference's det() and
//                    invert3x3Matrix (:838-847, :681-702).  No reference class: the opt-in block-Jacobi of the survey.
//
// One THREAD per node: it owns the node's S entries of every vector, so r can be updated in place and z = B r formed
// from the new values without reading a neighbour's half-written entries.  A warp still touches contiguous bytes (the
// S loads of a thread interleave with its neighbours'), each line is fetched once.
// precondition(r, r) -- the smoothing sweeps pass the same vector twice (conjugategradient.cpp:141, :282) -- lets row
// m of a block read the rows < m already overwritten, exactly as the reference's loop does; kept (k_smooth_blk).
#pragma once
#include "kernels_vec.cuh"
#include "device_utils.cuh"

enum { PRECOND_BLOCK = 2 } ;

template<int S>
__device__ __forceinline__ void blk_apply(const double * __restrict__ B, const double (&v)[S], double (&t)[S])
{
    #pragma unroll
    for(int r = 0 ; r < S ; r++)
    {
        double acc = __dadd_rn(__dmul_rn(B[r*S], v[0]), __dmul_rn(B[r*S+1], v[1])) ;
        if(S == 3) acc = __dadd_rn(acc, __dmul_rn(B[r*S+2], v[2])) ;
        t[r] = acc ;
    }
}

// the blocks: one thread per node; the diagonal block is found by a linear scan (a partitioned context's column
// numbering is not ascending)
template<int S>
static __global__ void k_block_inverse(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ col,
                                       const double * __restrict__ vals, uint64_t nb, double * __restrict__ blocks)
{
    for(uint64_t k = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; k < nb ; k += (uint64_t)gridDim.x*blockDim.x)
    {
        double m[S*S] ;                       // row-major m[r*S+c] = A(kS+r, kS+c); stored blocks are column-major
        #pragma unroll
        for(int i = 0 ; i < S*S ; i++) m[i] = 0. ;
        for(uint32_t l = rowptr[k] ; l < rowptr[k+1] ; l++)
            if(col[l] == (uint32_t)k)
            {
                #pragma unroll
                for(int r = 0 ; r < S ; r++)
                    #pragma unroll
                    for(int c = 0 ; c < S ; c++) m[r*S+c] = vals[(size_t)l*S*S+c*S+r] ;
                break ;
            }
        double * B = blocks+k*(S*S) ;
        if(S == 2)
        {
            const double a = m[0], b = m[1], c = m[2], d = m[3] ;
            const double dt = fma(a, d, -__dmul_rn(b, c)) ;
            if(fabs(dt) > 1e-8)
            {
                if(fabs(a) < 1e-24)
                {
                    const double inv = __ddiv_rn(1., __dsub_rn(__dmul_rn(a, d), __dmul_rn(b, c))) ;
                    B[0] = __dmul_rn(d, inv) ; B[1] = __dmul_rn(-b, inv) ; B[2] = __dmul_rn(-c, inv) ; B[3] = __dmul_rn(a, inv) ;
                }
                else
                {
                    const double r1 = __ddiv_rn(1., a) ;
                    const double r3 = __dmul_rn(r1, b) ;
                    const double r6 = __ddiv_rn(1., __dsub_rn(__dmul_rn(c, r3), d)) ;
                    const double b2 = __dmul_rn(__dmul_rn(r6, c), r1) ;
                    B[1] = __dmul_rn(r3, r6) ;
                    B[2] = b2 ;
                    B[0] = __dsub_rn(r1, __dmul_rn(r3, b2)) ;
                    B[3] = -r6 ;
                }
            }
            else
            {
                B[1] = 0. ; B[2] = 0. ;
                B[0] = fabs(a) > 1e-8 ? __ddiv_rn(1., a) : 1. ;
                B[3] = fabs(d) > 1e-8 ? __ddiv_rn(1., d) : 1. ;
            }
        }
        else
        {
            const double dt = fma(m[0], __dmul_rn(m[4], m[8]), fma(m[5], __dmul_rn(m[6], m[1]), fma(m[2], __dmul_rn(m[7], m[3]),
                              fma(-m[0], __dmul_rn(m[5], m[7]), fma(-m[1], __dmul_rn(m[3], m[8]), __dmul_rn(__dmul_rn(-m[6], m[4]), m[2])))))) ;
            if(fabs(dt) > 1e-8)
            {
                const double r11 = fma(m[4], m[8], -__dmul_rn(m[5], m[7])) ;
                const double r21 = fma(m[5], m[6], -__dmul_rn(m[3], m[8])) ;
                const double r31 = fma(m[3], m[7], -__dmul_rn(m[4], m[6])) ;
                const double inv = __ddiv_rn(1., fma(m[0], r11, fma(m[1], r21, __dmul_rn(m[2], r31)))) ;
                const double v[9] = { r11, fma(m[2], m[7], -__dmul_rn(m[1], m[8])), fma(m[1], m[5], -__dmul_rn(m[2], m[4])),
                                      r21, fma(m[0], m[8], -__dmul_rn(m[2], m[6])), fma(m[2], m[3], -__dmul_rn(m[0], m[5])),
                                      r31, fma(m[1], m[6], -__dmul_rn(m[0], m[7])), fma(m[4], m[0], -__dmul_rn(m[1], m[3])) } ;
                #pragma unroll
                for(int i = 0 ; i < 9 ; i++) B[i] = __dmul_rn(v[i], inv) ;
            }
            else
            {
                #pragma unroll
                for(int i = 0 ; i < S*S ; i++) B[i] = 0. ;
                #pragma unroll
                for(int i = 0 ; i < S ; i++) B[i*S+i] = fabs(m[i*S+i]) > 1e-8 ? __ddiv_rn(1., m[i*S+i]) : 1. ;
            }
        }
    }
}

#define BLK_LOOP(k, a, S) for(uint64_t k = (a).begin/(S)+(uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; k < (a).end/(S) ; k += (uint64_t)gridDim.x*blockDim.x)

// z = B r ; p = beta p + z   (restart form: p = z ; rho0 = r.z)
template<int S, bool FIRST>
__global__ void __launch_bounds__(AMIE_VEC_THREADS) k_cg_dir_blk(VecArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    double sum[1] = {0.} ;
    const double beta = FIRST ? 0. : a.st->beta ;
    BLK_LOOP(k, a, S)
    {
        double rv[S], zv[S] ;
        #pragma unroll
        for(int m = 0 ; m < S ; m++) rv[m] = a.r[k*S+m] ;
        blk_apply<S>(a.d+k*(S*S), rv, zv) ;
        #pragma unroll
        for(int m = 0 ; m < S ; m++)
        {
            const uint64_t i = k*S+m ;
            if(FIRST)
            {
                a.z[i] = zv[m] ;
                a.p[i] = zv[m] ;
                sum[0] = fma(rv[m], zv[m], sum[0]) ;
            }
            else
                a.p[i] = __dadd_rn(__dmul_rn(a.p[i], beta), zv[m]) ;
        }
    }
    if(FIRST)
    {
        double tot[1] ;
        if(grid_sum<1, AMIE_VEC_THREADS>(sum, a.partials, a.st->ticket+TICKET_DIR, tot) && threadIdx.x == 0)
            krylov_finalize(a.st, a.finalize, tot[0], 0.) ;
    }
}

// x += alpha p ; r -= alpha q (Kahan) ; rho' = r.(B r)
template<int S>
__global__ void __launch_bounds__(AMIE_VEC_THREADS) k_cg_update_blk(VecArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    const double alpha = a.st->alpha ;
    double sum[1] = {0.} ;
    BLK_LOOP(k, a, S)
    {
        double rn[S], zv[S] ;
        #pragma unroll
        for(int m = 0 ; m < S ; m++)
        {
            const uint64_t i = k*S+m ;
            const double ri = a.r[i], xi = a.x[i] ;
            const double yr = __dsub_rn(__dmul_rn(-a.q[i], alpha), a.rc[i]) ;
            const double yx = __dsub_rn(__dmul_rn( a.p[i], alpha), a.xc[i]) ;
            const double rtot = __dadd_rn(ri, yr) ;
            const double xtot = __dadd_rn(xi, yx) ;
            a.rc[i] = __dsub_rn(__dsub_rn(rtot, ri), yr) ;
            a.xc[i] = __dsub_rn(__dsub_rn(xtot, xi), yx) ;
            a.r[i] = rtot ;
            a.x[i] = xtot ;
            rn[m] = rtot ;
        }
        blk_apply<S>(a.d+k*(S*S), rn, zv) ;
        #pragma unroll
        for(int m = 0 ; m < S ; m++) sum[0] = fma(rn[m], zv[m], sum[0]) ;
    }
    double tot[1] ;
    if(grid_sum<1, AMIE_VEC_THREADS>(sum, a.partials, a.st->ticket+TICKET_UPDATE, tot) && threadIdx.x == 0)
        krylov_finalize(a.st, a.finalize, tot[0], 0.) ;
}

// smoothing sweep given r = A x - b: precondition(r, r) IN PLACE (row m of a block reads the rows < m already
// replaced, as the reference's loop over the aliased vectors does), x -= r (Kahan), |r|^2
template<int S>
__global__ void __launch_bounds__(AMIE_VEC_THREADS) k_smooth_blk(VecArgs a)
{
    double sum[1] = {0.} ;
    BLK_LOOP(k, a, S)
    {
        const double * B = a.d+k*(S*S) ;
        double v[S] ;
        #pragma unroll
        for(int m = 0 ; m < S ; m++) v[m] = a.r[k*S+m] ;
        #pragma unroll
        for(int m = 0 ; m < S ; m++)
        {
            double acc = __dadd_rn(__dmul_rn(B[m*S], v[0]), __dmul_rn(B[m*S+1], v[1])) ;
            if(S == 3) acc = __dadd_rn(acc, __dmul_rn(B[m*S+2], v[2])) ;
            v[m] = acc ;                                  // later rows see it
        }
        #pragma unroll
        for(int m = 0 ; m < S ; m++)
        {
            const uint64_t i = k*S+m ;
            const double ri = v[m] ;
            a.r[i] = ri ;
            const double xi = a.x[i] ;
            const double yx = __dsub_rn(-ri, a.xc[i]) ;
            const double xtot = __dadd_rn(xi, yx) ;
            a.xc[i] = __dsub_rn(__dsub_rn(xtot, xi), yx) ;
            a.x[i] = xtot ;
            sum[0] = fma(ri, ri, sum[0]) ;
        }
    }
    double tot[1] ;
    if(grid_sum<1, AMIE_VEC_THREADS>(sum, a.partials, a.st->ticket+TICKET_MISC, tot) && threadIdx.x == 0)
        krylov_finalize(a.st, a.finalize, tot[0], 0.) ;
}
