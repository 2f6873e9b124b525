// kernels_fields.cuh -- the kernels of the per-element field recovery (fields.cu holds the description, the host side
// and the C-ABI).  Depends on device_utils.cuh only (see there).
#pragma once
#include "device_utils.cuh"

#ifndef NO_NODE
#define NO_NODE 0xFFFFFFFFu
#endif
#define FIELD_THREADS 128

// in[e*K + k] -> out[k*n + e]  (once per topology)
template<typename T>
static __global__ void k_to_component_major(const T * __restrict__ in, T * __restrict__ out, uint64_t n, int K)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    const uint64_t total = n*(uint64_t)K ;
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < total ; i += stride)
    {
        const uint64_t e = i/K ;
        const int k = (int)(i-e*K) ;
        out[(uint64_t)k*n+e] = in[i] ;
    }
}

// One thread per element.  g[c][l] accumulates d(shape_j)/d(local_l) * u_j[c] over the element's slots in slot
// order, exactly as the reference's x_xi ... z_zeta accumulators do.
template<int DIM>
static __global__ void __launch_bounds__(FIELD_THREADS)
k_element_fields(const uint32_t * __restrict__ ids, const double * __restrict__ dshape, const double * __restrict__ jinv,
                 const double * __restrict__ tensors, const double * __restrict__ istrain, const double * __restrict__ istress,
                 const uint32_t * __restrict__ tensor_of_elem, const double * __restrict__ u, uint64_t n_u,
                 uint64_t n_elem, int npe, double * __restrict__ total_out, double * __restrict__ mech_out,
                 double * __restrict__ stress_out)
{
    constexpr int NC = DIM == 2 ? 3 : 6 ;
    __shared__ double stage[FIELD_THREADS*NC] ;
    const uint64_t ntiles = (n_elem+FIELD_THREADS-1)/FIELD_THREADS ;
    for(uint64_t tile = blockIdx.x ; tile < ntiles ; tile += gridDim.x)
    {
        const uint64_t e0 = tile*FIELD_THREADS ;
        const uint64_t e = e0+threadIdx.x ;
        const bool live = e < n_elem ;
        const int cnt = (int)min((uint64_t)FIELD_THREADS, n_elem-e0) ;
        double t[NC], m[NC], s[NC] ;
        #pragma unroll
        for(int i = 0 ; i < NC ; i++) { t[i] = 0. ; m[i] = 0. ; s[i] = 0. ; }
        if(live)
        {
            double g[DIM][DIM] ;
            #pragma unroll
            for(int c = 0 ; c < DIM ; c++)
                #pragma unroll
                for(int l = 0 ; l < DIM ; l++) g[c][l] = 0. ;
            for(int j = 0 ; j < npe ; j++)
            {
                const uint32_t id = __ldg(ids+(uint64_t)j*n_elem+e) ;
                if(id == NO_NODE) continue ;
                double f[DIM] ;
                #pragma unroll
                for(int l = 0 ; l < DIM ; l++) f[l] = ld_stream(dshape+(uint64_t)(j*DIM+l)*n_elem+e) ;
                #pragma unroll
                for(int c = 0 ; c < DIM ; c++)
                {
                    const uint64_t k = (uint64_t)id*DIM+c ;
                    const double d = k < n_u ? __ldg(u+k) : 0. ;          // ElementState::step, :3641-3648
                    #pragma unroll
                    for(int l = 0 ; l < DIM ; l++) g[c][l] = __dadd_rn(g[c][l], __dmul_rn(f[l], d)) ;
                }
            }
            double J[DIM*DIM] ;
            #pragma unroll
            for(int k = 0 ; k < DIM*DIM ; k++) J[k] = ld_stream(jinv+(uint64_t)k*n_elem+e) ;
            // a[0]*J[r][0] + a[1]*J[r][1] (+ a[2]*J[r][2]), left to right
            auto row = [&](const double * a, int r)
            {
                double v = __dadd_rn(__dmul_rn(a[0], J[r*DIM]), __dmul_rn(a[1], J[r*DIM+1])) ;
                if constexpr(DIM == 3) v = __dadd_rn(v, __dmul_rn(a[2], J[r*DIM+2])) ;
                return v ;
            } ;
            // ... continued with b[0]*J[q][0] + b[1]*J[q][1] (+ b[2]*J[q][2])
            auto row2 = [&](const double * a, int r, const double * b, int q)
            {
                double v = row(a, r) ;
                v = __dadd_rn(v, __dmul_rn(b[0], J[q*DIM])) ;
                v = __dadd_rn(v, __dmul_rn(b[1], J[q*DIM+1])) ;
                if constexpr(DIM == 3) v = __dadd_rn(v, __dmul_rn(b[2], J[q*DIM+2])) ;
                return v ;
            } ;
            if constexpr(DIM == 2)
            {
                t[0] = row(g[0], 0) ;                         // :1015
                t[1] = row(g[1], 1) ;                         // :1016
                t[2] = row2(g[0], 1, g[1], 0) ;               // :1017
            }
            else
            {
                t[0] = row(g[0], 0) ;                         // :1072-1074
                t[1] = row(g[1], 1) ;
                t[2] = row(g[2], 2) ;
                t[3] = row2(g[1], 2, g[2], 1) ;               // :1076-1081
                t[4] = row2(g[0], 2, g[2], 0) ;               // :1083-1088
                t[5] = row2(g[1], 0, g[0], 1) ;               // :1090-1095
            }
            const uint64_t ti = tensor_of_elem ? __ldg(tensor_of_elem+e) : e ;
            #pragma unroll
            for(int i = 0 ; i < NC ; i++) m[i] = __dsub_rn(t[i], __ldg(istrain+ti*NC+i)) ;      // :967-968
            const double * C = tensors+ti*NC*NC ;
            #pragma unroll
            for(int i = 0 ; i < NC ; i++)
            {
                double acc = 0. ;
                #pragma unroll
                for(int k = 0 ; k < NC ; k++) acc = __dadd_rn(acc, __dmul_rn(__ldg(C+i*NC+k), m[k])) ;   // matrixops.h:555
                s[i] = __dsub_rn(acc, __ldg(istress+ti*NC+i)) ;                              // :1392
            }
        }
        // element-major results through shared memory: one contiguous, coalesced store per field and tile
        double * outs[3] = { total_out, mech_out, stress_out } ;
        #pragma unroll
        for(int w = 0 ; w < 3 ; w++)
        {
            __syncthreads() ;
            #pragma unroll
            for(int i = 0 ; i < NC ; i++) stage[threadIdx.x*NC+i] = w == 0 ? t[i] : (w == 1 ? m[i] : s[i]) ;
            __syncthreads() ;
            double * dst = outs[w]+e0*NC ;
            for(int i = threadIdx.x ; i < cnt*NC ; i += FIELD_THREADS) dst[i] = stage[i] ;
        }
    }
}
