// kernels_history.cuh -- the two kernels of Assembly::extrapolate / displacementHistory on the device (cgsolve.cu).
// Depends on device_utils.cuh only (see there).
#pragma once
#include "device_utils.cuh"
#ifndef AMIE_VEC_THREADS
#define AMIE_VEC_THREADS 256
#endif

// x = back + (back - prev)*factor + 0.5*dxxddb*factor*factor  with dxxddb == 0   (:1791-1812)
static __global__ void __launch_bounds__(AMIE_VEC_THREADS)
k_extrapolate(const double * __restrict__ prev, double * __restrict__ back, double * __restrict__ x, uint64_t n, double factor)
{
    const double second = __dmul_rn(__dmul_rn(__dmul_rn(0.5, 0.), factor), factor) ;
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < n ; i += (uint64_t)gridDim.x*blockDim.x)
    {
        double b = back[i] ;
        if(b != b) { b = 0. ; back[i] = 0. ; }                    // :1793-1794
        const double dxdb = __dsub_rn(b, prev[i]) ;               // :1795
        x[i] = __dadd_rn(__dadd_rn(b, __dmul_rn(dxdb, factor)), second) ;
    }
}

// displacementHistory.push_back(displacements*0.)   (:1866)
static __global__ void __launch_bounds__(AMIE_VEC_THREADS)
k_times_zero(const double * __restrict__ x, double * __restrict__ out, uint64_t n)
{
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < n ; i += (uint64_t)gridDim.x*blockDim.x)
        out[i] = __dmul_rn(x[i], 0.) ;
}
