// device_utils.cuh -- two device helpers shared by every kernel file.  No other dependency than <stdint.h>, so the
// simple kernels that include only this header can also be compiled for the host by the test-only emulation
// (tests/emu/, which supplies its own definitions when AMIE_B200_EMU is set).
#pragma once
#include <stdint.h>

#ifndef AMIE_B200_EMU
__device__ __forceinline__ double ld_stream(const double * p)
{
    double v ;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)) ;
    return v ;
}

__device__ __forceinline__ uint32_t row_lower_bound(const uint32_t * col, uint32_t k0, uint32_t k1, uint32_t key)
{
    while(k0 < k1)
    {
        uint32_t mid = k0+((k1-k0) >> 1) ;
        if(__ldg(col+mid) < key) k0 = mid+1 ; else k1 = mid ;
    }
    return k0 ;
}
#endif
