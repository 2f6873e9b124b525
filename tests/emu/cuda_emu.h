// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  Lets g++ compile the dependency-light kernel headers of
// xfem-amie_b200/csrc (kernels_setup.cuh, kernels_assemble.cuh, kernels_fields.cuh, kernels_history.cuh) for the host,
// so that `pytest -m "not gpu"` exercises the index logic and the arithmetic order of the very kernel sources the
// GPU runs, against the oracle, on a box without a GPU.  This is not a CPU path of the product: nothing under
// xfem-amie_b200/ includes it, the product library is built by nvcc for sm_100a only and fails without a device.
//
// Model: a launch runs its blocks one after the other; the threads of a block run one after the other too, unless the
// kernel uses __syncthreads(), in which case every thread of the block is a std::thread and the barrier is real.
// Not modelled (the kernels compiled here do not use them): warp shuffles, cp.async / TMA / mbarrier, tensor memory.
#pragma once
#define AMIE_B200_EMU 1
#include <stdint.h>
#include <math.h>
#include <algorithm>
#include <barrier>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

struct EmuDim3 { unsigned x = 1, y = 1, z = 1 ; } ;
inline thread_local EmuDim3 threadIdx, blockIdx ;
inline EmuDim3 blockDim, gridDim ;
inline thread_local std::barrier<> * emu_block_barrier = nullptr ;

// round-to-nearest, no contraction: this file is compiled with -ffp-contract=off
inline double __dmul_rn(double a, double b) { return a*b ; }
inline double __dadd_rn(double a, double b) { return a+b ; }
inline double __dsub_rn(double a, double b) { return a-b ; }
template<typename T> inline T __ldg(const T * p) { return *p ; }
inline double ld_stream(const double * p) { return *p ; }
inline unsigned int atomicAdd(unsigned int * p, unsigned int v) { unsigned int o = *p ; *p += v ; return o ; }   // sequential model
template<typename T> inline T min(T a, T b) { return a < b ? a : b ; }
template<typename T> inline T max(T a, T b) { return a > b ? a : b ; }
inline void __syncthreads() { emu_block_barrier->arrive_and_wait() ; }

inline uint32_t row_lower_bound(const uint32_t * col, uint32_t k0, uint32_t k1, uint32_t key)
{
    while(k0 < k1)
    {
        uint32_t mid = k0+((k1-k0) >> 1) ;
        if(col[mid] < key) k0 = mid+1 ; else k1 = mid ;
    }
    return k0 ;
}

// kernel<<<grid, block>>>(args...) for kernels without __syncthreads()
template<typename F>
inline void emu_launch(unsigned grid, unsigned block, F && body)
{
    gridDim.x = grid ; blockDim.x = block ;
    for(unsigned b = 0 ; b < grid ; b++)
        for(unsigned t = 0 ; t < block ; t++)
        {
            blockIdx.x = b ; threadIdx.x = t ;
            body() ;
        }
}

// ... and for kernels that synchronise inside the block
template<typename F>
inline void emu_launch_sync(unsigned grid, unsigned block, F && body)
{
    gridDim.x = grid ; blockDim.x = block ;
    for(unsigned b = 0 ; b < grid ; b++)
    {
        std::barrier<> bar(block) ;
        std::vector<std::thread> th ;
        for(unsigned t = 0 ; t < block ; t++)
            th.emplace_back([&, t]() { blockIdx.x = b ; threadIdx.x = t ; emu_block_barrier = &bar ; body() ; }) ;
        for(auto & x : th) x.join() ;
    }
}
