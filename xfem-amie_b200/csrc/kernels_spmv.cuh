// kernels_spmv.cuh -- FP64 block-row SpMV for stride-3 / stride-2 block-CSR on sm_100a.
//
// Replaces  assign(ret, A*v [- b], rowstart, colstart)  +  inner_product
//   (sparse/sparse_matrix.cpp:462-547, sparse/sparse_matrix.h:203-333)
// and the dot product that always follows it in the Krylov loops
//   (solvers/conjugategradient.cpp:234, biconjugategradientstabilized.cpp:100,115).
//
// Device layout (built once by K-Repack from the reference's padded array):
//   vals  : compact column-major blocks, block k at vals[k*S*S], element (r,c) at + c*S + r
//   col   : uint32 block-column index per block, ascending inside a row
//   rowptr: uint32 exclusive prefix of row_size (nb+1)
//
// Mapping (bandwidth-bound, no tensor cores: 18 flop per 76 B):
//   * one group of G lanes per block row; a row's values are ONE contiguous run of
//     S*S*len doubles, which the lanes read as a flat coalesced stream;
//   * S=3: 27 of the 32 lanes are active, lane l <-> (block slot l/9, element l%9), so the
//     element's (r,c) is a per-lane constant: no index arithmetic in the loop, one DFMA per
//     loaded value; 3 blocks (216 contiguous bytes) per warp-wide load;
//   * S=2: all G lanes active, lane l <-> (slot l/4, element l%4), G/4 blocks per load;
//   * the row's column indices are fetched with ONE coalesced load and handed round by
//     warp shuffles (no per-block index load);
//   * x is gathered through L1/L2 (neighbouring rows share 2/3 of their columns);
//     values are streamed with ld.global.nc.L1::no_allocate so they do not evict x;
//   * persistent grid (one wave), thread blocks sweep row tiles in a grid-stride front so
//     concurrently processed rows are adjacent (x window stays in L2);
//   * the dot product that follows is fused: per-row results are multiplied in registers
//     and reduced deterministically (grid_sum), and the loop-control scalars are updated by
//     the last block.
#pragma once
#include "common.cuh"
#include "krylov_scalars.cuh"
#include "device_utils.cuh"

#include "spmv_args.h"

// ---------------------------------------------------------------- stride 3 (27 active lanes of 32)
template<int UMAX>
__device__ __forceinline__ double s3_chunk(const double * __restrict__ vp, const double * __restrict__ x,
                                           uint32_t colreg, int nblk, int lane, int slot, int cc, double acc)
{
    double v[UMAX] ;
    uint32_t c[UMAX] ;
    bool ok[UMAX] ;
    #pragma unroll
    for(int u = 0 ; u < UMAX ; u++)
    {
        ok[u] = (lane < 27) && (3*u+slot < nblk) ;
        v[u] = ok[u] ? ld_stream(vp+u*27+lane) : 0. ;
    }
    #pragma unroll
    for(int u = 0 ; u < UMAX ; u++)
        c[u] = __shfl_sync(0xffffffffu, colreg, (3*u+slot) & 31) ;
    #pragma unroll
    for(int u = 0 ; u < UMAX ; u++)
    {
        double xv = ok[u] ? __ldg(x+(size_t)c[u]*3+cc) : 0. ;
        acc = fma(v[u], xv, acc) ;
    }
    return acc ;
}

template<int DOT, bool MINUS_B>
__global__ void __launch_bounds__(256) k_spmv_s3(SpmvArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    const int lane = threadIdx.x & 31 ;
    const int wid = threadIdx.x >> 5 ;
    const int slot = lane/9 ;                 // block slot inside a 3-block load (3 for idle lanes)
    const int e = lane-slot*9 ;
    const int cc = e/3 ;                      // column inside the block (constant per lane)
    const uint32_t ntiles = (a.nrows+7u) >> 3 ;
    double dsum[2] = {0., 0.} ;

    for(uint32_t tile = blockIdx.x ; tile < ntiles ; tile += gridDim.x)
    {
        const uint32_t lr = tile*8u+wid ;
        if(lr >= a.nrows) continue ;
        const uint32_t row = a.row0+lr ;
        uint32_t kk = __ldg(a.rowptr+row+(lane & 1)) ;
        uint32_t k0 = __shfl_sync(0xffffffffu, kk, 0) ;
        const uint32_t k1 = __shfl_sync(0xffffffffu, kk, 1) ;
        if(a.colstart_blk) k0 = row_lower_bound(a.col, k0, k1, a.colstart_blk) ;
        double acc = 0. ;
        for(uint32_t kb = k0 ; kb < k1 ; kb += 27u)
        {
            const int nblk = (int)min(27u, k1-kb) ;
            const uint32_t colreg = lane < nblk ? __ldg(a.col+kb+lane) : 0u ;
            const double * vp = a.vals+(size_t)kb*9 ;
            if(nblk > 18)      acc = s3_chunk<9>(vp, a.x, colreg, nblk, lane, slot, cc, acc) ;
            else if(nblk > 9)  acc = s3_chunk<6>(vp, a.x, colreg, nblk, lane, slot, cc, acc) ;
            else               acc = s3_chunk<3>(vp, a.x, colreg, nblk, lane, slot, cc, acc) ;
        }
        // lanes with equal lane%3 hold partial sums of the same row component
        acc += __shfl_down_sync(0xffffffffu, acc, 9) + __shfl_down_sync(0xffffffffu, acc, 18) ;   // valid in lanes 0..8
        acc += __shfl_down_sync(0xffffffffu, acc, 3) + __shfl_down_sync(0xffffffffu, acc, 6) ;    // valid in lanes 0..2
        if(lane < 3)
        {
            const size_t i = (size_t)row*3+lane ;
            double yv = acc ;
            if(MINUS_B) yv -= a.b[i] ;
            yv *= a.sign ;
            a.y[i] = yv ;
            if(DOT == DOT_YX) dsum[0] = fma(yv, a.x[i], dsum[0]) ;
            if(DOT == DOT_YY) dsum[0] = fma(yv, yv, dsum[0]) ;
            if(DOT == DOT_YW) dsum[0] = fma(yv, a.w[i], dsum[0]) ;
            if(DOT == DOT_OMEGA)
            {
                const double di = a.d ? a.d[i] : 1. ;
                const double t2 = yv*di, s2 = a.w[i]*di ;
                dsum[0] = fma(t2, s2, dsum[0]) ;
                dsum[1] = fma(t2, t2, dsum[1]) ;
            }
        }
    }
    if(DOT != DOT_NONE)
    {
        double tot[2] ;
        if(grid_sum<2, 256>(dsum, a.partials, a.st->ticket+TICKET_SPMV, tot) && threadIdx.x == 0)
            krylov_finalize(a.st, a.finalize, tot[0], tot[1]) ;
    }
}

// ---------------------------------------------------------------- stride 2 (G lanes per row, G/4 blocks per load)
template<int G, int DOT, bool MINUS_B>
__global__ void __launch_bounds__(256) k_spmv_s2(SpmvArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    constexpr int GPB = 256/G ;               // rows per tile
    constexpr int BPL = G/4 ;                 // blocks per group-wide load
    const int lane = threadIdx.x & 31 ;
    const int gl = threadIdx.x & (G-1) ;      // lane inside the group
    const int grp = threadIdx.x/G ;
    const int gbase = lane & ~(G-1) ;         // first warp lane of this group
    const int slot = gl >> 2 ;
    const int cc = (gl >> 1) & 1 ;
    const uint32_t ntiles = (a.nrows+GPB-1)/GPB ;
    double dsum[2] = {0., 0.} ;

    for(uint32_t tile = blockIdx.x ; tile < ntiles ; tile += gridDim.x)
    {
        const uint32_t lr = tile*GPB+grp ;
        const bool live = lr < a.nrows ;
        const uint32_t row = a.row0+(live ? lr : 0u) ;
        uint32_t k0 = 0, k1 = 0 ;
        if(live)
        {
            k0 = __ldg(a.rowptr+row) ;
            k1 = __ldg(a.rowptr+row+1) ;
            if(a.colstart_blk) k0 = row_lower_bound(a.col, k0, k1, a.colstart_blk) ;
        }
        // groups of one warp may have different lengths: iterate to the warp-wide maximum so
        // the shuffles stay convergent
        uint32_t len = k1-k0 ;
        uint32_t maxlen = len ;
        #pragma unroll
        for(int o = 16 ; o >= G ; o >>= 1)
            maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o)) ;
        double acc = 0. ;
        for(uint32_t cb = 0 ; cb < maxlen ; cb += G)
        {
            // G column indices per group per chunk
            const uint32_t colreg = (cb+gl < len) ? __ldg(a.col+k0+cb+gl) : 0u ;
            #pragma unroll
            for(int u = 0 ; u < 4 ; u++)
            {
                const uint32_t bi = cb+u*BPL+slot ;
                const bool ok = bi < len ;
                const double v = ok ? ld_stream(a.vals+((size_t)(k0+bi) << 2)+(gl & 3)) : 0. ;
                const uint32_t c = __shfl_sync(0xffffffffu, colreg, gbase+((u*BPL+slot) & (G-1))) ;
                const double xv = ok ? __ldg(a.x+(size_t)c*2+cc) : 0. ;
                acc = fma(v, xv, acc) ;
            }
        }
        // lanes of a group with equal parity hold the same row component
        #pragma unroll
        for(int o = G/2 ; o >= 2 ; o >>= 1)
            acc += __shfl_xor_sync(0xffffffffu, acc, o) ;
        if(live && gl < 2)
        {
            const size_t i = (size_t)row*2+gl ;
            double yv = acc ;
            if(MINUS_B) yv -= a.b[i] ;
            yv *= a.sign ;
            a.y[i] = yv ;
            if(DOT == DOT_YX) dsum[0] = fma(yv, a.x[i], dsum[0]) ;
            if(DOT == DOT_YY) dsum[0] = fma(yv, yv, dsum[0]) ;
            if(DOT == DOT_YW) dsum[0] = fma(yv, a.w[i], dsum[0]) ;
            if(DOT == DOT_OMEGA)
            {
                const double di = a.d ? a.d[i] : 1. ;
                const double t2 = yv*di, s2 = a.w[i]*di ;
                dsum[0] = fma(t2, s2, dsum[0]) ;
                dsum[1] = fma(t2, t2, dsum[1]) ;
            }
        }
    }
    if(DOT != DOT_NONE)
    {
        double tot[2] ;
        if(grid_sum<2, 256>(dsum, a.partials, a.st->ticket+TICKET_SPMV, tot) && threadIdx.x == 0)
            krylov_finalize(a.st, a.finalize, tot[0], tot[1]) ;
    }
}

// ---------------------------------------------------------------- other strides (1, 4, 6): one thread per scalar row
// The reference's inner_product also has stride 1, 4, 6 and generic cases (sparse/sparse_matrix.h:222-233,
// :335-676): diffusion problems (1 DOF per node) and space-time elements (4, 6).  They are off the 2D/3D
// elasticity headline path; this kernel keeps the drop-in usable for them (correct, persistent grid, fused dot,
// no staging).
template<int S, int DOT, bool MINUS_B>
__global__ void __launch_bounds__(256) k_spmv_gen(SpmvArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    double dsum[2] = {0., 0.} ;
    const uint64_t n = (uint64_t)a.nrows*S ;
    for(uint64_t t = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; t < n ; t += (uint64_t)gridDim.x*blockDim.x)
    {
        const uint32_t row = a.row0+(uint32_t)(t/S) ;
        const int r = (int)(t%S) ;
        uint32_t k0 = __ldg(a.rowptr+row) ;
        const uint32_t k1 = __ldg(a.rowptr+row+1) ;
        if(a.colstart_blk) k0 = row_lower_bound(a.col, k0, k1, a.colstart_blk) ;
        double acc = 0. ;
        for(uint32_t k = k0 ; k < k1 ; k++)
        {
            const double * v = a.vals+(size_t)k*S*S+r ;
            const double * px = a.x+(size_t)__ldg(a.col+k)*S ;
            #pragma unroll
            for(int c = 0 ; c < S ; c++)
                acc = fma(ld_stream(v+c*S), __ldg(px+c), acc) ;
        }
        const size_t i = (size_t)row*S+r ;
        double yv = acc ;
        if(MINUS_B) yv -= a.b[i] ;
        yv *= a.sign ;
        a.y[i] = yv ;
        if(DOT == DOT_YX) dsum[0] = fma(yv, a.x[i], dsum[0]) ;
        if(DOT == DOT_YY) dsum[0] = fma(yv, yv, dsum[0]) ;
        if(DOT == DOT_YW) dsum[0] = fma(yv, a.w[i], dsum[0]) ;
        if(DOT == DOT_OMEGA)
        {
            const double di = a.d ? a.d[i] : 1. ;
            const double t2 = yv*di, s2 = a.w[i]*di ;
            dsum[0] = fma(t2, s2, dsum[0]) ;
            dsum[1] = fma(t2, t2, dsum[1]) ;
        }
    }
    if(DOT != DOT_NONE)
    {
        double tot[2] ;
        if(grid_sum<2, 256>(dsum, a.partials, a.st->ticket+TICKET_SPMV, tot) && threadIdx.x == 0)
            krylov_finalize(a.st, a.finalize, tot[0], tot[1]) ;
    }
}
