// kernels_spmv_rt2.cuh -- stride-2 (2D) block-row SpMV: the row-thread pipeline of kernels_spmv_rt.cuh
// for 2x2 blocks.
//
//   tile  = 16 consecutive block rows = 32 scalar rows = the 32 lanes of one compute warp
//   block = 4 doubles (32 B, always 16-byte aligned: the TMA copies need no lead skipping)
//   x     = one 16-byte cp.async per block (x node = 2 doubles, 16-byte aligned)
//   2D rows are short (T3 meshes: ~7 blocks), so a stage is small (CAP 144 blocks = 8.6 KB) and there is
//   room for 8 compute warps, 4 producer warps and 20 stages (spmv_launch.cu): in 2D the vector traffic
//   weighs as much as the matrix (144 B/DOF of SpMV against 120 B/DOF of vector kernels), and per-row
//   fixed costs dominate.
//   lane  = (block row, block COLUMN) for the products, (block row, row component) for the store (below)
#pragma once
#include "kernels_spmv_rt.cuh"

#define RT2_ROWS 16

template<int NST, int CAP>
struct Rt2Layout
{
    static constexpr int VAL_BYTES = (CAP*32+16+15)/16*16 ;
    static constexpr int COL_BYTES = (CAP*4+16+15)/16*16 ;
    static constexpr int XS_BYTES = CAP*16 ;
    static constexpr int AUX_BYTES = 4*32*8 ;
    static constexpr int META_BYTES = 128 ;                        // 17 row pointers + 3 words
    static constexpr int STAGE_BYTES = (VAL_BYTES+COL_BYTES+XS_BYTES+AUX_BYTES+META_BYTES+127)/128*128 ;
    static constexpr int META_OFF = VAL_BYTES+COL_BYTES+XS_BYTES+AUX_BYTES ;
    static constexpr int TOTAL_BYTES = NST*STAGE_BYTES+2*NST*8+16 ;
} ;

// Column mapping: lane (rl, c) holds COLUMN c of its block row's blocks -- A(0,c), A(1,c) are 16 contiguous, 16-byte
// aligned bytes (one LDS.128) and only x_c is needed (one LDS.64): 2 shared-memory instructions / 6 wavefronts per block
// instead of 4 / 8, conflict-free for rows of 7 blocks (row stride 224 B: four block rows x 2 columns fill the 128 B of
// a quarter-warp pass; a lane per scalar row collides block rows rl and rl + 4: measured 192 M -> 150 M shared-memory
// wavefronts per launch on S2-tri-4096, profiles/r02_notes.md section 8).  The two partial sums per row are the FMA chains
// a lane per scalar row forms (acc_c over the row's blocks), exchanged once per tile: y_r = acc_r(col 0) + acc_r(col 1) --
// same bits (checked on B200 against that mapping before it was deleted).
template<int OFF>
__device__ __forceinline__ void lds_v2f64(uint32_t base, double & v0, double & v1)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v0), "=d"(v1) : "r"(base), "n"(OFF)) ;
}
template<int Q, int NB>
struct Rt2LoadCM
{
    static __device__ __forceinline__ void run(uint32_t va, uint32_t xa, double (&v0)[NB], double (&v1)[NB], double (&x)[NB])
    {
        lds_v2f64<Q*32>(va, v0[Q], v1[Q]) ;
        x[Q] = lds_f64<Q*16>(xa) ;
        Rt2LoadCM<Q+1, NB>::run(va, xa, v0, v1, x) ;
    }
} ;
template<int NB>
struct Rt2LoadCM<NB, NB>
{
    static __device__ __forceinline__ void run(uint32_t, uint32_t, double (&)[NB], double (&)[NB], double (&)[NB]) { }
} ;
// acc0 += A(0,c) x_c ; acc1 += A(1,c) x_c   (va: shared address of A(0,c) of the first block, xa: of its x_c)
template<int NB>
__device__ __forceinline__ void rt2_blocks_cm(uint32_t va, uint32_t xa, double & acc0, double & acc1)
{
    double v0[NB], v1[NB], x[NB] ;
    Rt2LoadCM<0, NB>::run(va, xa, v0, v1, x) ;
    #pragma unroll
    for(int q = 0 ; q < NB ; q++)
    {
        acc0 = fma(v0[q], x[q], acc0) ;
        acc1 = fma(v1[q], x[q], acc1) ;
    }
}

template<int DOT, bool MINUS_B, int W, int NST, int CAP, int G, int NP = 1>
__global__ void __launch_bounds__((W+NP)*32) k_spmv_s2_rt(SpmvArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    static_assert(NST >= (G+1)*W, "stages: W tiles in compute + G*W tiles being gathered") ;
    constexpr int R = RT2_ROWS ;
    using L = Rt2Layout<NST, CAP> ;
    extern __shared__ __align__(128) unsigned char smem[] ;
    uint64_t * full_v = reinterpret_cast<uint64_t *>(smem+NST*L::STAGE_BYTES) ;
    uint64_t * empty = full_v+NST ;
    const int lane = threadIdx.x & 31 ;
    const int wid = threadIdx.x >> 5 ;
    const uint32_t ntiles = (a.nrows+R-1)/R ;

    if(threadIdx.x == 0)
    {
        for(int s = 0 ; s < NST ; s++)
        {
            mbar_init(full_v+s, 1) ;
            mbar_init(empty+s, 1) ;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory") ;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory") ;
    }
    __syncthreads() ;

    double dsum[2] = {0., 0.} ;

    if(wid >= W)
    {
        tile_producer<R, NST, CAP, L::STAGE_BYTES, L::VAL_BYTES, L::META_OFF, 4, 32>(a, smem, full_v, empty, ntiles, lane, wid-W, NP) ;
    }
    else
    {
        const int rl = lane >> 1 ;              // block row inside the tile
        const int r = lane & 1 ;                // the block COLUMN this lane multiplies, and the row component it stores

        auto issue_gather = [&](uint32_t j)
        {
            const uint32_t tile = blockIdx.x+j*gridDim.x ;
            if(tile < ntiles)
            {
                const int s = j%NST ;
                unsigned char * stage = smem+s*L::STAGE_BYTES ;
                const uint32_t * meta = reinterpret_cast<const uint32_t *>(stage+L::META_OFF) ;
                const uint32_t r0 = a.row0+tile*R ;
                const uint32_t nr = min((uint32_t)R, a.row0+a.nrows-r0) ;
                double * aux = reinterpret_cast<double *>(stage+L::VAL_BYTES+L::COL_BYTES+L::XS_BYTES) ;
                mbar_wait(full_v+s, (j/NST) & 1u) ;          // also: the stage is free (see kernels_spmv_rt.cuh)
                if(rl < (int)nr)
                {
                    const size_t i = (size_t)(r0+rl)*2+r ;
                    if(MINUS_B) cp_async_8(aux+lane, a.b+i) ;
                    if(DOT == DOT_YX) cp_async_8(aux+32+lane, a.x+i) ;
                    if(DOT == DOT_YW || DOT == DOT_OMEGA) cp_async_8(aux+64+lane, a.w+i) ;
                    if(DOT == DOT_OMEGA && a.d) cp_async_8(aux+96+lane, a.d+i) ;
                }
                if(meta[R+3] != 0u)
                {
                    const uint32_t * cs = reinterpret_cast<const uint32_t *>(stage+L::VAL_BYTES+meta[R+2]) ;
                    double * xs = reinterpret_cast<double *>(stage+L::VAL_BYTES+L::COL_BYTES) ;
                    const uint32_t nblk = meta[nr]-meta[0] ;
                    constexpr int GI = (CAP+31)/32 ;
                    uint32_t cidx[GI] ;
                    #pragma unroll
                    for(int g = 0 ; g < GI ; g++)
                        cidx[g] = (lane+32u*g < nblk) ? cs[lane+32u*g] : 0u ;
                    #pragma unroll
                    for(int g = 0 ; g < GI ; g++)
                    {
                        const uint32_t bk = lane+32u*g ;
                        if(bk < nblk) cp_async_16(xs+(size_t)bk*2, a.x+(size_t)cidx[g]*2) ;
                    }
                }
            }
            cp_async_commit() ;
        } ;

        #pragma unroll
        for(int g = 0 ; g < G ; g++) issue_gather(wid+g*W) ;

        for(uint32_t j = wid ; blockIdx.x+j*gridDim.x < ntiles ; j += W)
        {
            const uint32_t tile = blockIdx.x+j*gridDim.x ;
            issue_gather(j+G*W) ;
            cp_async_wait_group<G>() ;
            __syncwarp() ;
            const int s = j%NST ;
            const unsigned char * stage = smem+s*L::STAGE_BYTES ;
            const uint32_t * meta = reinterpret_cast<const uint32_t *>(stage+L::META_OFF) ;
            const double * aux = reinterpret_cast<const double *>(stage+L::VAL_BYTES+L::COL_BYTES+L::XS_BYTES) ;
            const uint32_t r0 = a.row0+tile*R ;
            const uint32_t nr = min((uint32_t)R, a.row0+a.nrows-r0) ;
            const uint32_t k_lo = meta[0] ;
            const bool staged = meta[R+3] != 0u ;
            if(rl < (int)nr)
            {
                uint32_t k0 = meta[rl] ;
                const uint32_t k1 = meta[rl+1] ;
                double acc0 = 0., acc1 = 0. ;
                if(staged)
                {
                    const uint32_t * cs = reinterpret_cast<const uint32_t *>(stage+L::VAL_BYTES+meta[R+2]) ;
                    if(a.colstart_blk)
                    {
                        uint32_t lo = k0, hi = k1 ;
                        while(lo < hi)
                        {
                            const uint32_t mid = lo+((hi-lo) >> 1) ;
                            if(cs[mid-k_lo] < a.colstart_blk) lo = mid+1 ; else hi = mid ;
                        }
                        k0 = lo ;
                    }
                    const uint32_t n = k1-k0 ;
                    uint32_t t = 0 ;
                    const uint32_t va = smem_u32(stage+meta[R+1])+(k0-k_lo)*32u+(uint32_t)r*16u ;
                    const uint32_t xa = smem_u32(stage+L::VAL_BYTES+L::COL_BYTES)+(k0-k_lo)*16u+(uint32_t)r*8u ;
                    if(n == 7u)
                    {
                        rt2_blocks_cm<7>(va, xa, acc0, acc1) ;       // the interior row of a T3 mesh
                        t = 7 ;
                    }
                    for( ; t+4 <= n ; t += 4) rt2_blocks_cm<4>(va+t*32, xa+t*16, acc0, acc1) ;
                    for( ; t < n ; t++)       rt2_blocks_cm<1>(va+t*32, xa+t*16, acc0, acc1) ;
                }
                else
                {
                    if(a.colstart_blk) k0 = row_lower_bound(a.col, k0, k1, a.colstart_blk) ;
                    for(uint32_t k = k0 ; k < k1 ; k++)
                    {
                        const double * v = a.vals+(size_t)k*4+2*r ;
                        const double xc = __ldg(a.x+(size_t)__ldg(a.col+k)*2+r) ;
                        acc0 = fma(ld_stream(v), xc, acc0) ;
                        acc1 = fma(ld_stream(v+1), xc, acc1) ;
                    }
                }
                const size_t i = (size_t)(r0+rl)*2+r ;
                // lane c holds the (row 0, row 1) partial sums of column c; it stores row r = c.  Both lanes of a block row
                // are active together (rl < nr), so the exchange is safe inside this branch.
                const double give = r == 0 ? acc1 : acc0 ;
                const double got = __shfl_xor_sync(__activemask(), give, 1) ;
                double yv = r == 0 ? acc0+got : got+acc1 ;
                if(MINUS_B) yv -= aux[lane] ;
                yv *= a.sign ;
                a.y[i] = yv ;
                if(DOT == DOT_YX) dsum[0] = fma(yv, aux[32+lane], dsum[0]) ;
                if(DOT == DOT_YY) dsum[0] = fma(yv, yv, dsum[0]) ;
                if(DOT == DOT_YW) dsum[0] = fma(yv, aux[64+lane], dsum[0]) ;
                if(DOT == DOT_OMEGA)
                {
                    const double di = a.d ? aux[96+lane] : 1. ;
                    const double t2 = yv*di, s2 = aux[64+lane]*di ;
                    dsum[0] = fma(t2, s2, dsum[0]) ;
                    dsum[1] = fma(t2, t2, dsum[1]) ;
                }
            }
            __syncwarp() ;
            if(lane == 0) mbar_arrive(empty+s) ;
        }
        cp_async_wait_group<0>() ;
    }
    if(DOT != DOT_NONE)
    {
        double tot[2] ;
        if(grid_sum<2, (W+NP)*32>(dsum, a.partials, a.st->ticket+TICKET_SPMV, tot) && threadIdx.x == 0)
            krylov_finalize(a.st, a.finalize, tot[0], tot[1]) ;
    }
}
