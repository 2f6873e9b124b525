// kernels_spmv_rtm.cuh -- stride-3 row-thread pipeline with TEAMS of warps per tile (sm_100a).
//
// kernels_spmv_rt.cuh keeps one compute warp per scheduler (3 compute warps + 1 producer per SM: the 197 KB stage
// ring is what limits the CTA count, not the warps).  ncu on that kernel inside the solve
// (profiles/r01_ncu_insolve_S3hex256_rt_final.txt): issue slots 19 % busy, 1.83 `wait` + 1.17 `short_scoreboard`
// stall cycles per issued instruction, nothing to switch to -- the kernel follows the SM clock, and the SM clock
// follows the power cap.  Here the SAME pipeline (same producer, same stages, same bytes through shared memory)
// gives every tile to a TEAM of TM warps: member m multiplies the m-th part of every row of the tile
// (27 blocks -> 9 + 9 + 9), members 1.. hand their 30 partial sums to member 0 through 240 B of shared memory, and
// member 0 adds them in a fixed order and finishes the rows (b, sign, store, fused dot).  The x gather of a tile is
// shared between the members as well.  T teams x TM warps are resident: the schedulers now hold 2-3 warps each, so
// a warp waiting on a DFMA chain or an LDS no longer idles its scheduler.
//
//   named barrier 1+team (TM*32 threads): (a) "everybody's cp.async gathers of this tile have landed",
//                                          (b) "the partial sums are in shared memory / the stage is no longer read".
#pragma once
#include "kernels_spmv_rt.cuh"

__device__ __forceinline__ void bar_sync_named(int id, int count)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory") ;
}

template<int NST, int CAP, int T, int TM>
struct RtmLayout
{
    using Base = RtLayout<NST, CAP> ;
    static constexpr int SCRATCH_OFF = (Base::TOTAL_BYTES+15)/16*16 ;
    static constexpr int SCRATCH_BYTES = T*(TM > 1 ? TM-1 : 1)*32*8 ;
    static constexpr int TOTAL_BYTES = SCRATCH_OFF+SCRATCH_BYTES ;
} ;

template<int DOT, bool MINUS_B, int T, int TM, int NST, int CAP, int G, int NP = 1>
__global__ void __launch_bounds__((T*TM+NP)*32) k_spmv_s3_rtm(SpmvArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    static_assert(NST >= (G+1)*T, "stages: T tiles in compute + G*T tiles being gathered") ;
    static_assert(T*TM+1 <= 15, "one named barrier per team") ;
    constexpr int R = RT_ROWS ;
    constexpr int NB = 9 ;
    using L = RtLayout<NST, CAP> ;
    using LM = RtmLayout<NST, CAP, T, TM> ;
    extern __shared__ __align__(128) unsigned char smem[] ;
    uint64_t * full_v = reinterpret_cast<uint64_t *>(smem+NST*L::STAGE_BYTES) ;
    uint64_t * empty = full_v+NST ;
    double * scratch = reinterpret_cast<double *>(smem+LM::SCRATCH_OFF) ;
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

    if(wid >= T*TM)
    {
        tile_producer<R, NST, CAP, L::STAGE_BYTES, L::VAL_BYTES, L::META_OFF>(a, smem, full_v, empty, ntiles, lane, wid-T*TM, NP) ;
    }
    else
    {
        const int team = wid/TM ;
        const int m = wid-team*TM ;
        const int rl = lane/3 ;                 // block row inside the tile (10 = idle lanes 30, 31)
        const int r = lane-rl*3 ;               // row component
        double * my_scratch = scratch+(size_t)team*(TM > 1 ? TM-1 : 1)*32 ;

        // CTA-local tile j: this member's share of the x gather (and, member 0, the own-row aux entries)
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
                // full_v[s] of this phase also says the stage is FREE (see kernels_spmv_rt.cuh)
                mbar_wait(full_v+s, (j/NST) & 1u) ;
                if(m == 0 && rl < (int)nr)
                {
                    const size_t i = (size_t)(r0+rl)*3+r ;
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
                    constexpr int GM = (GI+TM-1)/TM ;          // rounds of 32 blocks per member
                    uint32_t cidx[GM] ;
                    #pragma unroll
                    for(int g = 0 ; g < GM ; g++)
                    {
                        const uint32_t bk = lane+32u*(uint32_t)(g*TM+m) ;
                        cidx[g] = bk < nblk ? cs[bk] : 0u ;
                    }
                    #pragma unroll
                    for(int g = 0 ; g < GM ; g++)
                    {
                        const uint32_t bk = lane+32u*(uint32_t)(g*TM+m) ;
                        if(bk < nblk)
                        {
                            const double * px = a.x+(size_t)cidx[g]*3 ;
                            double * d = xs+(size_t)bk*3 ;
                            cp_async_8(d, px) ;
                            cp_async_8(d+1, px+1) ;
                            cp_async_8(d+2, px+2) ;
                        }
                    }
                }
            }
            cp_async_commit() ;
        } ;

        #pragma unroll
        for(int g = 0 ; g < G ; g++) issue_gather(team+g*T) ;

        for(uint32_t j = team ; blockIdx.x+j*gridDim.x < ntiles ; j += T)
        {
            const uint32_t tile = blockIdx.x+j*gridDim.x ;
            issue_gather(j+G*T) ;
            cp_async_wait_group<G>() ;
            if(TM > 1) bar_sync_named(1+team, TM*32) ; else __syncwarp() ;
            const int s = j%NST ;
            const unsigned char * stage = smem+s*L::STAGE_BYTES ;
            const uint32_t * meta = reinterpret_cast<const uint32_t *>(stage+L::META_OFF) ;
            const double * aux = reinterpret_cast<const double *>(stage+L::VAL_BYTES+L::COL_BYTES+L::XS_BYTES) ;
            const uint32_t r0 = a.row0+tile*R ;
            const uint32_t nr = min((uint32_t)R, a.row0+a.nrows-r0) ;
            const uint32_t k_lo = meta[0] ;
            const bool staged = meta[R+3] != 0u ;
            double part = 0. ;
            if(rl < (int)nr)
            {
                uint32_t k0 = meta[rl] ;
                const uint32_t k1 = meta[rl+1] ;
                double acc0 = 0., acc1 = 0., acc2 = 0. ;
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
                    // this member's part of the row
                    const uint32_t nrow = k1-k0 ;
                    const uint32_t c0 = k0+(nrow*(uint32_t)m)/TM, c1 = k0+(nrow*(uint32_t)(m+1))/TM ;
                    const uint32_t va = smem_u32(stage+meta[R+1])+((c0-k_lo)*9u+(uint32_t)r)*8u ;
                    const uint32_t xa = smem_u32(stage+L::VAL_BYTES+L::COL_BYTES)+(c0-k_lo)*24u ;
                    const uint32_t n = c1-c0 ;
                    if(n == 9u)
                        rt_blocks<9>(va, xa, acc0, acc1, acc2) ;
                    else
                    {
                        uint32_t t = 0 ;
                        for( ; t+9 <= n ; t += 9) rt_blocks<9>(va+t*72, xa+t*24, acc0, acc1, acc2) ;
                        for( ; t+3 <= n ; t += 3) rt_blocks<3>(va+t*72, xa+t*24, acc0, acc1, acc2) ;
                        for( ; t < n ; t++)       rt_blocks<1>(va+t*72, xa+t*24, acc0, acc1, acc2) ;
                    }
                }
                else
                {
                    // oversize tile (more than CAP blocks): operands straight from global memory
                    if(a.colstart_blk) k0 = row_lower_bound(a.col, k0, k1, a.colstart_blk) ;
                    const uint32_t nrow = k1-k0 ;
                    const uint32_t c0 = k0+(nrow*(uint32_t)m)/TM, c1 = k0+(nrow*(uint32_t)(m+1))/TM ;
                    for(uint32_t k = c0 ; k < c1 ; k++)
                    {
                        const double * v = a.vals+(size_t)k*9+r ;
                        const double * px = a.x+(size_t)__ldg(a.col+k)*3 ;
                        acc0 = fma(ld_stream(v), __ldg(px), acc0) ;
                        acc1 = fma(ld_stream(v+3), __ldg(px+1), acc1) ;
                        acc2 = fma(ld_stream(v+6), __ldg(px+2), acc2) ;
                    }
                }
                part = (acc0+acc1)+acc2 ;
            }
            if(TM > 1)
            {
                if(m > 0) my_scratch[(m-1)*32+lane] = part ;
                bar_sync_named(1+team, TM*32) ;
            }
            if(m == 0)
            {
                if(rl < (int)nr)
                {
                    double yv = part ;
                    #pragma unroll
                    for(int q = 1 ; q < TM ; q++) yv += my_scratch[(q-1)*32+lane] ;       // fixed order: run-to-run identical bits
                    const size_t i = (size_t)(r0+rl)*3+r ;
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
                // the aux reads above are member 0's own; members 1.. left the stage at the barrier
                if(lane == 0) mbar_arrive(empty+s) ;
            }
        }
        cp_async_wait_group<0>() ;
    }
    if(DOT != DOT_NONE)
    {
        double tot[2] ;
        if(grid_sum<2, (T*TM+NP)*32>(dsum, a.partials, a.st->ticket+TICKET_SPMV, tot) && threadIdx.x == 0)
            krylov_finalize(a.st, a.finalize, tot[0], tot[1]) ;
    }
}
