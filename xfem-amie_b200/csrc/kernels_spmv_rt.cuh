// kernels_spmv_rt.cuh -- stride-3 block-row SpMV, "row-thread" shared-memory pipeline (sm_100a).
//
// Third design of the K-SpMV kernel (history and ncu numbers: profiles/r01_notes.md):
//   1. plain warp-per-row (kernels_spmv.cuh)        : right traffic, latency-bound (dependent loads)
//   2. TMA-staged values, warp-per-row compute       : compute warps stall on the x gather and burn
//      (kernels_spmv_tma.cuh; a variant that also      150-200 instructions per row on reduction
//       staged x with cp.async was no better)          shuffles, predicates and pointer set-up
//   3. this one: operands staged in shared memory by asynchronous copies, and a compute mapping
//      with NO reduction: one THREAD per scalar row.
//
//   tile  = 10 consecutive block rows = 30 scalar rows = 30 lanes of ONE compute warp
//   warp W (producer): cp.async.bulk (1D TMA) of the tile's values + column indices -> full_v[s]
//   warps 0..W-1     : tile j of this CTA belongs to warp j % W.  G of its own tiles ahead, a warp
//                      waits full_v of that future tile and issues 8-byte cp.async gathers
//                      x[3 col .. 3 col + 2] -> xs[stage] for every block (lane <-> element of xs), plus
//                      the own-row entries of b / x / w / d it will need at the end (aux[stage]).
//                      cp.async.wait_group<G> = "the gathers of the tile I compute now landed".
//                      Compute: lane (row rl = lane/3, component r = lane%3) walks its row's blocks:
//                      3 x (LDS value, LDS x, DFMA) per block; shared-memory reads are
//                      bank-conflict-free for equal-length rows (row stride 243 doubles = 6 banks).
//                      y store is one coalesced 240-byte write per tile; fused dot as before.
#pragma once
#include "kernels_spmv_tma.cuh"

#define RT_ROWS 10

template<int NST, int CAP>
struct RtLayout
{
    static constexpr int VAL_BYTES = (CAP*72+16+15)/16*16 ;        // every region starts 16-byte aligned (TMA destinations)
    static constexpr int COL_BYTES = (CAP*4+16+15)/16*16 ;
    static constexpr int XS_BYTES = (CAP*24+15)/16*16 ;            // (>= 64 B: the gather's unpredicated index reads run up to 44 B past COL)
    static constexpr int AUX_BYTES = 4*32*8 ;                      // b, x, w, d of the tile's 30 scalar rows
    static constexpr int META_BYTES = 64 ;                         // 11 row pointers + 3 words
    static constexpr int STAGE_BYTES = (VAL_BYTES+COL_BYTES+XS_BYTES+AUX_BYTES+META_BYTES+127)/128*128 ;
    static constexpr int META_OFF = VAL_BYTES+COL_BYTES+XS_BYTES+AUX_BYTES ;
    static constexpr int TOTAL_BYTES = NST*STAGE_BYTES+2*NST*8+16 ;
} ;

// ld.shared.f64 with a compile-time byte offset folded into the instruction (no address arithmetic)
template<int OFF>
__device__ __forceinline__ double lds_f64(uint32_t base)
{
    double v ;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(base), "n"(OFF)) ;
    return v ;
}

// Loads of block Q.. of a batch, unrolled by template recursion so that every offset is an immediate.
// The volatile asm keeps the 6*NB loads together in front of the DFMAs (the compiler otherwise interleaves
// load/use pairs to save registers, which serialises on the LDS latency: one warp per scheduler here).
template<int Q, int NB>
struct RtLoad
{
    static __device__ __forceinline__ void run(uint32_t va, uint32_t xa, double (&v0)[NB], double (&v1)[NB], double (&v2)[NB],
                                               double (&x0)[NB], double (&x1)[NB], double (&x2)[NB])
    {
        v0[Q] = lds_f64<Q*72>(va) ; v1[Q] = lds_f64<Q*72+24>(va) ; v2[Q] = lds_f64<Q*72+48>(va) ;
        x0[Q] = lds_f64<Q*24>(xa) ; x1[Q] = lds_f64<Q*24+8>(xa) ;  x2[Q] = lds_f64<Q*24+16>(xa) ;
        RtLoad<Q+1, NB>::run(va, xa, v0, v1, v2, x0, x1, x2) ;
    }
} ;
template<int NB>
struct RtLoad<NB, NB>
{
    static __device__ __forceinline__ void run(uint32_t, uint32_t, double (&)[NB], double (&)[NB], double (&)[NB],
                                               double (&)[NB], double (&)[NB], double (&)[NB]) { }
} ;

// NB consecutive 3x3 blocks of one scalar row: acc_c += A(r,c) x_c   (va: shared address of A(r,0) of the first
// block, xa: of x_0 of the first block)
template<int NB>
__device__ __forceinline__ void rt_blocks(uint32_t va, uint32_t xa, double & acc0, double & acc1, double & acc2)
{
    double v0[NB], v1[NB], v2[NB], x0[NB], x1[NB], x2[NB] ;
    RtLoad<0, NB>::run(va, xa, v0, v1, v2, x0, x1, x2) ;
    #pragma unroll
    for(int q = 0 ; q < NB ; q++)
    {
        acc0 = fma(v0[q], x0[q], acc0) ;
        acc1 = fma(v1[q], x1[q], acc1) ;
        acc2 = fma(v2[q], x2[q], acc2) ;
    }
}

template<int DOT, bool MINUS_B, int W, int NST, int CAP, int G, int NB = 9, int NP = 1>
__global__ void __launch_bounds__((W+NP)*32) k_spmv_s3_rt(SpmvArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    static_assert(NST >= (G+1)*W, "stages: W tiles in compute + G*W tiles being gathered") ;
    constexpr int R = RT_ROWS ;
    using L = RtLayout<NST, CAP> ;
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
        tile_producer<R, NST, CAP, L::STAGE_BYTES, L::VAL_BYTES, L::META_OFF>(a, smem, full_v, empty, ntiles, lane, wid-W, NP) ;
    }
    else
    {
        const int rl = lane/3 ;                 // block row inside the tile (10 = idle lanes 30, 31)
        const int r = lane-rl*3 ;               // row component

        // tile number j (CTA-local): x of every block and the own-row aux entries -> shared memory, asynchronously
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
                // full_v[s] of this phase also says the stage is FREE: the producer only refills it after the
                // warp that computed its previous tile arrived on empty[s].  Nothing may be written into the
                // stage (aux included) before this wait.
                mbar_wait(full_v+s, (j/NST) & 1u) ;
                if(rl < (int)nr)
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
                    // lane <-> ELEMENT of the stage's x array (xs[e] = x[3 col(e/3) + e%3], e = lane + 32 g): the 32
                    // copies of one LDGSTS land in 256 contiguous bytes of shared memory and read ~4 global lines
                    // (a row's columns come in runs), where lane <-> block (three copies per lane, 24-byte stride)
                    // touched ~11 lines per instruction and cost 10.9 shared-memory wavefronts each: 303 of the
                    // kernel's 754 wavefronts per tile (ncu source page of the round-2 in-solve capture,
                    // profiles/r02_notes.md section 8).  Column-index reads first, then the copies.
                    constexpr int GE = (3*CAP+31)/32 ;
                    const uint32_t nel = nblk*3u ;
                    const uint32_t q0 = (uint32_t)lane/3u, q1 = ((uint32_t)lane+1u)/3u, q2 = ((uint32_t)lane+2u)/3u ;
                    const uint32_t c0 = (uint32_t)lane-3u*q0, c1 = (uint32_t)lane+1u-3u*q1, c2 = (uint32_t)lane+2u-3u*q2 ;
                    // index reads are NOT predicated: past the tile's last block they return whatever follows in the
                    // stage (the region is followed by xs; RtLayout keeps (32 GE)/3 indices inside the stage) and
                    // the value is only used under the copy's own predicate
                    const uint32_t * cs0 = cs+q0, * cs1 = cs+q1, * cs2 = cs+q2 ;
                    uint32_t cidx[GE] ;
                    #pragma unroll
                    for(int g = 0 ; g < GE ; g++)
                    {
                        const uint32_t m = (32u*g)%3u ;
                        cidx[g] = (m == 0u ? cs0 : (m == 1u ? cs1 : cs2))[(32u*g)/3u] ;
                    }
                    // 3 col + c < 2^32: a block column index beyond 1.4e9 would need terabytes of matrix
                    double * xl = xs+lane ;
                    const int left = (int)nel-lane ;       // lane + 32 g < nel  <=>  32 g < left
                    #pragma unroll
                    for(int g = 0 ; g < GE ; g++)
                    {
                        const uint32_t m = (32u*g)%3u ;
                        const uint32_t c = m == 0u ? c0 : (m == 1u ? c1 : c2) ;
                        if(32*g < left) cp_async_8(xl+32*g, a.x+(cidx[g]*3u+c)) ;
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
                    const uint32_t va = smem_u32(stage+meta[R+1])+((k0-k_lo)*9u+(uint32_t)r)*8u ;
                    const uint32_t xa = smem_u32(stage+L::VAL_BYTES+L::COL_BYTES)+(k0-k_lo)*24u ;
                    const uint32_t n = k1-k0 ;
                    if(n == 27u && NB == 9)
                    {
                        // the common row of a hexahedral mesh: straight-line code, every offset an immediate
                        rt_blocks<9>(va, xa, acc0, acc1, acc2) ;
                        rt_blocks<9>(va+9*72, xa+9*24, acc0, acc1, acc2) ;
                        rt_blocks<9>(va+18*72, xa+18*24, acc0, acc1, acc2) ;
                    }
                    else
                    {
                        uint32_t t = 0 ;
                        if(NB == 9)
                            for( ; t+9 <= n ; t += 9) rt_blocks<9>(va+t*72, xa+t*24, acc0, acc1, acc2) ;
                        for( ; t+3 <= n ; t += 3) rt_blocks<3>(va+t*72, xa+t*24, acc0, acc1, acc2) ;
                        for( ; t < n ; t++)       rt_blocks<1>(va+t*72, xa+t*24, acc0, acc1, acc2) ;
                    }
                }
                else
                {
                    // oversize tile (more than CAP blocks): operands straight from global memory
                    if(a.colstart_blk) k0 = row_lower_bound(a.col, k0, k1, a.colstart_blk) ;
                    for(uint32_t k = k0 ; k < k1 ; k++)
                    {
                        const double * v = a.vals+(size_t)k*9+r ;
                        const double * px = a.x+(size_t)__ldg(a.col+k)*3 ;
                        acc0 = fma(ld_stream(v), __ldg(px), acc0) ;
                        acc1 = fma(ld_stream(v+3), __ldg(px+1), acc1) ;
                        acc2 = fma(ld_stream(v+6), __ldg(px+2), acc2) ;
                    }
                }
                const size_t i = (size_t)(r0+rl)*3+r ;
                double yv = (acc0+acc1)+acc2 ;
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
