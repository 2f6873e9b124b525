// kernels_spmv_tma.cuh -- stride-3 block-row SpMV with a TMA bulk-copy pipeline (sm_100a).
//
// Why: the plain kernel (kernels_spmv.cuh) moves the right bytes (ncu: dram bytes = algorithmic
// bytes) but is latency-bound -- a warp's loads (row pointer -> column indices -> x gather, with
// the values beside them) are serialised, and 32 warps/SM cannot keep ~45 KB/SM in flight
// (profiles/r01_spmv_s3_plain.md: 50 % DRAM utilisation, long-scoreboard stalls).
//
// How: a block row's values are one contiguous run in HBM, and so are the values of a TILE of
// consecutive rows.  One producer warp per CTA streams whole tiles (values + column indices) into
// shared memory with cp.async.bulk (the 1D TMA engine), completion on an mbarrier, NST stages
// ahead of the 8 consumer warps.  Bytes in flight are then set by NST x tile size per CTA and no
// longer by how many warps are resident or where they stall.  Consumers read values and indices
// from shared memory (conflict-free: lane l <-> element l of a 3-block group), gather x through
// L1/L2, and do one DFMA per value; the dot product that follows is fused as before.
//
//   tile      = R consecutive block rows (row = tile*R + j*8 + warp)
//   stage     = [ values: CAP*72 B | column indices: CAP*4 B | row pointers: (R+1)*4 B ]
//   alignment = bulk copies need 16-byte addresses and sizes: the source is rounded down to 16 B
//               and the few leading bytes are skipped on the shared-memory side
//               (device arrays carry 16 B of tail padding for the rounded-up end).
//   oversize  = a tile with more than CAP blocks is not staged: consumers fall back to global loads.
#pragma once
#include "kernels_spmv.cuh"

__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p) ; }

__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)) ;
}
__device__ __forceinline__ void mbar_arrive(uint64_t * bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory") ;
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t * bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory") ;
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory") ;
}
// one lane polls the barrier, the others wait at the warp barrier: 31 lanes fewer hammering the shared-memory pipe
// while a warp waits (the producer spends most of its life here).  The elected lane's acquire and the __syncwarp order
// the stage's contents before every lane's later reads.
__device__ __forceinline__ void mbar_wait_elect(uint64_t * bar, uint32_t parity, int lane)
{
    if(lane == 0) mbar_wait(bar, parity) ;
    __syncwarp() ;
}
__device__ __forceinline__ void tma_bulk_g2s(void * dst_smem, const void * src_gmem, uint32_t bytes, uint64_t * bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory") ;
}


// Producer warp shared by the TMA kernels: streams tiles (values + column indices) into the ring.
// The row pointers of the next PD tiles are prefetched in registers: without that, the producer's
// own dependent rowptr load (a DRAM miss per tile) caps a CTA at one tile per memory latency.
__device__ __forceinline__ void cp_async_8(void * dst_smem, const void * src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(dst_smem)), "l"(src) : "memory") ;
}
__device__ __forceinline__ void cp_async_16(void * dst_smem, const void * src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_u32(dst_smem)), "l"(src) : "memory") ;
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;" ::: "memory") ;
}
template<int N> __device__ __forceinline__ void cp_async_wait_group()
{
    asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory") ;
}
__device__ __forceinline__ void tma_prefetch_l2(const void * src_gmem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src_gmem), "r"(bytes) : "memory") ;
}

template<int R, int NST, int CAP, int STAGE_BYTES, int VAL_BYTES, int META_OFF, int PFD = 4, int BB = 72, bool ELECT = false>
__device__ __forceinline__ void tile_producer(const SpmvArgs & a, unsigned char * smem, uint64_t * full, uint64_t * empty,
                                              uint32_t ntiles, int lane, uint32_t first = 0, uint32_t step = 1)
{
    // `first`/`step`: several producer warps share the CTA's tile sequence (warp p takes tiles p, p+step, ...):
    // one producer spends ~0.3 us per tile, which caps a CTA at tile_bytes/0.3 us -- too little for short rows
    // PD row-pointer sets in flight; tile it+PFD is pulled into L2 (TMA prefetch, no shared memory
    // needed) while tile it is copied into its stage: the stage copies then see L2 latency, not HBM's.
    constexpr int PD = 8 ;
    static_assert(PFD < PD, "prefetch distance") ;
    uint32_t rpq[PD] ;
    auto load_rp = [&](uint32_t t) -> uint32_t
    {
        if(t >= ntiles) return 0u ;
        const uint32_t r0 = a.row0+t*R ;
        const uint32_t nr = min((uint32_t)R, a.row0+a.nrows-r0) ;
        return lane <= (int)nr ? __ldg(a.rowptr+r0+lane) : 0u ;
    } ;
    const uint32_t tstride = step*gridDim.x ;
    #pragma unroll
    for(int j = 0 ; j < PD ; j++) rpq[j] = load_rp(blockIdx.x+(first+j*step)*gridDim.x) ;
    uint32_t it = first ;
    uint32_t tile = blockIdx.x+first*gridDim.x ;
    while(tile < ntiles)
    {
        #pragma unroll
        for(int j = 0 ; j < PD ; j++)
        {
            if(tile >= ntiles) break ;
            const uint32_t rp = rpq[j] ;
            rpq[j] = load_rp(tile+PD*tstride) ;
            if(PFD > 0)
            {
                const uint32_t tp = tile+PFD*tstride ;
                if(tp < ntiles)
                {
                    const uint32_t rpp = rpq[(j+PFD)%PD] ;
                    const uint32_t r0p = a.row0+tp*R ;
                    const uint32_t nrp = min((uint32_t)R, a.row0+a.nrows-r0p) ;
                    const uint32_t p_lo = __shfl_sync(0xffffffffu, rpp, 0) ;
                    const uint32_t p_hi = __shfl_sync(0xffffffffu, rpp, nrp & 31) ;
                    if(lane == 0 && p_hi > p_lo && p_hi-p_lo <= (uint32_t)CAP)
                    {
                        const uint64_t va_lo = ((uint64_t)p_lo*BB) & ~15ull ;
                        const uint64_t ca_lo = ((uint64_t)p_lo*4) & ~15ull ;
                        tma_prefetch_l2(reinterpret_cast<const unsigned char *>(a.vals)+va_lo, (uint32_t)((((uint64_t)p_hi*BB-va_lo)+15ull) & ~15ull)) ;
                        tma_prefetch_l2(reinterpret_cast<const unsigned char *>(a.col)+ca_lo, (uint32_t)((((uint64_t)p_hi*4-ca_lo)+15ull) & ~15ull)) ;
                    }
                }
            }
            const int s = it%NST ;
            const uint32_t ph = (it/NST) & 1u ;
            if(ELECT) mbar_wait_elect(empty+s, ph^1u, lane) ; else mbar_wait(empty+s, ph^1u) ;
            unsigned char * stage = smem+s*STAGE_BYTES ;
            uint32_t * meta = reinterpret_cast<uint32_t *>(stage+META_OFF) ;
            const uint32_t r0 = a.row0+tile*R ;
            const uint32_t nr = min((uint32_t)R, a.row0+a.nrows-r0) ;
            if(lane <= R) meta[lane] = rp ;
            const uint32_t k_lo = __shfl_sync(0xffffffffu, rp, 0) ;
            const uint32_t k_hi = __shfl_sync(0xffffffffu, rp, nr & 31) ;
            const uint32_t nblk = k_hi-k_lo ;
            const bool staged = nblk <= (uint32_t)CAP && nblk > 0 ;
            if(lane == 0)
            {
                meta[R+1] = (uint32_t)(((uint64_t)k_lo*BB) & 15ull) ;
                meta[R+2] = (uint32_t)(((uint64_t)k_lo*4) & 15ull) ;
                meta[R+3] = staged ? 1u : 0u ;
            }
            __syncwarp() ;
            if(lane == 0)
            {
                if(staged)
                {
                    const uint64_t vb_lo = (uint64_t)k_lo*BB, vb_hi = (uint64_t)k_hi*BB ;
                    const uint64_t va_lo = vb_lo & ~15ull ;
                    const uint32_t vbytes = (uint32_t)(((vb_hi-va_lo)+15ull) & ~15ull) ;
                    const uint64_t cb_lo = (uint64_t)k_lo*4, cb_hi = (uint64_t)k_hi*4 ;
                    const uint64_t ca_lo = cb_lo & ~15ull ;
                    const uint32_t cbytes = (uint32_t)(((cb_hi-ca_lo)+15ull) & ~15ull) ;
                    mbar_arrive_expect_tx(full+s, vbytes+cbytes) ;
                    tma_bulk_g2s(stage, reinterpret_cast<const unsigned char *>(a.vals)+va_lo, vbytes, full+s) ;
                    tma_bulk_g2s(stage+VAL_BYTES, reinterpret_cast<const unsigned char *>(a.col)+ca_lo, cbytes, full+s) ;
                }
                else
                    mbar_arrive(full+s) ;
            }
            tile += tstride ;
            it += step ;
        }
    }
}

template<int R, int NST, int CAP>
struct TmaStageLayout
{
    static constexpr int VAL_BYTES = CAP*72+16 ;
    static constexpr int COL_BYTES = CAP*4+16 ;
    static constexpr int META_BYTES = ((R+1+3)*4+15)/16*16 ;       // row pointers + (first value byte offset, first col offset, staged flag)
    static constexpr int STAGE_BYTES = VAL_BYTES+COL_BYTES+META_BYTES ;
    static constexpr int TOTAL_BYTES = NST*STAGE_BYTES+2*NST*8+16 ;
} ;

template<int UMAX>
__device__ __forceinline__ double s3_chunk_smem(const double * __restrict__ vs, const uint32_t * __restrict__ cs,
                                                const double * __restrict__ x, int nblk, int lane, int slot, int cc, double acc)
{
    double v[UMAX] ;
    double xv[UMAX] ;
    #pragma unroll
    for(int u = 0 ; u < UMAX ; u++)
    {
        const bool ok = (lane < 27) && (3*u+slot < nblk) ;
        v[u] = ok ? vs[u*27+lane] : 0. ;
        const uint32_t c = ok ? cs[3*u+slot] : 0u ;
        xv[u] = ok ? __ldg(x+(size_t)c*3+cc) : 0. ;
    }
    #pragma unroll
    for(int u = 0 ; u < UMAX ; u++)
        acc = fma(v[u], xv[u], acc) ;
    return acc ;
}

template<int DOT, bool MINUS_B, int R, int NST, int CAP, bool PF>
__global__ void __launch_bounds__(PF ? 320 : 288) k_spmv_s3_tma(SpmvArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    using L = TmaStageLayout<R, NST, CAP> ;
    extern __shared__ __align__(128) unsigned char smem[] ;
    uint64_t * full = reinterpret_cast<uint64_t *>(smem+NST*L::STAGE_BYTES) ;
    uint64_t * empty = full+NST ;
    const int lane = threadIdx.x & 31 ;
    const int wid = threadIdx.x >> 5 ;
    const uint32_t ntiles = (a.nrows+R-1)/R ;

    if(threadIdx.x == 0)
    {
        for(int s = 0 ; s < NST ; s++)
        {
            mbar_init(full+s, 1) ;
            mbar_init(empty+s, PF ? 9 : 8) ;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory") ;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory") ;
    }
    __syncthreads() ;

    double dsum[2] = {0., 0.} ;

    if(wid == 8)
    {
        static_assert(R < 32, "a tile's row pointers are held by one warp") ;
        tile_producer<R, NST, CAP, L::STAGE_BYTES, L::VAL_BYTES, L::VAL_BYTES+L::COL_BYTES>(a, smem, full, empty, ntiles, lane) ;
    }
    else if(PF && wid == 9)
    {
        // ---------------- prefetch warp: pull the x lines of landed tiles into L2 ahead of the consumers
        uint32_t it = 0 ;
        for(uint32_t tile = blockIdx.x ; tile < ntiles ; tile += gridDim.x, it++)
        {
            const int s = it%NST ;
            const uint32_t ph = (it/NST) & 1u ;
            const unsigned char * stage = smem+s*L::STAGE_BYTES ;
            const uint32_t * meta = reinterpret_cast<const uint32_t *>(stage+L::VAL_BYTES+L::COL_BYTES) ;
            const uint32_t r0 = a.row0+tile*R ;
            const uint32_t nr = min((uint32_t)R, a.row0+a.nrows-r0) ;
            mbar_wait(full+s, ph) ;
            if(meta[R+3] != 0u)
            {
                const uint32_t * cbase = reinterpret_cast<const uint32_t *>(stage+L::VAL_BYTES+meta[R+2]) ;
                const uint32_t nblk = meta[nr]-meta[0] ;
                for(uint32_t b = lane ; b < nblk ; b += 32)
                {
                    const uint32_t c = cbase[b] ;
                    const bool run_start = (b == 0) || (cbase[b-1]+1u != c) ;
                    if(run_start)
                    {
                        const double * px = a.x+(size_t)c*3 ;
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(px)) ;
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(px+8)) ;
                    }
                }
            }
            __syncwarp() ;
            if(lane == 0) mbar_arrive(empty+s) ;
        }
    }
    else
    {
        // ---------------- 8 consumer warps
        const int slot = lane/9 ;
        const int e = lane-slot*9 ;
        const int cc = e/3 ;
        uint32_t it = 0 ;
        for(uint32_t tile = blockIdx.x ; tile < ntiles ; tile += gridDim.x, it++)
        {
            const int s = it%NST ;
            const uint32_t ph = (it/NST) & 1u ;
            const unsigned char * stage = smem+s*L::STAGE_BYTES ;
            const uint32_t * meta = reinterpret_cast<const uint32_t *>(stage+L::VAL_BYTES+L::COL_BYTES) ;
            const uint32_t r0 = a.row0+tile*R ;
            const uint32_t nr = min((uint32_t)R, a.row0+a.nrows-r0) ;
            mbar_wait(full+s, ph) ;
            const uint32_t k_lo = meta[0] ;
            const bool staged = meta[R+3] != 0u ;
            const double * vbase = reinterpret_cast<const double *>(stage+meta[R+1]) ;
            const uint32_t * cbase = reinterpret_cast<const uint32_t *>(stage+L::VAL_BYTES+meta[R+2]) ;
            #pragma unroll 1
            for(uint32_t lr = wid ; lr < nr ; lr += 8)
            {
                const uint32_t row = r0+lr ;
                uint32_t k0 = meta[lr] ;
                const uint32_t k1 = meta[lr+1] ;
                double acc = 0. ;
                if(staged)
                {
                    if(a.colstart_blk)
                    {
                        // lower_bound over the row's staged column indices
                        uint32_t lo = k0, hi = k1 ;
                        while(lo < hi)
                        {
                            const uint32_t mid = lo+((hi-lo) >> 1) ;
                            if(cbase[mid-k_lo] < a.colstart_blk) lo = mid+1 ; else hi = mid ;
                        }
                        k0 = lo ;
                    }
                    for(uint32_t kb = k0 ; kb < k1 ; kb += 27u)
                    {
                        const int nblk = (int)min(27u, k1-kb) ;
                        const double * vs = vbase+(size_t)(kb-k_lo)*9 ;
                        const uint32_t * cs = cbase+(kb-k_lo) ;
                        if(nblk > 18)      acc = s3_chunk_smem<9>(vs, cs, a.x, nblk, lane, slot, cc, acc) ;
                        else if(nblk > 9)  acc = s3_chunk_smem<6>(vs, cs, a.x, nblk, lane, slot, cc, acc) ;
                        else               acc = s3_chunk_smem<3>(vs, cs, a.x, nblk, lane, slot, cc, acc) ;
                    }
                }
                else
                {
                    if(a.colstart_blk) k0 = row_lower_bound(a.col, k0, k1, a.colstart_blk) ;
                    for(uint32_t kb = k0 ; kb < k1 ; kb += 27u)
                    {
                        const int nblk = (int)min(27u, k1-kb) ;
                        const uint32_t colreg = lane < nblk ? __ldg(a.col+kb+lane) : 0u ;
                        acc = s3_chunk<9>(a.vals+(size_t)kb*9, a.x, colreg, nblk, lane, slot, cc, acc) ;
                    }
                }
                acc += __shfl_down_sync(0xffffffffu, acc, 9)+__shfl_down_sync(0xffffffffu, acc, 18) ;
                acc += __shfl_down_sync(0xffffffffu, acc, 3)+__shfl_down_sync(0xffffffffu, acc, 6) ;
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
            __syncwarp() ;
            if(lane == 0) mbar_arrive(empty+s) ;
        }
    }
    if(DOT != DOT_NONE)
    {
        double tot[2] ;
        if(grid_sum<2, PF ? 320 : 288>(dsum, a.partials, a.st->ticket+TICKET_SPMV, tot) && threadIdx.x == 0)
            krylov_finalize(a.st, a.finalize, tot[0], tot[1]) ;
    }
}
