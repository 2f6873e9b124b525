// kernels_spmv_tma.cuh -- the TMA bulk-copy machinery of the SpMV pipelines (sm_100a): mbarrier / cp.async.bulk /
// cp.async wrappers and the tile producer shared by kernels_spmv_rt.cuh (3x3 blocks) and kernels_spmv_rt2.cuh (2x2).
// The warp-per-row consumer kernel that first used it (k_spmv_s3_tma, 5.2-5.4 TB/s) lost to the row-thread pipeline and
// is gone; its description stays below because the producer design is unchanged.
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

template<int R, int NST, int CAP, int STAGE_BYTES, int VAL_BYTES, int META_OFF, int PFD = 4, int BB = 72>
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
    // Unconditional, clamped loads (tile -> the last tile, row -> the end of the range; lanes past a short tile's rows
    // read a valid word nobody uses): a predicated load goes through a temporary register and a MOV that waits for it
    // right here, which stalled this warp for a whole memory latency per tile (7 % of the kernel's warp samples,
    // profiles/r02_notes.md section 8) instead of leaving the load in flight until its tile comes up PD tiles later.
    const uint32_t row_end = a.row0+a.nrows ;
    auto load_rp = [&](uint32_t t) -> uint32_t
    {
        const uint32_t r0 = a.row0+min(t, ntiles-1u)*R ;
        return __ldg(a.rowptr+min(r0+(uint32_t)lane, row_end)) ;
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
            mbar_wait(empty+s, ph^1u) ;
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
