// spmv_launch.cu -- kernel selection and launch of the block-row SpMV: the ONE translation unit that instantiates the
// SpMV kernels (kernels_spmv*.cuh).  Every caller (the solvers, dist.cu, the C-ABI) goes through launch_spmv /
// launch_spmv_range declared in launch.cuh.
#include "launch.cuh"
#include "kernels_spmv.cuh"
#include "kernels_spmv_tma.cuh"
#include "kernels_spmv_rt.cuh"
#include "kernels_spmv_rt2.cuh"

static inline int persistent_grid(const amie_b200_ctx * ctx, int per_sm, uint32_t ntiles)
{
    if(per_sm < 1) per_sm = 1 ;
    uint64_t cap = (uint64_t)ctx->num_sms*per_sm ;
    if(cap > AMIE_MAX_PARTIALS) cap = AMIE_MAX_PARTIALS ;
    uint64_t g = ntiles < cap ? ntiles : cap ;
    return (int)(g ? g : 1) ;
}

// Occupancy and the opt-in to > 48 KB of dynamic shared memory are properties of (kernel, DEVICE): one cache slot per
// device ordinal and kernel instantiation.  A process may hold contexts on several devices (group.cu does).
#define AMIE_MAX_DEVICES 64
static inline int device_slot(const amie_b200_ctx * ctx) { return ctx->device >= 0 && ctx->device < AMIE_MAX_DEVICES ? ctx->device : 0 ; }

#define SPMV_LAUNCH(KERNEL, ROWS_PER_TILE) do { \
        static int per_sm_of[AMIE_MAX_DEVICES] = {} ; \
        int & per_sm = per_sm_of[device_slot(ctx)] ; \
        if(!per_sm) { int q = 0 ; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, KERNEL, 256, 0) ; per_sm = q < 1 ? 1 : q ; } \
        uint32_t ntiles = (uint32_t)((args.nrows+(ROWS_PER_TILE)-1)/(ROWS_PER_TILE)) ; \
        int grid = persistent_grid(ctx, per_sm, ntiles) ; \
        KERNEL<<<grid, 256, 0, ctx->stream>>>(args) ; } while(0)

// first use of a kernel instantiation on a device: raise its dynamic shared-memory limit there, query its occupancy
template<typename K>
static inline int smem_kernel_per_sm(const amie_b200_ctx * ctx, K kern, int threads, int smem, int * cache)
{
    int & per_sm = cache[device_slot(ctx)] ;
    if(!per_sm)
    {
        int q = 0 ;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) ;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, threads, smem) ;
        per_sm = q < 1 ? 1 : q ;
    }
    return per_sm ;
}

// row-thread pipeline: W compute warps (one 10-row tile each at a time), NST stages, CAP blocks per stage
template<int DOT, bool MINUS_B, int W, int NST, int CAP, int G, int NB = 9, int NP = 1>
static inline void launch_s3_rt(amie_b200_ctx * ctx, const SpmvArgs & args)
{
    auto kern = k_spmv_s3_rt<DOT, MINUS_B, W, NST, CAP, G, NB, NP> ;
    constexpr int smem = RtLayout<NST, CAP>::TOTAL_BYTES ;
    constexpr int threads = (W+NP)*32 ;
    static int cache[AMIE_MAX_DEVICES] = {} ;
    const int per_sm = smem_kernel_per_sm(ctx, kern, threads, smem, cache) ;
    const uint32_t ntiles = (args.nrows+RT_ROWS-1)/RT_ROWS ;
    const int grid = persistent_grid(ctx, per_sm, ntiles) ;
    kern<<<grid, threads, smem, ctx->stream>>>(args) ;
}

template<int DOT, bool MINUS_B, int W, int NST, int CAP, int G, int NP = 1>
static inline void launch_s2_rt(amie_b200_ctx * ctx, const SpmvArgs & args)
{
    auto kern = k_spmv_s2_rt<DOT, MINUS_B, W, NST, CAP, G, NP> ;
    constexpr int smem = Rt2Layout<NST, CAP>::TOTAL_BYTES ;
    constexpr int threads = (W+NP)*32 ;
    static int cache[AMIE_MAX_DEVICES] = {} ;
    const int per_sm = smem_kernel_per_sm(ctx, kern, threads, smem, cache) ;
    const uint32_t ntiles = (args.nrows+RT2_ROWS-1)/RT2_ROWS ;
    const int grid = persistent_grid(ctx, per_sm, ntiles) ;
    kern<<<grid, threads, smem, ctx->stream>>>(args) ;
}


template<int DOT, bool MINUS_B>
static inline void spmv_dispatch(amie_b200_ctx * ctx, const SpmvArgs & args)
{
    if(ctx->S != 2 && ctx->S != 3)
    {
        const uint32_t nt = (uint32_t)(((uint64_t)args.nrows*ctx->S+255)/256) ;
        const int grid = persistent_grid(ctx, 8, nt) ;
        if(ctx->S == 1)      k_spmv_gen<1, DOT, MINUS_B><<<grid, 256, 0, ctx->stream>>>(args) ;
        else if(ctx->S == 4) k_spmv_gen<4, DOT, MINUS_B><<<grid, 256, 0, ctx->stream>>>(args) ;
        else                 k_spmv_gen<6, DOT, MINUS_B><<<grid, 256, 0, ctx->stream>>>(args) ;
        return ;
    }
    if(ctx->S == 3)
    {
        // default: row-thread pipeline (a smaller-stage configuration for short rows was measured
        // and brought nothing: profiles/r01_notes.md)
        if(ctx->opt_variant == 1)
            SPMV_LAUNCH((k_spmv_s3<DOT, MINUS_B>), 8) ;
        else
        {
            // row-thread TMA pipeline; the stage capacity follows the mean row length so that a tile of 10 rows
            // fills its stage (27 blocks/row for Q1 hexahedra, 12-15 for linear tetrahedra).  Short rows mean small
            // tiles, and ONE producer warp (~0.3 us per tile) then caps the CTA: three producers there.
            const double avg = ctx->nb ? (double)ctx->nnzb/(double)ctx->nb : 0. ;
            // Stages fill the 227 KB: 8 x 28 KB / 13 x 17 KB.  (Round 1 kept ~30 KB back as L1 for the x gather; since the
            // gather touches a third of the lines it did, the extra stage in flight is worth more: hexahedra 6.27 -> 6.23 ms
            // in the solve, tetrahedra 3.50 -> 3.38 ms, profiles/r02_notes.md section 9.)
            if(avg > 15.5 || ctx->opt_variant == 3)
                launch_s3_rt<DOT, MINUS_B, 3, 8, 270, 1>(ctx, args) ;
            else
                launch_s3_rt<DOT, MINUS_B, 5, 13, 160, 1, 9, 3>(ctx, args) ;
        }
    }
    else
    {
        // rows are short in 2D (about 7 blocks): 8 lanes per row unless rows are long
        const double avg = ctx->nb ? (double)ctx->nnzb/(double)ctx->nb : 0. ;
        int G = avg > 24. ? 32 : (avg > 10. ? 16 : 8) ;
        if((ctx->opt_variant == 0 && avg <= 9.) || ctx->opt_variant == 3)
        {
            // row-thread TMA pipeline for 2x2 blocks (kernels_spmv_rt2.cuh).  A 16-row tile is only ~3.6 KB, so the per-tile
            // producer cost dominates: FOUR producer warps (1 producer: 1.4 TB/s, 4: 4.8 TB/s on S2-tri-4096); 24 stages
            // instead of 20 change nothing (profiles/r02_notes.md section 9)
            launch_s2_rt<DOT, MINUS_B, 8, 20, 144, 1, 4>(ctx, args) ;
            return ;
        }
        if(ctx->opt_variant == 8 || ctx->opt_variant == 16 || ctx->opt_variant == 32) G = ctx->opt_variant ;
        if(G == 32)      SPMV_LAUNCH((k_spmv_s2<32, DOT, MINUS_B>), 8) ;
        else if(G == 16) SPMV_LAUNCH((k_spmv_s2<16, DOT, MINUS_B>), 16) ;
        else             SPMV_LAUNCH((k_spmv_s2<8, DOT, MINUS_B>), 32) ;
    }
}

// block rows [blk_row0, blk_row0+blk_nrows) of the local matrix; `finalize` overrides c.finalize
int launch_spmv_range(amie_b200_ctx * ctx, const SpmvCall & c, uint32_t blk_row0, uint32_t blk_nrows, int finalize)
{
    SpmvArgs args ;
    args.rowptr = ctx->rowptr ; args.col = ctx->col ; args.vals = ctx->vals ;
    args.x = c.x ; args.b = c.b ; args.y = c.y ; args.w = c.w ; args.d = c.d ;
    args.row0 = blk_row0 ;
    args.nrows = blk_nrows ;
    args.colstart_blk = (uint32_t)(c.colstart/ctx->S) ;
    args.sign = c.sign ;
    args.st = ctx->st ;
    args.partials = ctx->partials ;
    args.finalize = finalize ;
    args.check_stop = c.check_stop ;
    if(c.dot == DOT_NONE && !c.minus_b)      spmv_dispatch<DOT_NONE, false>(ctx, args) ;
    else if(c.dot == DOT_NONE && c.minus_b)  spmv_dispatch<DOT_NONE, true>(ctx, args) ;
    else if(c.dot == DOT_YY && c.minus_b)    spmv_dispatch<DOT_YY, true>(ctx, args) ;
    else if(c.dot == DOT_YX && !c.minus_b)   spmv_dispatch<DOT_YX, false>(ctx, args) ;
    else if(c.dot == DOT_YW && !c.minus_b)   spmv_dispatch<DOT_YW, false>(ctx, args) ;
    else if(c.dot == DOT_OMEGA && !c.minus_b) spmv_dispatch<DOT_OMEGA, false>(ctx, args) ;
    else { ctx->set_error("launch_spmv: unsupported combination") ; return AMIE_B200_ERR_ARG ; }
    ctx->stats.kernel_launches++ ;
    // a refused launch (shared-memory opt-in missing on this device, bad configuration) must not go unnoticed: the
    // solver would iterate on stale reduction scalars
    const cudaError_t le = cudaPeekAtLastError() ;
    if(le != cudaSuccess) { ctx->set_error(std::string("SpMV launch: ")+cudaGetErrorString(le)) ; return AMIE_B200_ERR_CUDA ; }
    return AMIE_B200_OK ;
}

int launch_spmv(amie_b200_ctx * ctx, const SpmvCall & c)
{
    if(ctx->dist) return dist_spmv(ctx, c) ;
    cudaEvent_t e0 = nullptr, e1 = nullptr ;
    if(ctx->opt_time_spmv && ctx->ev_used+2 <= ctx->ev_pool.size())
    {
        e0 = ctx->ev_pool[ctx->ev_used++] ;
        e1 = ctx->ev_pool[ctx->ev_used++] ;
        cudaEventRecord(e0, ctx->stream) ;
    }
    const uint32_t row0 = (uint32_t)(c.rowstart/ctx->S) ;
    int rc = launch_spmv_range(ctx, c, row0, (uint32_t)(ctx->nb-row0), c.finalize) ;
    if(rc) return rc ;
    if(e1) cudaEventRecord(e1, ctx->stream) ;
    ctx->stats.spmv_launches++ ;
    if(c.smoothing) ctx->stats.smoothing_spmv++ ;
    return AMIE_B200_OK ;
}


// one launch of the selection amie_b200_spmv_resident asks for (the option "spmv_variant" of the shipped dispatch was
// set by the caller; variant >= 100 = the in-solve form, fused p.q)
int launch_spmv_variant(amie_b200_ctx * ctx, const SpmvCall & c, int variant, bool insolve)
{
    // (the tuning configurations of round 1 -- stage counts, capacities, warps per tile -- were measured and are gone;
    // their numbers are in profiles/r01_notes.md)
    (void)variant ; (void)insolve ;
    return launch_spmv(ctx, c) ;
}
