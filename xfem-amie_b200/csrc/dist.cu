// dist.cu -- row-partitioned execution over the GPUs of one NVSwitch box.
//
// One process (and one amie_b200_ctx) per GPU.  Rank r owns the contiguous block rows
// [bounds[r], bounds[r+1]).  Local column numbering: owned columns first (c - r0), then the halo
// (the distinct off-range columns the local rows reference, ascending = grouped by owner rank),
// so every SpMV input vector is  [ owned N_local | halo tail ]  and a peer's contribution lands
// in one contiguous slice of the tail -- no unpack kernel.
//
// Per SpMV:   pack owned entries the peers need -> grouped ncclSend/ncclRecv on a second stream,
//             overlapped with the SpMV of the interior rows (the longest run of rows that touch no
//             halo column; for slab partitions of a lexicographic mesh that is everything but one
//             node plane per side) -> boundary rows after the exchange.
// Per fused reduction: every rank's last block stores its partial sums; one ncclAllReduce of 2
//             doubles; a one-thread kernel then runs the same scalar recurrence / loop test
//             (krylov_scalars.cuh) on every rank, so all ranks take identical decisions and the
//             speculative batches of iterations stay matched across ranks.
//
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy torch already loaded when the caller
// is bench.py, else the system one) so that the single-GPU library has no NCCL dependency.
#include "launch.cuh"
#include "synth.h"
#include "dist.h"
#include "group.h"
#include <nccl.h>
#include <dlfcn.h>
#include <thrust/device_ptr.h>
#include <thrust/copy.h>
#include <thrust/sort.h>
#include <thrust/unique.h>
#include <thrust/execution_policy.h>
#include <algorithm>
#include <vector>

namespace {

struct NcclApi
{
    void * handle = nullptr ;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr ;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr ;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr ;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr ;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr ;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr ;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr ;
    ncclResult_t (*GroupStart)() = nullptr ;
    ncclResult_t (*GroupEnd)() = nullptr ;
    const char * (*GetErrorString)(ncclResult_t) = nullptr ;
} ;

NcclApi g_nccl ;

bool load_nccl(std::string & err)
{
    if(g_nccl.handle) return true ;
    // AMIE_B200_NCCL_LIB names the copy to use; a process that also loads another NCCL user (torch) must make
    // both resolve to the SAME libnccl.so.2 -- ld.so shares objects by SONAME, whichever is loaded first wins.
    const char * names[] = { getenv("AMIE_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so" } ;
    for(const char * n : names)
    {
        if(!n || !*n) continue ;
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL) ;
        if(g_nccl.handle) break ;
    }
    if(!g_nccl.handle) { err = std::string("dlopen libnccl.so.2: ")+dlerror() ; return false ; }
#define NCCL_SYM(field, name) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.handle, name)) ; \
    if(!g_nccl.field) { err = std::string("dlsym ")+name ; g_nccl.handle = nullptr ; return false ; }
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(AllReduce, "ncclAllReduce")
    NCCL_SYM(AllGather, "ncclAllGather")
    NCCL_SYM(Send, "ncclSend")
    NCCL_SYM(Recv, "ncclRecv")
    NCCL_SYM(GroupStart, "ncclGroupStart")
    NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
    return true ;
}

#define NCCL_TRY(ctx, expr) do { ncclResult_t _r = (expr) ; if(_r != ncclSuccess) { \
        (ctx)->set_error(std::string(#expr)+": "+g_nccl.GetErrorString(_r)) ; return AMIE_B200_ERR_NCCL ; } } while(0)

struct OffRange
{
    uint32_t r0, r1 ;
    __host__ __device__ bool operator()(uint32_t c) const { return c < r0 || c >= r1 ; }
} ;

// global block column -> local: owned c - r0 ; halo nb_local + position in the sorted halo list
__global__ void k_remap_cols(uint32_t * col, uint64_t nnzb, uint32_t r0, uint32_t r1, const uint32_t * halo, uint32_t nhalo, uint32_t nb_local)
{
    for(uint64_t k = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; k < nnzb ; k += (uint64_t)gridDim.x*blockDim.x)
    {
        const uint32_t c = col[k] ;
        if(c >= r0 && c < r1) col[k] = c-r0 ;
        else
        {
            uint32_t lo = 0, hi = nhalo ;
            while(lo < hi)
            {
                const uint32_t mid = lo+((hi-lo) >> 1) ;
                if(halo[mid] < c) lo = mid+1 ; else hi = mid ;
            }
            col[k] = nb_local+lo ;
        }
    }
}

__global__ void k_row_touches_halo(const uint32_t * rowptr, const uint32_t * col, uint32_t nb_local, unsigned char * flag)
{
    for(uint64_t r = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; r < nb_local ; r += (uint64_t)gridDim.x*blockDim.x)
    {
        unsigned char f = 0 ;
        for(uint32_t k = rowptr[r] ; k < rowptr[r+1] ; k++)
            if(col[k] >= nb_local) f = 1 ;
        flag[r] = f ;
    }
}

template<int S>
__global__ void k_pack(const double * v, const uint32_t * idx, uint64_t n, double * out)
{
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < n*S ; i += (uint64_t)gridDim.x*blockDim.x)
    {
        const uint64_t k = i/S ;
        out[i] = v[(uint64_t)idx[k]*S+(i-k*S)] ;
    }
}

// a SpMV whose first row range is empty (rowstart beyond the interior rows): red_local = 0 before the others add to it
__global__ void k_reset_red(KrylovState * st, int check_stop)
{
    if(check_stop && st->stop) return ;
    st->red_local[0] = 0. ;
    st->red_local[1] = 0. ;
}

__global__ void k_finalize(KrylovState * st, int kind)
{
    if(kind != FIN_STORE && st->stop) return ;      // mirrors the early return of the fused kernels
    krylov_finalize(st, kind, st->red_global[0], st->red_global[1]) ;
}

// ------------------------------------------------------------------ NVLink peer-memory transport
// Instead of NCCL calls on the critical path, kernels store straight into the neighbours' memory
// (cudaIpc-mapped) and signal with sequence-numbered flags; consumers spin on their OWN memory.
//   halo   : k_halo_push writes this rank's boundary entries of the SpMV input vector directly into the halo tail
//            of the same vector on each neighbour, then raises halo_flag[me] there; k_halo_wait (one warp) spins
//            until every neighbour's flag reached the current push number; the boundary rows follow in-stream.
//   reduce : k_finalize_peer writes (partial0, partial1, seq) into mailbox slot [seq&1][me] of EVERY rank, spins
//            until its own mailbox holds `world` entries of this seq, adds them in rank order (same bits on every
//            rank) and runs the scalar step.  Two slots suffice because a rank cannot be two executed reductions
//            ahead of another (each one needs everybody's contribution).
// Sequence numbers count EXECUTED operations (device-side counters that outlive the per-restart KrylovState):
// kernels that return early after `stop` do so on every rank alike.
#define PEER_MAX 8
#define SYNC_FLAG_OFF 0
#define SYNC_MBOX_OFF 512
#define SYNC_BYTES 8192

struct Mbox { double v0, v1 ; unsigned long long seq ; unsigned long long pad ; } ;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long * p)
{
    unsigned long long v ;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory") ;
    return v ;
}
__device__ __forceinline__ void st_release_sys(unsigned long long * p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory") ;
}

struct PushArgs
{
    const double * v ;
    const uint32_t * idx ;
    int npeers ;
    int S ;
    uint64_t begin[PEER_MAX], end[PEER_MAX] ;       // block ranges inside idx
    double * dst[PEER_MAX] ;                         // neighbour's vector + its owned length + my slice of its halo
    unsigned long long * flag[PEER_MAX] ;            // neighbour's halo_flag[me]
    KrylovState * st ;
    unsigned int * ticket ;
    unsigned long long * counters ;                  // [0] executed reductions, [1] executed halo pushes
    int check_stop ;
} ;

__global__ void __launch_bounds__(256) k_halo_push(PushArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    __shared__ bool is_last ;
    const uint64_t total = a.end[a.npeers-1] ;
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < total*a.S ; i += (uint64_t)gridDim.x*blockDim.x)
    {
        const uint64_t k = i/a.S ;
        const int c = (int)(i-k*a.S) ;
        int q = 0 ;
        #pragma unroll
        for(int t = 1 ; t < PEER_MAX ; t++) if(t < a.npeers && k >= a.begin[t]) q = t ;
        a.dst[q][(k-a.begin[q])*a.S+c] = a.v[(uint64_t)a.idx[k]*a.S+c] ;
    }
    __threadfence_system() ;
    __syncthreads() ;
    if(threadIdx.x == 0)
    {
        const unsigned int t = atomicAdd(a.ticket, 1u) ;
        is_last = (t == gridDim.x-1) ;
    }
    __syncthreads() ;
    if(!is_last) return ;
    __threadfence_system() ;
    if(threadIdx.x == 0)
    {
        *a.ticket = 0u ;
        const unsigned long long seq = a.counters[1]+1ull ;
        a.counters[1] = seq ;
        for(int q = 0 ; q < a.npeers ; q++) st_release_sys(a.flag[q], seq) ;
    }
}

struct WaitArgs
{
    int npeers ;
    const unsigned long long * flag[PEER_MAX] ;      // my halo_flag[neighbour]
    KrylovState * st ;
    const unsigned long long * counters ;
    int check_stop ;
} ;

__global__ void k_halo_wait(WaitArgs a)
{
    if(a.check_stop && a.st->stop) return ;
    const unsigned long long seq = a.counters[1] ;
    if((int)threadIdx.x < a.npeers)
        while(ld_acquire_sys(a.flag[threadIdx.x]) < seq) { }
    __threadfence_system() ;
}

struct ReduceArgs
{
    int world, rank ;
    Mbox * mbox[PEER_MAX] ;                          // mailbox of every rank (own included), [2][PEER_MAX]
    KrylovState * st ;
    unsigned long long * counters ;
    int kind ;
} ;

__global__ void k_finalize_peer(ReduceArgs a)
{
    if(a.kind != FIN_STORE && a.kind != FIN_DEFER_SET && a.st->stop) return ;
    const int lane = threadIdx.x ;
    const unsigned long long seq = a.counters[0]+1ull ;
    const int slot = (int)(seq & 1ull) ;
    if(lane < a.world)
    {
        Mbox * m = a.mbox[lane]+slot*PEER_MAX+a.rank ;
        m->v0 = a.st->red_local[0] ;
        m->v1 = a.st->red_local[1] ;
        __threadfence_system() ;
        st_release_sys(&m->seq, seq) ;
    }
    double v0 = 0., v1 = 0. ;
    if(lane < a.world)
    {
        Mbox * m = a.mbox[a.rank]+slot*PEER_MAX+lane ;
        while(ld_acquire_sys(&m->seq) != seq) { }
        v0 = *(volatile double *)&m->v0 ;
        v1 = *(volatile double *)&m->v1 ;
    }
    // rank order, the same on every rank
    double s0 = 0., s1 = 0. ;
    for(int q = 0 ; q < a.world ; q++)
    {
        s0 += __shfl_sync(0xffffffffu, v0, q) ;
        s1 += __shfl_sync(0xffffffffu, v1, q) ;
    }
    if(lane == 0)
    {
        a.counters[0] = seq ;
        a.st->red_global[0] = s0 ;
        a.st->red_global[1] = s1 ;
        if(a.kind != FIN_DEFER_SET) krylov_finalize(a.st, a.kind, s0, s1) ;      // FIN_DEFER_SET = barrier only
    }
}

// diagonal of a row whose (remapped) column indices are no longer sorted: linear scan
template<int S>
__global__ void k_inverse_diagonal_linear(const uint32_t * rowptr, const uint32_t * col, const double * vals, uint64_t nrows, double * d)
{
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < nrows*S ; i += (uint64_t)gridDim.x*blockDim.x)
    {
        const uint32_t row = (uint32_t)(i/S) ;
        const int m = (int)(i-(uint64_t)row*S) ;
        double v = 0. ;
        for(uint32_t k = rowptr[row] ; k < rowptr[row+1] ; k++)
            if(col[k] == row) { v = vals[(size_t)k*S*S+m*S+m] ; break ; }
        d[i] = fabs(v) > 1e-12 ? 1./v : 0. ;
    }
}

}

struct DistPeer
{
    int rank ;
    uint64_t recv_off, recv_cnt ;      // block columns, offset inside the halo tail
    uint64_t send_off, send_cnt ;      // block columns, offset inside send_idx / sendbuf
} ;

struct DistState
{
    int rank = 0, world = 1 ;
    ncclComm_t comm = nullptr ;
    cudaStream_t comm_stream = nullptr ;
    cudaEvent_t ev_pack = nullptr, ev_comm = nullptr ;
    std::vector<uint64_t> bounds ;
    uint64_t nhalo = 0 ;
    uint64_t nsend = 0 ;
    std::vector<DistPeer> peers ;
    uint32_t * send_idx = nullptr ;     // device: local block rows to pack
    double * sendbuf = nullptr ;        // device
    uint32_t int_a = 0, int_b = 0 ;     // interior block rows [a, b)
    double * scratch = nullptr ;        // device, 2 doubles (max reductions)
    std::vector<long long> need_all ;   // need_all[i*world+j] = halo block columns rank i receives from rank j

    // ---- peer-memory transport (NVLink loads/stores through cudaIpc mappings)
    bool peer_on = false ;
    unsigned char * sync = nullptr ;                 // this rank's flags + mailboxes
    std::vector<unsigned char *> sync_of ;           // [world]: every rank's sync buffer as seen from here
    std::vector<void *> ipc_opened ;                 // mappings to close
    struct VecMap { const double * base ; uint64_t gen ; std::vector<double *> of_peer ; } ;
    std::vector<VecMap> vec_maps ;                   // SpMV input vectors already exchanged
    unsigned int * push_ticket = nullptr ;
    unsigned long long * counters = nullptr ;        // device: executed reductions / halo pushes (never reset)

    // ---- in-process group (group.h): set-up collectives through shared host memory instead of NCCL / cudaIpc
    LocalGroup * local = nullptr ;
    std::vector<uint32_t> halo_host ;                // this rank's halo list while the send lists are being built
    double * cs_save = nullptr ;                     // colstart > 0: the owned entries in front of colstart, parked during a SpMV
    uint64_t cs_save_len = 0 ;
} ;

// AMIE_B200_TRACE=1: stage markers of the partitioned paths on stderr (debugging hangs between parts)
static bool trace_on() { static int v = -1 ; if(v < 0) { const char * e = getenv("AMIE_B200_TRACE") ; v = (e && atoi(e)) ? 1 : 0 ; } return v == 1 ; }
#define TRACE(ctx, ...) do { if(trace_on()) { fprintf(stderr, "[amie_b200 part %d dev %d] ", (ctx)->dist ? (ctx)->dist->rank : -1, (ctx)->device) ; fprintf(stderr, __VA_ARGS__) ; fprintf(stderr, "\n") ; fflush(stderr) ; } } while(0)

// ---- collectives used at set-up time, on either plumbing -------------------------------------------------------
#define GROUP_TRY(ctx, ok) do { if(!(ok)) { (ctx)->set_error("group: a collective step was abandoned (another device failed)") ; return AMIE_B200_ERR_STATE ; } } while(0)

// op: 0 min, 1 max
static int comm_allreduce_scalar(amie_b200_ctx * ctx, double * value, int op)
{
    DistState * d = ctx->dist ;
    if(d->local)
    {
        LocalGroup * g = d->local ;
        g->dbl[d->rank] = *value ;
        GROUP_TRY(ctx, g->barrier()) ;
        double m = g->dbl[0] ;
        for(int r = 1 ; r < d->world ; r++)
        {
            const double y = g->dbl[r] ;
            if(op == 0 ? (y < m) : (y > m || y != y)) m = y ;
        }
        GROUP_TRY(ctx, g->barrier()) ;
        *value = m ;
        return AMIE_B200_OK ;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(d->scratch, value, sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    NCCL_TRY(ctx, g_nccl.AllReduce(d->scratch, d->scratch+1, 1, ncclDouble, op == 0 ? ncclMin : ncclMax, d->comm, ctx->stream)) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(value, d->scratch+1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return AMIE_B200_OK ;
}

int dist_world(const amie_b200_ctx * ctx) { return ctx->dist ? ctx->dist->world : 1 ; }
int dist_rank(const amie_b200_ctx * ctx) { return ctx->dist ? ctx->dist->rank : 0 ; }

uint64_t dist_local_rowstart(const amie_b200_ctx * ctx, uint64_t rowstart_global)
{
    const uint64_t base = ctx->row_base*(uint64_t)ctx->S ;
    if(rowstart_global <= base) return 0 ;
    return std::min<uint64_t>(rowstart_global-base, ctx->N) ;
}

void dist_destroy(amie_b200_ctx * ctx)
{
    DistState * d = ctx->dist ;
    if(!d) return ;
    for(void * m : d->ipc_opened) cudaIpcCloseMemHandle(m) ;
    if(d->sync) cudaFree(d->sync) ;
    if(d->push_ticket) cudaFree(d->push_ticket) ;
    if(d->counters) cudaFree(d->counters) ;
    if(d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm) ;
    if(d->send_idx) cudaFree(d->send_idx) ;
    if(d->sendbuf) cudaFree(d->sendbuf) ;
    if(d->cs_save) cudaFree(d->cs_save) ;
    if(d->scratch) cudaFree(d->scratch) ;
    if(d->ev_pack) cudaEventDestroy(d->ev_pack) ;
    if(d->ev_comm) cudaEventDestroy(d->ev_comm) ;
    if(d->comm_stream) cudaStreamDestroy(d->comm_stream) ;
    delete d ;
    ctx->dist = nullptr ;
}

// sum st->red_local over the ranks, then run the scalar step `kind` on every rank
static void launch_finalize_peer(amie_b200_ctx * ctx, int kind)
{
    DistState * d = ctx->dist ;
    ReduceArgs ra ;
    ra.world = d->world ; ra.rank = d->rank ; ra.st = ctx->st ; ra.kind = kind ; ra.counters = d->counters ;
    for(int r = 0 ; r < d->world ; r++) ra.mbox[r] = reinterpret_cast<Mbox *>(d->sync_of[r]+SYNC_MBOX_OFF) ;
    k_finalize_peer<<<1, 32, 0, ctx->stream>>>(ra) ;
    ctx->stats.kernel_launches++ ;
}

int dist_finalize(amie_b200_ctx * ctx, int kind)
{
    DistState * d = ctx->dist ;
    if(d->peer_on)
    {
        launch_finalize_peer(ctx, kind) ;
        return AMIE_B200_OK ;
    }
    NCCL_TRY(ctx, g_nccl.AllReduce(ctx->st->red_local, ctx->st->red_global, 2, ncclDouble, ncclSum, d->comm, ctx->stream)) ;
    k_finalize<<<1, 1, 0, ctx->stream>>>(ctx->st, kind) ;
    ctx->stats.kernel_launches++ ;
    return AMIE_B200_OK ;
}

int dist_allreduce_max(amie_b200_ctx * ctx, double * value)
{
    return comm_allreduce_scalar(ctx, value, 1) ;
}

// In-process groups: every part has finished what it was doing on the HOST side (allocations above all) before any
// part goes on to queue kernels that wait for the others on the device.  Nothing to do between processes.
int dist_host_barrier(amie_b200_ctx * ctx)
{
    DistState * d = ctx->dist ;
    if(!d || !d->local) return AMIE_B200_OK ;
    GROUP_TRY(ctx, d->local->barrier()) ;
    return AMIE_B200_OK ;
}

int dist_inverse_diagonal(amie_b200_ctx * ctx)
{
    int grid = vec_grid(ctx, ctx->N) ;
    if(ctx->S == 3) k_inverse_diagonal_linear<3><<<grid, 256, 0, ctx->stream>>>(ctx->rowptr, ctx->col, ctx->vals, ctx->nb, ctx->dinv) ;
    else            k_inverse_diagonal_linear<2><<<grid, 256, 0, ctx->stream>>>(ctx->rowptr, ctx->col, ctx->vals, ctx->nb, ctx->dinv) ;
    CUDA_TRY(ctx, cudaGetLastError()) ;
    return AMIE_B200_OK ;
}

// all ranks exchange the cudaIpc handle of one allocation each; out[r] = rank r's pointer as seen from here
// (nullptr for ranks that are not opened; `only_peers`: open the neighbours' mappings only)
static int ipc_exchange(amie_b200_ctx * ctx, void * mine, bool only_peers, std::vector<void *> & out)
{
    DistState * d = ctx->dist ;
    if(d->local)
    {
        // one address space, peer access enabled at group creation: the pointers themselves travel
        LocalGroup * g = d->local ;
        (void)only_peers ;
        g->ptr[d->rank] = mine ;
        GROUP_TRY(ctx, g->barrier()) ;
        out.assign(g->ptr, g->ptr+d->world) ;
        GROUP_TRY(ctx, g->barrier()) ;
        return AMIE_B200_OK ;
    }
    cudaIpcMemHandle_t h ;
    CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, mine)) ;
    unsigned char * dbuf = nullptr ;
    CUDA_TRY(ctx, cudaMalloc(&dbuf, (size_t)(d->world+1)*sizeof(h))) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(dbuf, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream)) ;
    NCCL_TRY(ctx, g_nccl.AllGather(dbuf, dbuf+sizeof(h), sizeof(h), ncclChar, d->comm, ctx->stream)) ;
    std::vector<cudaIpcMemHandle_t> all(d->world) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(all.data(), dbuf+sizeof(h), (size_t)d->world*sizeof(h), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    cudaFree(dbuf) ;
    out.assign(d->world, nullptr) ;
    out[d->rank] = mine ;
    for(int r = 0 ; r < d->world ; r++)
    {
        if(r == d->rank) continue ;
        if(only_peers)
        {
            bool is_peer = false ;
            for(const DistPeer & p : d->peers) if(p.rank == r) is_peer = true ;
            if(!is_peer) continue ;
        }
        void * ptr = nullptr ;
        CUDA_TRY(ctx, cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess)) ;
        d->ipc_opened.push_back(ptr) ;
        out[r] = ptr ;
    }
    return AMIE_B200_OK ;
}

// flags + mailboxes of every rank mapped everywhere; called once the halo lists exist.
// Collective.  If ANY rank cannot map its peers (no P2P path, IPC disabled in the container ...) every rank
// falls back to the NCCL transport: the decision is agreed with one all-reduce.
static int peer_setup_local(amie_b200_ctx * ctx)
{
    DistState * d = ctx->dist ;
    if(d->sync) return AMIE_B200_OK ;
    CUDA_TRY(ctx, cudaMalloc(&d->sync, SYNC_BYTES)) ;
    CUDA_TRY(ctx, cudaMalloc(&d->push_ticket, sizeof(unsigned int))) ;
    CUDA_TRY(ctx, cudaMemset(d->push_ticket, 0, sizeof(unsigned int))) ;
    CUDA_TRY(ctx, cudaMalloc(&d->counters, 2*sizeof(unsigned long long))) ;
    CUDA_TRY(ctx, cudaMemset(d->counters, 0, 2*sizeof(unsigned long long))) ;
    CUDA_TRY(ctx, cudaMemset(d->sync, 0, SYNC_BYTES)) ;
    return AMIE_B200_OK ;
}

static int peer_setup(amie_b200_ctx * ctx)
{
    DistState * d = ctx->dist ;
    d->peer_on = false ;
    if(d->local)
    {
        // the only transport of an in-process group (group_create refused devices without a peer path)
        TRACE(ctx, "peer_setup") ;
        int rc = peer_setup_local(ctx) ;
        double ok = rc == AMIE_B200_OK ? 1. : 0. ;
        CUDA_TRY(ctx, cudaDeviceSynchronize()) ;           // the memsets above are done before anybody signals
        int rc2 = comm_allreduce_scalar(ctx, &ok, 0) ;
        if(rc2) return rc2 ;
        if(ok < 1.) return rc ? rc : AMIE_B200_ERR_CUDA ;
        std::vector<void *> out ;
        if((rc = ipc_exchange(ctx, d->sync, false, out))) return rc ;
        d->sync_of.resize(d->world) ;
        for(int r = 0 ; r < d->world ; r++) d->sync_of[r] = static_cast<unsigned char *>(out[r]) ;
        d->peer_on = true ;
        return AMIE_B200_OK ;
    }
    const char * e = getenv("AMIE_B200_TRANSPORT") ;
    const bool want = !(e && std::string(e) == "nccl") && d->world <= PEER_MAX && (int)d->peers.size() <= PEER_MAX ;
    double ok = want ? 1. : 0. ;
    if(want && peer_setup_local(ctx) != AMIE_B200_OK) ok = 0. ;
    // the handle exchange itself is collective and must run on every rank once any rank wants it
    CUDA_TRY(ctx, cudaMemcpyAsync(d->scratch, &ok, sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    NCCL_TRY(ctx, g_nccl.AllReduce(d->scratch, d->scratch+1, 1, ncclDouble, ncclMin, d->comm, ctx->stream)) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(&ok, d->scratch+1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    if(ok < 1.) return AMIE_B200_OK ;                      // NCCL transport on every rank
    if(d->sync_of.empty())
    {
        std::vector<void *> out ;
        double mapped = ipc_exchange(ctx, d->sync, false, out) == AMIE_B200_OK ? 1. : 0. ;
        cudaGetLastError() ;                               // a failed cudaIpcOpenMemHandle must not poison later calls
        CUDA_TRY(ctx, cudaMemcpyAsync(d->scratch, &mapped, sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
        NCCL_TRY(ctx, g_nccl.AllReduce(d->scratch, d->scratch+1, 1, ncclDouble, ncclMin, d->comm, ctx->stream)) ;
        CUDA_TRY(ctx, cudaMemcpyAsync(&mapped, d->scratch+1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
        if(mapped < 1.) { ctx->err.clear() ; return AMIE_B200_OK ; }
        d->sync_of.resize(d->world) ;
        for(int r = 0 ; r < d->world ; r++) d->sync_of[r] = static_cast<unsigned char *>(out[r]) ;
    }
    // (the all-reduces above also guarantee that every rank's memset finished before anybody signals)
    d->peer_on = true ;
    return AMIE_B200_OK ;
}

// neighbours' views of one SpMV input vector (exchanged collectively the first time the vector is used)
static int peer_vector(amie_b200_ctx * ctx, const double * base, const std::vector<double *> ** out)
{
    DistState * d = ctx->dist ;
    for(auto & m : d->vec_maps)
        if(m.base == base && m.gen == ctx->alloc_gen) { *out = &m.of_peer ; return AMIE_B200_OK ; }
    std::vector<void *> all ;
    TRACE(ctx, "peer_vector: first use of %p as a SpMV input", (const void *)base) ;
    int rc = ipc_exchange(ctx, const_cast<double *>(base), true, all) ;
    if(rc) return rc ;
    TRACE(ctx, "peer_vector: exchanged") ;
    DistState::VecMap m ;
    m.base = base ; m.gen = ctx->alloc_gen ;
    for(const DistPeer & p : d->peers) m.of_peer.push_back(static_cast<double *>(all[p.rank])) ;
    d->vec_maps.push_back(m) ;
    *out = &d->vec_maps.back().of_peer ;
    return AMIE_B200_OK ;
}

// y = A x on the local rows with the halo exchange of x overlapped with the interior rows
int dist_spmv(amie_b200_ctx * ctx, const SpmvCall & c)
{
    DistState * d = ctx->dist ;
    const int S = ctx->S ;
    double * xv = const_cast<double *>(c.x) ;
    const bool dot = c.dot != DOT_NONE ;
    // c.rowstart is LOCAL (dist_local_rowstart), c.colstart GLOBAL.  Skipping the block columns < colstart is the same
    // as multiplying by a vector whose entries in front of colstart are zero: every rank parks the owned part of that
    // prefix, zeroes it for the duration of this SpMV (the halo pushes then carry zeros too) and puts it back.
    // The local column numbering (owned | halo) is not ascending, so the kernels' lower_bound skip cannot be used.
    uint64_t cs_local = 0 ;
    if(c.colstart > d->bounds[d->rank]*S)
        cs_local = std::min<uint64_t>(c.colstart-d->bounds[d->rank]*S, ctx->N) ;
    if(cs_local)
    {
        // (the parking buffer is allocated with the structure: no allocation may happen here, where another part of
        // an in-process group can already be waiting on this one ON THE DEVICE -- cudaMalloc / cudaFree may synchronise
        // the device, and two parts may share one)
        if(d->cs_save_len < cs_local) { ctx->set_error("distributed SpMV: colstart buffer missing") ; return AMIE_B200_ERR_STATE ; }
        CUDA_TRY(ctx, cudaMemcpyAsync(d->cs_save, xv, cs_local*sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream)) ;
        CUDA_TRY(ctx, cudaMemsetAsync(xv, 0, cs_local*sizeof(double), ctx->stream)) ;
    }
    const uint32_t row_lo = (uint32_t)(c.rowstart/S) ;
    cudaEvent_t e0 = nullptr, e1 = nullptr ;
    if(ctx->opt_time_spmv && ctx->ev_used+2 <= ctx->ev_pool.size())
    {
        e0 = ctx->ev_pool[ctx->ev_used++] ;
        e1 = ctx->ev_pool[ctx->ev_used++] ;
        cudaEventRecord(e0, ctx->stream) ;
    }
    const bool peer = d->peer_on && !d->peers.empty() ;
    static int traced = 0 ;
    const bool tr = trace_on() && traced < 12 ;
    if(tr) { traced++ ; TRACE(ctx, "dist_spmv: x %p dot %d rowstart %llu colstart %llu peer %d", (const void *)c.x, c.dot, (unsigned long long)c.rowstart, (unsigned long long)c.colstart, (int)peer) ; }
    if(peer)
    {
        // stores straight into the neighbours' halo tails over NVLink, then a flag; no NCCL call
        const std::vector<double *> * of_peer = nullptr ;
        int prc = peer_vector(ctx, xv, &of_peer) ;
        if(prc) return prc ;
        PushArgs pa ;
        pa.v = xv ; pa.idx = d->send_idx ; pa.S = S ; pa.st = ctx->st ; pa.ticket = d->push_ticket ; pa.counters = d->counters ; pa.check_stop = c.check_stop ;
        pa.npeers = (int)d->peers.size() ;
        for(int t = 0 ; t < pa.npeers ; t++)
        {
            const DistPeer & p = d->peers[t] ;
            const int q = p.rank ;
            uint64_t off_q = 0 ;             // where my columns start inside q's halo
            for(int j = 0 ; j < d->rank ; j++) if(j != q) off_q += (uint64_t)d->need_all[(size_t)q*d->world+j] ;
            pa.begin[t] = p.send_off ; pa.end[t] = p.send_off+p.send_cnt ;
            pa.dst[t] = (*of_peer)[t]+(d->bounds[q+1]-d->bounds[q])*S+off_q*S ;
            pa.flag[t] = reinterpret_cast<unsigned long long *>(d->sync_of[q]+SYNC_FLAG_OFF)+d->rank ;
        }
        for(int t = pa.npeers ; t < PEER_MAX ; t++) { pa.begin[t] = pa.end[t] = pa.end[pa.npeers-1] ; pa.dst[t] = nullptr ; pa.flag[t] = nullptr ; }
        k_halo_push<<<vec_grid(ctx, std::max<uint64_t>(d->nsend*S, 1)), 256, 0, ctx->stream>>>(pa) ;
        ctx->stats.kernel_launches++ ;
    }
    else if(!d->peers.empty())
    {
        if(d->nsend)
        {
            if(S == 3) k_pack<3><<<vec_grid(ctx, d->nsend*3), 256, 0, ctx->stream>>>(xv, d->send_idx, d->nsend, d->sendbuf) ;
            else       k_pack<2><<<vec_grid(ctx, d->nsend*2), 256, 0, ctx->stream>>>(xv, d->send_idx, d->nsend, d->sendbuf) ;
            ctx->stats.kernel_launches++ ;
        }
        CUDA_TRY(ctx, cudaEventRecord(d->ev_pack, ctx->stream)) ;
        CUDA_TRY(ctx, cudaStreamWaitEvent(d->comm_stream, d->ev_pack, 0)) ;
        NCCL_TRY(ctx, g_nccl.GroupStart()) ;
        for(const DistPeer & p : d->peers)
        {
            if(p.send_cnt) NCCL_TRY(ctx, g_nccl.Send(d->sendbuf+p.send_off*S, p.send_cnt*S, ncclDouble, p.rank, d->comm, d->comm_stream)) ;
            if(p.recv_cnt) NCCL_TRY(ctx, g_nccl.Recv(xv+ctx->N+p.recv_off*S, p.recv_cnt*S, ncclDouble, p.rank, d->comm, d->comm_stream)) ;
        }
        NCCL_TRY(ctx, g_nccl.GroupEnd()) ;
        CUDA_TRY(ctx, cudaEventRecord(d->ev_comm, d->comm_stream)) ;
    }
    int rc ;
    bool first = true ;
    SpmvCall cl = c ;
    cl.colstart = 0 ;                                                     // done through the zeroed prefix above
    auto part = [&](uint32_t a, uint32_t b) -> int
    {
        if(a < row_lo) a = row_lo ;                                       // rows < rowstart keep what they hold
        if(b <= a) return AMIE_B200_OK ;
        int r = launch_spmv_range(ctx, cl, a, b-a, dot ? (first ? FIN_DEFER_SET : FIN_DEFER_ADD) : FIN_STORE) ;
        first = false ;
        return r ;
    } ;
    if((rc = part(d->int_a, d->int_b))) return rc ;                       // interior rows: no halo column
    if(peer)
    {
        WaitArgs wa ;
        wa.npeers = (int)d->peers.size() ; wa.st = ctx->st ; wa.counters = d->counters ; wa.check_stop = c.check_stop ;
        for(int t = 0 ; t < wa.npeers ; t++)
            wa.flag[t] = reinterpret_cast<const unsigned long long *>(d->sync+SYNC_FLAG_OFF)+d->peers[t].rank ;
        for(int t = wa.npeers ; t < PEER_MAX ; t++) wa.flag[t] = nullptr ;
        k_halo_wait<<<1, 32, 0, ctx->stream>>>(wa) ;
        ctx->stats.kernel_launches++ ;
    }
    else if(!d->peers.empty()) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, d->ev_comm, 0)) ;
    if((rc = part(0, d->int_a))) return rc ;                              // boundary rows
    if((rc = part(d->int_b, (uint32_t)ctx->nb))) return rc ;
    if(dot && first)
    {
        // no local row at or after rowstart: this rank adds nothing to the sums
        k_reset_red<<<1, 1, 0, ctx->stream>>>(ctx->st, c.check_stop) ;
        ctx->stats.kernel_launches++ ;
    }
    if(dot && (rc = dist_finalize(ctx, c.finalize))) return rc ;
    // no reduction follows a plain SpMV: a mailbox round keeps a fast neighbour from overwriting this rank's halo
    // tail with its NEXT push while the boundary rows above still read it
    if(!dot && peer) launch_finalize_peer(ctx, FIN_DEFER_SET) ;
    if(cs_local) CUDA_TRY(ctx, cudaMemcpyAsync(xv, d->cs_save, cs_local*sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream)) ;
    if(tr) TRACE(ctx, "dist_spmv: queued") ;
    if(e1) cudaEventRecord(e1, ctx->stream) ;
    ctx->stats.spmv_launches++ ;
    if(c.smoothing) ctx->stats.smoothing_spmv++ ;
    return AMIE_B200_OK ;
}

// ctx->rowptr / col (GLOBAL block columns) / vals of the local rows are on the device: build the halo,
// renumber the columns, exchange the send lists, find the interior rows, size the vectors.
static int dist_finish_structure(amie_b200_ctx * ctx)
{
    DistState * d = ctx->dist ;
    const uint32_t r0 = (uint32_t)d->bounds[d->rank], r1 = (uint32_t)d->bounds[d->rank+1] ;
    const uint32_t nbl = r1-r0 ;
    cudaStream_t st = ctx->stream ;
    auto pol = thrust::cuda::par.on(st) ;

    TRACE(ctx, "finish_structure: rows [%u, %u), %llu blocks", r0, r1, (unsigned long long)ctx->nnzb) ;
    // ---- halo = sorted distinct off-range columns
    uint32_t * tmp = nullptr ;
    CUDA_TRY(ctx, cudaMalloc(&tmp, std::max<uint64_t>(ctx->nnzb, 1)*sizeof(uint32_t))) ;
    thrust::device_ptr<uint32_t> colp(ctx->col), tmpp(tmp) ;
    auto end1 = thrust::copy_if(pol, colp, colp+ctx->nnzb, tmpp, OffRange{r0, r1}) ;
    thrust::sort(pol, tmpp, end1) ;
    auto end2 = thrust::unique(pol, tmpp, end1) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(st)) ;
    d->nhalo = (uint64_t)(end2-tmpp) ;
    std::vector<uint32_t> halo(d->nhalo) ;
    CUDA_TRY(ctx, cudaMemcpy(halo.data(), tmp, d->nhalo*sizeof(uint32_t), cudaMemcpyDeviceToHost)) ;
    k_remap_cols<<<vec_grid(ctx, ctx->nnzb), 256, 0, st>>>(ctx->col, ctx->nnzb, r0, r1, tmp, (uint32_t)d->nhalo, nbl) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(st)) ;
    CUDA_TRY(ctx, cudaGetLastError()) ;
    // the halo list stays on the device: the value assembly and the elimination translate global node ids with it
    if(ctx->halo_glob) { cudaFree(ctx->halo_glob) ; ctx->halo_glob = nullptr ; }
    CUDA_TRY(ctx, cudaMalloc(&ctx->halo_glob, std::max<uint64_t>(d->nhalo, 1)*sizeof(uint32_t))) ;
    CUDA_TRY(ctx, cudaMemcpy(ctx->halo_glob, tmp, d->nhalo*sizeof(uint32_t), cudaMemcpyDeviceToDevice)) ;
    cudaFree(tmp) ;

    // ---- who owns which slice of the halo
    std::vector<long long> need(d->world, 0) ;            // block columns this rank needs from rank q
    std::vector<uint64_t> need_off(d->world+1, 0) ;
    for(uint64_t h = 0 ; h < d->nhalo ; h++)
    {
        int q = (int)(std::upper_bound(d->bounds.begin(), d->bounds.end(), (uint64_t)halo[h])-d->bounds.begin())-1 ;
        need[q]++ ;
    }
    for(int q = 0 ; q < d->world ; q++) need_off[q+1] = need_off[q]+(uint64_t)need[q] ;

    // ---- all ranks learn the whole need matrix; then the lists travel
    std::vector<long long> all((size_t)d->world*d->world) ;
    if(d->local)
    {
        LocalGroup * g = d->local ;
        g->ll[d->rank] = need ;
        d->halo_host = halo ;
        g->cptr[d->rank] = d ;
        GROUP_TRY(ctx, g->barrier()) ;
        for(int i = 0 ; i < d->world ; i++)
            for(int j = 0 ; j < d->world ; j++) all[(size_t)i*d->world+j] = g->ll[i][j] ;
    }
    else
    {
        long long * dneed = nullptr, * dall = nullptr ;
        CUDA_TRY(ctx, cudaMalloc(&dneed, d->world*sizeof(long long))) ;
        CUDA_TRY(ctx, cudaMalloc(&dall, (size_t)d->world*d->world*sizeof(long long))) ;
        CUDA_TRY(ctx, cudaMemcpyAsync(dneed, need.data(), d->world*sizeof(long long), cudaMemcpyHostToDevice, st)) ;
        NCCL_TRY(ctx, g_nccl.AllGather(dneed, dall, d->world, ncclInt64, d->comm, st)) ;
        CUDA_TRY(ctx, cudaMemcpyAsync(all.data(), dall, all.size()*sizeof(long long), cudaMemcpyDeviceToHost, st)) ;
        CUDA_TRY(ctx, cudaStreamSynchronize(st)) ;
        cudaFree(dneed) ; cudaFree(dall) ;
    }

    d->need_all = all ;
    d->peers.clear() ;
    d->nsend = 0 ;
    for(int q = 0 ; q < d->world ; q++)
    {
        if(q == d->rank) continue ;
        const uint64_t rc = (uint64_t)all[(size_t)d->rank*d->world+q] ;     // I need from q
        const uint64_t sc = (uint64_t)all[(size_t)q*d->world+d->rank] ;     // q needs from me
        if(rc == 0 && sc == 0) continue ;
        DistPeer p ;
        p.rank = q ; p.recv_off = need_off[q] ; p.recv_cnt = rc ; p.send_off = d->nsend ; p.send_cnt = sc ;
        d->nsend += sc ;
        d->peers.push_back(p) ;
    }
    if(d->send_idx) { cudaFree(d->send_idx) ; d->send_idx = nullptr ; }
    if(d->sendbuf) { cudaFree(d->sendbuf) ; d->sendbuf = nullptr ; }
    CUDA_TRY(ctx, cudaMalloc(&d->send_idx, std::max<uint64_t>(d->nsend, 1)*sizeof(uint32_t))) ;
    CUDA_TRY(ctx, cudaMalloc(&d->sendbuf, std::max<uint64_t>(d->nsend, 1)*ctx->S*sizeof(double))) ;
    std::vector<uint32_t> sidx(d->nsend) ;
    if(d->local)
    {
        // "I need these columns of yours": read straight out of the peers' halo lists (owner-grouped, ascending)
        LocalGroup * g = d->local ;
        for(const DistPeer & p : d->peers)
        {
            if(!p.send_cnt) continue ;
            const DistState * dq = static_cast<const DistState *>(g->cptr[p.rank]) ;
            uint64_t off = 0 ;
            for(int j = 0 ; j < d->rank ; j++) off += (uint64_t)all[(size_t)p.rank*d->world+j] ;
            std::copy(dq->halo_host.begin()+off, dq->halo_host.begin()+off+p.send_cnt, sidx.begin()+p.send_off) ;
        }
        GROUP_TRY(ctx, g->barrier()) ;                     // everybody is done reading everybody's list
        d->halo_host.clear() ; d->halo_host.shrink_to_fit() ;
    }
    else
    {
        uint32_t * dhalo = nullptr ;
        CUDA_TRY(ctx, cudaMalloc(&dhalo, std::max<uint64_t>(d->nhalo, 1)*sizeof(uint32_t))) ;
        CUDA_TRY(ctx, cudaMemcpyAsync(dhalo, halo.data(), d->nhalo*sizeof(uint32_t), cudaMemcpyHostToDevice, st)) ;
        NCCL_TRY(ctx, g_nccl.GroupStart()) ;
        for(const DistPeer & p : d->peers)
        {
            if(p.recv_cnt) NCCL_TRY(ctx, g_nccl.Send(dhalo+p.recv_off, p.recv_cnt, ncclUint32, p.rank, d->comm, st)) ;      // "I need these columns of yours"
            if(p.send_cnt) NCCL_TRY(ctx, g_nccl.Recv(d->send_idx+p.send_off, p.send_cnt, ncclUint32, p.rank, d->comm, st)) ;
        }
        NCCL_TRY(ctx, g_nccl.GroupEnd()) ;
        CUDA_TRY(ctx, cudaMemcpyAsync(sidx.data(), d->send_idx, d->nsend*sizeof(uint32_t), cudaMemcpyDeviceToHost, st)) ;
        CUDA_TRY(ctx, cudaStreamSynchronize(st)) ;
        cudaFree(dhalo) ;
    }
    // global column -> local row of mine
    for(auto & v : sidx)
    {
        if(v < r0 || v >= r1) { ctx->set_error("distributed structure: a peer asked for a column this rank does not own") ; return AMIE_B200_ERR_ARG ; }
        v -= r0 ;
    }
    CUDA_TRY(ctx, cudaMemcpy(d->send_idx, sidx.data(), d->nsend*sizeof(uint32_t), cudaMemcpyHostToDevice)) ;

    // ---- interior rows: the longest run of rows that reference no halo column
    unsigned char * dflag = nullptr ;
    CUDA_TRY(ctx, cudaMalloc(&dflag, std::max<uint32_t>(nbl, 1))) ;
    k_row_touches_halo<<<vec_grid(ctx, nbl), 256, 0, st>>>(ctx->rowptr, ctx->col, nbl, dflag) ;
    std::vector<unsigned char> flag(nbl) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(flag.data(), dflag, nbl, cudaMemcpyDeviceToHost, st)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(st)) ;
    cudaFree(dflag) ;
    uint32_t best_a = 0, best_b = 0, run = 0 ;
    for(uint32_t r = 0 ; r <= nbl ; r++)
    {
        if(r < nbl && !flag[r]) { run++ ; continue ; }
        if(run > best_b-best_a) { best_a = r-run ; best_b = r ; }
        run = 0 ;
    }
    d->int_a = best_a ; d->int_b = best_b ;

    TRACE(ctx, "finish_structure: halo %llu, send %llu, peers %zu, interior [%u, %u)", (unsigned long long)d->nhalo, (unsigned long long)d->nsend, d->peers.size(), d->int_a, d->int_b) ;
    ctx->ncols_local = (uint64_t)nbl+d->nhalo ;
    int arc = ctx_alloc_vectors(ctx) ;
    if(arc) return arc ;
    if(d->cs_save) { cudaFree(d->cs_save) ; d->cs_save = nullptr ; d->cs_save_len = 0 ; }
    CUDA_TRY(ctx, cudaMalloc(&d->cs_save, std::max<uint64_t>(ctx->N, 1)*sizeof(double))) ;     // colstart > 0: see dist_spmv
    d->cs_save_len = ctx->N ;
    // vectors may have been re-allocated: earlier mappings of them are stale
    d->vec_maps.clear() ;
    return peer_setup(ctx) ;
}

extern "C" {

int amie_b200_nccl_unique_id(void * id128_out)
{
    std::string err ;
    if(!id128_out || !load_nccl(err)) return AMIE_B200_ERR_NCCL ;
    ncclUniqueId id ;
    if(g_nccl.GetUniqueId(&id) != ncclSuccess) return AMIE_B200_ERR_NCCL ;
    memcpy(id128_out, &id, sizeof(id)) ;
    return AMIE_B200_OK ;
}

int amie_b200_dist_init(amie_b200_ctx * ctx, int rank, int world, const void * id128, const uint64_t * bounds)
{
    if(!ctx || !id128 || !bounds || world < 1 || rank < 0 || rank >= world) return AMIE_B200_ERR_ARG ;
    std::string err ;
    if(!load_nccl(err)) { ctx->set_error(err) ; return AMIE_B200_ERR_NCCL ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    dist_destroy(ctx) ;
    DistState * d = new DistState ;
    ctx->dist = d ;
    d->rank = rank ; d->world = world ;
    d->bounds.assign(bounds, bounds+world+1) ;
    ncclUniqueId id ;
    memcpy(&id, id128, sizeof(id)) ;
    NCCL_TRY(ctx, g_nccl.CommInitRank(&d->comm, world, id, rank)) ;
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&d->comm_stream, cudaStreamNonBlocking)) ;
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&d->ev_pack, cudaEventDisableTiming)) ;
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&d->ev_comm, cudaEventDisableTiming)) ;
    CUDA_TRY(ctx, cudaMalloc(&d->scratch, 2*sizeof(double))) ;
    return AMIE_B200_OK ;
}

int amie_b200_dist_set_structure(amie_b200_ctx * ctx, int stride, uint64_t nb_global,
                                 const uint32_t * row_size_local, const uint32_t * column_index_local, uint64_t nnzb_local)
{
    return dist_set_structure_local(ctx, stride, nb_global, row_size_local, column_index_local, nnzb_local) ;
}

}

// rank `rank` of an in-process group: no communicator, no second stream -- peer memory is the only transport
int dist_init_local(amie_b200_ctx * ctx, int rank, LocalGroup * g)
{
    if(!ctx || !g || rank < 0 || rank >= g->world) return AMIE_B200_ERR_ARG ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    dist_destroy(ctx) ;
    DistState * d = new DistState ;
    ctx->dist = d ;
    d->rank = rank ; d->world = g->world ;
    d->bounds = g->bounds ;
    d->local = g ;
    return AMIE_B200_OK ;
}

int dist_set_structure_local(amie_b200_ctx * ctx, int stride, uint64_t nb_global, const uint32_t * row_size_local,
                             const uint32_t * column_index_local, uint64_t nnzb_local)
{
    if(!ctx || !ctx->dist) return AMIE_B200_ERR_STATE ;
    DistState * d = ctx->dist ;
    if(d->bounds.back() != nb_global) { ctx->set_error("dist_set_structure: bounds do not cover nb_global") ; return AMIE_B200_ERR_ARG ; }
    const uint64_t nbl = d->bounds[d->rank+1]-d->bounds[d->rank] ;
    // upload through the single-device path with GLOBAL columns (range check against nb_global), then renumber
    int rc = ctx_set_structure(ctx, stride, nbl, row_size_local, column_index_local, nnzb_local, nb_global) ;
    if(rc) return rc ;
    ctx->nb_global = nb_global ;
    ctx->row_base = d->bounds[d->rank] ;
    return dist_finish_structure(ctx) ;
}

extern "C" {

int amie_b200_dist_synth_to_device(amie_b200_ctx * ctx, const amie_b200_synth * s)
{
    if(!ctx || !s || !ctx->dist) return AMIE_B200_ERR_STATE ;
    DistState * d = ctx->dist ;
    const SynthRecipe & R = *synth_recipe_of(s) ;
    const uint64_t nbg = synth_num_nodes(R) ;
    if(d->bounds.back() != nbg) { ctx->set_error("dist_synth_to_device: bounds do not cover the mesh") ; return AMIE_B200_ERR_ARG ; }
    const uint64_t r0 = d->bounds[d->rank], r1 = d->bounds[d->rank+1] ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    ctx_free_matrix(ctx) ;
    ctx->S = R.stride ; ctx->nb = r1-r0 ; ctx->nb_global = nbg ; ctx->row_base = r0 ; ctx->N = ctx->nb*R.stride ;
    ctx->ncols_local = ctx->nb ;
    // b lives in the vectors: allocate them for the owned part first (re-sized once the halo is known)
    double * btmp = nullptr ;
    CUDA_TRY(ctx, cudaMalloc(&btmp, std::max<uint64_t>(ctx->N, 1)*sizeof(double))) ;
    uint64_t nnzb = 0 ;
    int rc = synth_rows_to_device(ctx, R, r0, r1, &ctx->rowptr, &ctx->col, &ctx->vals, btmp, &nnzb) ;
    if(rc) { cudaFree(btmp) ; return rc ; }
    ctx->nnzb = nnzb ;
    ctx->have_structure = ctx->have_values = true ;
    rc = dist_finish_structure(ctx) ;
    if(rc) { cudaFree(btmp) ; return rc ; }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->b, btmp, ctx->N*sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    cudaFree(btmp) ;
    ctx->have_rhs = true ;
    return AMIE_B200_OK ;
}

int amie_b200_dist_transport(const amie_b200_ctx * ctx)
{
    if(!ctx || !ctx->dist) return AMIE_B200_ERR_STATE ;
    return ctx->dist->peer_on ? 1 : 0 ;
}

int amie_b200_dist_info(const amie_b200_ctx * ctx, uint64_t * nhalo_out, uint64_t * nsend_out, uint64_t * interior_rows_out, int * npeers_out)
{
    if(!ctx || !ctx->dist) return AMIE_B200_ERR_STATE ;
    if(nhalo_out) *nhalo_out = ctx->dist->nhalo ;
    if(nsend_out) *nsend_out = ctx->dist->nsend ;
    if(interior_rows_out) *interior_rows_out = ctx->dist->int_b-ctx->dist->int_a ;
    if(npeers_out) *npeers_out = (int)ctx->dist->peers.size() ;
    return AMIE_B200_OK ;
}

}
