// batch_loop.cuh -- the speculative iteration loop shared by the two solvers.
//
// Iterations are queued in batches without waiting for their outcome; every kernel returns at once when the
// device-side `stop` flag is set (krylov_scalars.cuh).  After each batch a copy of the state is sent to pinned
// host memory; two polls are kept in flight so the device never waits for the host.
// On small systems the inner loop is launch-bound: the batch is then captured ONCE into a CUDA graph
// (kernel arguments are pointers into the context, identical from solve to solve) and replayed.
#pragma once
#include "launch.cuh"

// what a captured batch depends on besides the arguments checked below: the SpMV kernel selection
static inline int graph_variant_key(const amie_b200_ctx * ctx) { return ctx->opt_variant ; }

template<typename QueueOne>
static int run_iteration_batches(amie_b200_ctx * ctx, amie_b200_ctx::GraphSlot & slot, bool use_graph, int batch,
                                 int precond, uint64_t rowstart, uint64_t colstart, int kernels_per_iter, int spmv_per_iter,
                                 QueueOne queue_one)
{
    if(use_graph)
    {
        const bool hit = slot.exec && slot.batch == batch && slot.precond == precond && slot.variant == graph_variant_key(ctx)
                         && slot.rowstart == rowstart && slot.colstart == colstart && slot.alloc_gen == ctx->alloc_gen ;
        if(!hit)
        {
            if(slot.exec) { cudaGraphExecDestroy(slot.exec) ; slot.exec = nullptr ; }
            const int saved_t = ctx->opt_time_spmv ;
            const uint64_t kl = ctx->stats.kernel_launches, sl = ctx->stats.spmv_launches ;
            ctx->opt_time_spmv = 0 ;                       // event pairs are not graph material
            cudaGraph_t graph = nullptr ;
            CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal)) ;
            int qrc = AMIE_B200_OK ;
            for(int i = 0 ; i < batch && !qrc ; i++) qrc = queue_one() ;
            cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph) ;
            if(qrc) { if(graph) cudaGraphDestroy(graph) ; ctx->opt_time_spmv = saved_t ; return qrc ; }
            ctx->opt_time_spmv = saved_t ;
            ctx->stats.kernel_launches = kl ; ctx->stats.spmv_launches = sl ;     // counted per replay below
            CUDA_TRY(ctx, e) ;
            e = cudaGraphInstantiate(&slot.exec, graph, 0) ;
            cudaGraphDestroy(graph) ;
            CUDA_TRY(ctx, e) ;
            slot.batch = batch ; slot.precond = precond ; slot.variant = graph_variant_key(ctx) ;
            slot.rowstart = rowstart ; slot.colstart = colstart ; slot.alloc_gen = ctx->alloc_gen ;
        }
    }
    int pslot = 0, pending = 0 ;
    bool stopped = false ;
    // safety net: the device-side test ends the loop after at most n_limit iterations (krylov_scalars.cuh)
    const uint64_t batches_cap = (ctx->nb_global*(uint64_t)ctx->S*4+16)/(uint64_t)batch+4 ;
    uint64_t batches = 0 ;
    while(!stopped)
    {
        if(++batches > batches_cap) { ctx->set_error("iteration loop: the device never reported the end of the loop") ; return AMIE_B200_ERR_CUDA ; }
        if(use_graph)
        {
            CUDA_TRY(ctx, cudaGraphLaunch(slot.exec, ctx->stream)) ;
            ctx->stats.kernel_launches += (uint64_t)kernels_per_iter*batch ;
            ctx->stats.spmv_launches += (uint64_t)spmv_per_iter*batch ;
        }
        else
            for(int i = 0 ; i < batch ; i++)
            {
                // a failed launch or exchange must end the loop here: nothing would ever set `stop`
                const int qrc = queue_one() ;
                if(qrc) return qrc ;
            }
        CUDA_TRY(ctx, cudaPeekAtLastError()) ;
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->st_host+pslot, ctx->st, sizeof(KrylovState), cudaMemcpyDeviceToHost, ctx->stream)) ;
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_poll[pslot], ctx->stream)) ;
        pending++ ;
        if(pending == 2)
        {
            const int old = pslot^1 ;
            CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_poll[old])) ;
            pending-- ;
            if(ctx->st_host[old].stop) stopped = true ;
        }
        pslot ^= 1 ;
    }
    return AMIE_B200_OK ;
}

// launch-bound when one iteration moves less than ~0.5 GB (< ~100 us of HBM time)
static inline bool want_graph(const amie_b200_ctx * ctx, double iter_bytes)
{
    if(ctx->dist || ctx->opt_time_spmv) return false ;
    if(ctx->opt_graph >= 0) return ctx->opt_graph != 0 ;
    return iter_bytes < 0.5e9 ;
}
