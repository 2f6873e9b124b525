// api.cu -- C-ABI: context, matrix upload (K-Repack), SpMV / inverse-diagonal entry points, stats.
// The solvers are in solve_cg.cu and solve_bicg.cu.  No CPU fallback anywhere: every compute
// entry point needs a CUDA device and reports AMIE_B200_ERR_CUDA without one.
#include "launch.cuh"
#include "kernels_vec_block.cuh"
#include "group.h"
#include <cstring>
#include <cmath>
#include <cstdlib>
#include <vector>
#include <algorithm>

static std::string g_error ;

#define G_TRY(expr) do { cudaError_t _e = (expr) ; if(_e != cudaSuccess) { \
        g_error = std::string(#expr) + ": " + cudaGetErrorString(_e) ; return nullptr ; } } while(0)

template<typename T> static void dfree(T *& p) { if(p) cudaFree(p) ; p = nullptr ; }

void ctx_free_matrix(amie_b200_ctx * ctx)
{
    ctx->alloc_gen++ ;
    assembly_map_destroy(ctx) ;          // the gather lists index the stored blocks of this topology
    field_map_destroy(ctx) ;             // element data belongs to the topology too
    dfree(ctx->rowptr) ; dfree(ctx->col) ; dfree(ctx->vals) ; dfree(ctx->dinv) ; dfree(ctx->user_diag) ; dfree(ctx->block_to) ; dfree(ctx->halo_glob) ;
    ctx->dinv_len = 0 ;
    ctx->have_structure = ctx->have_values = ctx->dinv_valid = false ;
}

static void free_vectors(amie_b200_ctx * ctx)
{
    dfree(ctx->b) ; dfree(ctx->x) ; dfree(ctx->r) ; dfree(ctx->z) ; dfree(ctx->p) ; dfree(ctx->q) ;
    dfree(ctx->xc) ; dfree(ctx->rc) ; dfree(ctx->xmin) ;
    for(int i = 0 ; i < 8 ; i++) dfree(ctx->w[i]) ;
    ctx->vec_len = 0 ;
    ctx->have_rhs = false ;
}

static uint64_t device_bytes(const amie_b200_ctx * ctx)
{
    uint64_t nv = 9 ;
    for(int i = 0 ; i < 8 ; i++) if(ctx->w[i]) nv++ ;
    uint64_t bytes = ctx->vec_len*8*nv ;
    if(ctx->have_structure) bytes += (ctx->nb+1)*4+ctx->nnzb*4+ctx->nnzb*(uint64_t)(ctx->S*ctx->S)*8+ctx->N*8 ;
    if(ctx->user_diag) bytes += ctx->N*8 ;
    if(ctx->hist[0]) bytes += 2*ctx->hist_n*8 ;
    bytes += assembly_map_bytes(ctx)+field_map_bytes(ctx) ;       // rows next to the solve (assemble.cu, fields.cu)
    return bytes ;
}

int ctx_alloc_vectors(amie_b200_ctx * ctx)
{
    // SpMV inputs may be read up to ncols_local*S (owned rows + halo tail on a distributed context)
    uint64_t len = std::max<uint64_t>(ctx->N, ctx->ncols_local*ctx->S) ;
    if(len == 0) len = 1 ;
    if(ctx->vec_len == len) return AMIE_B200_OK ;
    free_vectors(ctx) ;
    ctx->alloc_gen++ ;
    double ** v[] = { &ctx->b, &ctx->x, &ctx->r, &ctx->z, &ctx->p, &ctx->q, &ctx->xc, &ctx->rc, &ctx->xmin } ;
    for(auto pp : v)
    {
        CUDA_TRY(ctx, cudaMalloc(pp, len*sizeof(double))) ;
        CUDA_TRY(ctx, cudaMemsetAsync(*pp, 0, len*sizeof(double), ctx->stream)) ;
    }
    ctx->vec_len = len ;
    return AMIE_B200_OK ;
}

int ctx_ensure_bicg_vectors(amie_b200_ctx * ctx)
{
    for(int i = 0 ; i < 6 ; i++)
        if(!ctx->w[i])
        {
            CUDA_TRY(ctx, cudaMalloc(&ctx->w[i], ctx->vec_len*sizeof(double))) ;
            CUDA_TRY(ctx, cudaMemsetAsync(ctx->w[i], 0, ctx->vec_len*sizeof(double), ctx->stream)) ;
        }
    return AMIE_B200_OK ;
}

int ctx_ensure_dinv(amie_b200_ctx * ctx, int kind)
{
    if(ctx->dinv_valid && ctx->dinv_kind == kind) return AMIE_B200_OK ;
    if(kind == AMIE_B200_PRECOND_DIAGONAL)
    {
        if(!ctx->user_diag) { ctx->set_error("diagonal preconditioner: amie_b200_set_preconditioner_diagonal was not called") ; return AMIE_B200_ERR_STATE ; }
    }
    else if(!ctx->have_values) { ctx->set_error("inverse diagonal: no values") ; return AMIE_B200_ERR_STATE ; }
    const bool block = kind == AMIE_B200_PRECOND_BLOCK2X2 || kind == AMIE_B200_PRECOND_BLOCK3X3 ;
    if(block && ctx->S != (kind == AMIE_B200_PRECOND_BLOCK2X2 ? 2 : 3))
    {
        ctx->set_error("block preconditioner: kind 5 (Inverse2x2Diagonal) takes stride 2, kind 6 stride 3") ;
        return AMIE_B200_ERR_UNSUPPORTED ;
    }
    if(!block && kind != AMIE_B200_PRECOND_JACOBI && kind != AMIE_B200_PRECOND_DIAGONAL && ctx->dist)
    {
        ctx->set_error("InverseDiagonalSquared / InverseLumpedDiagonal: not available on a row-partitioned context (pass the diagonal)") ;
        return AMIE_B200_ERR_UNSUPPORTED ;
    }
    const uint64_t need = std::max<uint64_t>(block ? ctx->N*(uint64_t)ctx->S : ctx->N, 1) ;
    if(ctx->dinv_len < need)
    {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
        dfree(ctx->dinv) ;
        ctx->dinv_len = 0 ;
        CUDA_TRY(ctx, cudaMalloc(&ctx->dinv, need*sizeof(double))) ;
        ctx->dinv_len = need ;
    }
    ctx->dinv_valid = false ;
    int grid = vec_grid(ctx, ctx->N) ;
#define DIAG_BY_STRIDE(KERNEL, ...) do { \
        if(ctx->S == 3)      KERNEL<3><<<grid, 256, 0, ctx->stream>>>(__VA_ARGS__) ; \
        else if(ctx->S == 2) KERNEL<2><<<grid, 256, 0, ctx->stream>>>(__VA_ARGS__) ; \
        else if(ctx->S == 1) KERNEL<1><<<grid, 256, 0, ctx->stream>>>(__VA_ARGS__) ; \
        else if(ctx->S == 4) KERNEL<4><<<grid, 256, 0, ctx->stream>>>(__VA_ARGS__) ; \
        else                 KERNEL<6><<<grid, 256, 0, ctx->stream>>>(__VA_ARGS__) ; } while(0)
    if(kind == AMIE_B200_PRECOND_DIAGONAL)
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dinv, ctx->user_diag, ctx->N*sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream)) ;
    else if(kind == AMIE_B200_PRECOND_DIAGONAL_SQUARED)
        DIAG_BY_STRIDE(k_inverse_diagonal_squared, ctx->rowptr, ctx->col, ctx->vals, ctx->nb, ctx->dinv) ;
    else if(kind == AMIE_B200_PRECOND_LUMPED)
        DIAG_BY_STRIDE(k_inverse_lumped_diagonal, ctx->rowptr, ctx->col, ctx->vals, ctx->nb, ctx->dinv) ;
    else if(block)
    {
        const int g = vec_grid(ctx, ctx->nb) ;
        if(ctx->S == 2) k_block_inverse<2><<<g, 256, 0, ctx->stream>>>(ctx->rowptr, ctx->col, ctx->vals, ctx->nb, ctx->dinv) ;
        else            k_block_inverse<3><<<g, 256, 0, ctx->stream>>>(ctx->rowptr, ctx->col, ctx->vals, ctx->nb, ctx->dinv) ;
    }
    else if(kind != AMIE_B200_PRECOND_JACOBI) { ctx->set_error("unknown diagonal preconditioner kind") ; return AMIE_B200_ERR_ARG ; }
    else if(ctx->dist)
    {
        int rc = dist_inverse_diagonal(ctx) ;       // renumbered columns are not sorted: linear scan for the diagonal block
        if(rc) return rc ;
    }
    else DIAG_BY_STRIDE(k_inverse_diagonal, ctx->rowptr, ctx->col, ctx->vals, 0u, ctx->nb, ctx->dinv) ;
#undef DIAG_BY_STRIDE
    CUDA_TRY(ctx, cudaGetLastError()) ;
    ctx->stats.kernel_launches++ ;
    ctx->dinv_valid = true ;
    ctx->dinv_kind = kind ;
    return AMIE_B200_OK ;
}

int ctx_sync_state(amie_b200_ctx * ctx, int slot)
{
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->st_host+slot, ctx->st, sizeof(KrylovState), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return AMIE_B200_OK ;
}

int ctx_push_state(amie_b200_ctx * ctx, const KrylovState & s)
{
    ctx->st_host[3] = s ;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->st, ctx->st_host+3, sizeof(KrylovState), cudaMemcpyHostToDevice, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;       // st_host[3] is reused
    return AMIE_B200_OK ;
}

void ctx_reset_solve_stats(amie_b200_ctx * ctx)
{
    ctx->stats.spmv_launches = ctx->stats.kernel_launches = ctx->stats.smoothing_spmv = 0 ;
    ctx->stats.iterations = ctx->stats.restarts = 0 ;
    ctx->stats.spmv_ms_total = 0. ; ctx->stats.spmv_timed = 0 ;
    ctx->stats.solve_ms = 0. ;
    ctx->stats.early_return = 0 ;
    ctx->ev_used = 0 ;
}

void ctx_collect_spmv_times(amie_b200_ctx * ctx)
{
    for(size_t i = 0 ; i+1 < ctx->ev_used ; i += 2)
    {
        float ms = 0.f ;
        if(cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i+1]) == cudaSuccess)
        {
            ctx->stats.spmv_ms_total += ms ;
            ctx->stats.spmv_timed++ ;
        }
    }
    ctx->ev_used = 0 ;
}

int ctx_max(amie_b200_ctx * ctx, const double * v, uint64_t n, int mode, double * out)
{
    int grid = vec_grid(ctx, n) ;
    k_max<<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(v, n, mode, ctx->partials+AMIE_MAX_PARTIALS*2) ;
    ctx->stats.kernel_launches++ ;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->partials_host, ctx->partials+AMIE_MAX_PARTIALS*2, grid*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    double m = ctx->partials_host[0] ;
    for(int i = 1 ; i < grid ; i++)
    {
        double y = ctx->partials_host[i] ;
        if(y > m || y != y) m = y ;
    }
    *out = m ;
    if(ctx->dist) return dist_allreduce_max(ctx, out) ;
    return AMIE_B200_OK ;
}

extern "C" {

const char * amie_b200_version(void) { return "amie_b200 0.1 (sm_100a)" ; }
const char * amie_b200_global_error(void) { return g_error.c_str() ; }
const char * amie_b200_last_error(const amie_b200_ctx * ctx) { return ctx ? ctx->err.c_str() : g_error.c_str() ; }

amie_b200_ctx * amie_b200_create(const int * devices, int ndev)
{
    // env AMIE_B200_DEVICES="0,1,2,3": what a host that cannot pass arguments (the drop-in shim) uses
    std::vector<int> env_devs ;
    if((!devices || ndev < 1))
        if(const char * e = getenv("AMIE_B200_DEVICES"))
        {
            for(const char * q = e ; *q ; )
            {
                char * end = nullptr ;
                long v = strtol(q, &end, 10) ;
                if(end == q) break ;
                env_devs.push_back((int)v) ;
                q = (*end == ',') ? end+1 : end ;
                if(*end && *end != ',') break ;
            }
            if(!env_devs.empty()) { devices = env_devs.data() ; ndev = (int)env_devs.size() ; }
        }
    if(ndev > 1) return group_create(devices, ndev, g_error) ;
    int dev = 0 ;
    if(devices && ndev == 1) dev = devices[0] ;
    else if(const char * e = getenv("AMIE_B200_DEVICE")) dev = atoi(e) ;
    int count = 0 ;
    G_TRY(cudaGetDeviceCount(&count)) ;
    if(dev < 0 || dev >= count) { g_error = "amie_b200_create: no such CUDA device" ; return nullptr ; }
    G_TRY(cudaSetDevice(dev)) ;
    cudaDeviceProp prop ;
    G_TRY(cudaGetDeviceProperties(&prop, dev)) ;
    if(prop.major != 10)
    {
        g_error = std::string("amie_b200_create: device ")+prop.name+" is not sm_100 (this library is built for sm_100a only)" ;
        return nullptr ;
    }
    amie_b200_ctx * ctx = new amie_b200_ctx ;
    ctx->device = dev ;
    ctx->num_sms = prop.multiProcessorCount ;
    // from here on a failure releases what was created so far (amie_b200_destroy tolerates null members)
#undef G_TRY
#define G_TRY(expr) do { cudaError_t _e = (expr) ; if(_e != cudaSuccess) { \
        g_error = std::string(#expr) + ": " + cudaGetErrorString(_e) ; amie_b200_destroy(ctx) ; cudaGetLastError() ; return nullptr ; } } while(0)
    G_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) ;
    G_TRY(cudaMalloc(&ctx->st, sizeof(KrylovState))) ;
    G_TRY(cudaMemset(ctx->st, 0, sizeof(KrylovState))) ;
    G_TRY(cudaMallocHost(&ctx->st_host, 4*sizeof(KrylovState))) ;
    G_TRY(cudaMalloc(&ctx->partials, 4*AMIE_MAX_PARTIALS*sizeof(double))) ;
    G_TRY(cudaMallocHost(&ctx->partials_host, AMIE_MAX_PARTIALS*sizeof(double))) ;
    G_TRY(cudaMalloc(&ctx->flag, sizeof(int))) ;
    G_TRY(cudaEventCreate(&ctx->ev_a)) ;
    G_TRY(cudaEventCreate(&ctx->ev_b)) ;
    G_TRY(cudaEventCreateWithFlags(&ctx->ev_poll[0], cudaEventDisableTiming)) ;
    G_TRY(cudaEventCreateWithFlags(&ctx->ev_poll[1], cudaEventDisableTiming)) ;
    return ctx ;
}

void amie_b200_destroy(amie_b200_ctx * ctx)
{
    if(!ctx) return ;
    if(ctx->group) { group_destroy(ctx) ; return ; }
    cudaSetDevice(ctx->device) ;
    if(ctx->stream) cudaStreamSynchronize(ctx->stream) ;
    dist_destroy(ctx) ;
    if(ctx->graph_cg.exec) cudaGraphExecDestroy(ctx->graph_cg.exec) ;
    if(ctx->graph_bicg.exec) cudaGraphExecDestroy(ctx->graph_bicg.exec) ;
    ctx_free_matrix(ctx) ;
    free_vectors(ctx) ;
    history_destroy(ctx) ;
    dfree(ctx->st) ; dfree(ctx->partials) ; dfree(ctx->flag) ;
    if(ctx->st_host) cudaFreeHost(ctx->st_host) ;
    if(ctx->partials_host) cudaFreeHost(ctx->partials_host) ;
    for(auto e : ctx->ev_pool) cudaEventDestroy(e) ;
    if(ctx->ev_a) cudaEventDestroy(ctx->ev_a) ;
    if(ctx->ev_b) cudaEventDestroy(ctx->ev_b) ;
    for(int i = 0 ; i < 2 ; i++) if(ctx->ev_poll[i]) cudaEventDestroy(ctx->ev_poll[i]) ;
    if(ctx->stream) cudaStreamDestroy(ctx->stream) ;
    delete ctx ;
}

int amie_b200_set_option(amie_b200_ctx * ctx, const char * key, int64_t value)
{
    if(!ctx || !key) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_set_option(ctx, key, value) ;
    std::string k(key) ;
    if(k == "time_spmv")
    {
        ctx->opt_time_spmv = (int)value ;
        if(value && ctx->ev_pool.empty())
        {
            cudaSetDevice(ctx->device) ;
            ctx->ev_pool.resize(16384) ;
            for(auto & e : ctx->ev_pool) CUDA_TRY(ctx, cudaEventCreate(&e)) ;
        }
    }
    else if(k == "spmv_variant") ctx->opt_variant = (int)value ;
    else if(k == "verbose") ctx->opt_verbose = (int)value ;
    else if(k == "iters_per_batch") ctx->opt_batch = (int)value ;
    else if(k == "graph") ctx->opt_graph = (int)value ;
    else if(k == "fields_variant") ctx->opt_fields_variant = (int)value ;
    else { ctx->set_error("unknown option "+k) ; return AMIE_B200_ERR_ARG ; }
    return AMIE_B200_OK ;
}

int amie_b200_get_stats(const amie_b200_ctx * ctx, amie_b200_stats * out)
{
    if(!ctx || !out) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_get_stats(ctx, out) ;
    *out = ctx->stats ;
    out->stride = ctx->S ; out->nb = ctx->nb ; out->nnzb = ctx->nnzb ; out->ndof = ctx->N ;
    out->spmv_algorithmic_bytes = ctx->nnzb*(uint64_t)(8*ctx->S*ctx->S+4)+4*(ctx->nb+1)+16*ctx->N ;
    out->device_bytes = device_bytes(ctx) ;
    return AMIE_B200_OK ;
}

int amie_b200_set_structure(amie_b200_ctx * ctx, int stride, uint64_t nb, const uint32_t * row_size,
                            const uint32_t * column_index, uint64_t nnzb)
{
    if(ctx && ctx->group) return group_set_structure(ctx, stride, nb, row_size, column_index, nnzb) ;
    if(ctx && ctx->dist) { ctx->set_error("set_structure on a distributed context: use amie_b200_dist_set_structure") ; return AMIE_B200_ERR_STATE ; }
    return ctx_set_structure(ctx, stride, nb, row_size, column_index, nnzb, nb) ;
}

}

int ctx_set_structure(amie_b200_ctx * ctx, int stride, uint64_t nb, const uint32_t * row_size,
                      const uint32_t * column_index, uint64_t nnzb, uint64_t ncols)
{
    if(!ctx || !row_size || (!column_index && nnzb)) return AMIE_B200_ERR_ARG ;
    if(stride != 1 && stride != 2 && stride != 3 && stride != 4 && stride != 6)
    {
        ctx->set_error("set_structure: strides 1, 2, 3, 4 and 6 are on the device path (2 and 3 on the tuned kernels)") ;
        return AMIE_B200_ERR_UNSUPPORTED ;
    }
    if(ctx->dist && stride != 2 && stride != 3)
    {
        ctx->set_error("set_structure: a row-partitioned context takes stride 2 or 3") ;
        return AMIE_B200_ERR_UNSUPPORTED ;
    }
    if(nnzb >= 0xffffffffull || nb >= 0xffffffffull) { ctx->set_error("set_structure: more than 2^32-1 blocks") ; return AMIE_B200_ERR_UNSUPPORTED ; }
    double t0 = wall_now() ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    ctx_free_matrix(ctx) ;
    // accumulated_row_size (sparse/sparse_matrix.cpp:59-66), with the total appended
    std::vector<uint32_t> rp(nb+1) ;
    uint64_t acc = 0 ;
    for(uint64_t i = 0 ; i < nb ; i++) { rp[i] = (uint32_t)acc ; acc += row_size[i] ; }
    rp[nb] = (uint32_t)acc ;
    if(acc != nnzb) { ctx->set_error("set_structure: sum(row_size) != nnzb") ; return AMIE_B200_ERR_ARG ; }
    ctx->S = stride ; ctx->nb = ctx->nb_global = nb ; ctx->row_base = 0 ; ctx->nnzb = nnzb ;
    ctx->N = nb*stride ; ctx->ncols_local = nb ;
    CUDA_TRY(ctx, cudaMalloc(&ctx->rowptr, (nb+1)*sizeof(uint32_t))) ;
    // +16 B: the TMA bulk copies round their end up to 16 bytes
    CUDA_TRY(ctx, cudaMalloc(&ctx->col, std::max<uint64_t>(nnzb, 1)*sizeof(uint32_t)+16)) ;
    CUDA_TRY(ctx, cudaMalloc(&ctx->vals, std::max<uint64_t>(nnzb, 1)*stride*stride*sizeof(double)+16)) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->rowptr, rp.data(), (nb+1)*sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream)) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->col, column_index, nnzb*sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream)) ;
    // validate: indices in range and strictly ascending per row (binary searches depend on it)
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->flag, 0, sizeof(int), ctx->stream)) ;
    k_rowptr_from_sizes_check<<<vec_grid(ctx, nb), 256, 0, ctx->stream>>>(ctx->col, ctx->rowptr, nb, (uint32_t)ncols, ctx->flag) ;
    int bad = 0 ;
    CUDA_TRY(ctx, cudaMemcpyAsync(&bad, ctx->flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    if(bad)
    {
        ctx_free_matrix(ctx) ;
        ctx->set_error(bad == 1 ? "set_structure: column index out of range" : "set_structure: column indices not strictly ascending in a row") ;
        return AMIE_B200_ERR_ARG ;
    }
    ctx->have_structure = true ;
    int rc = ctx_alloc_vectors(ctx) ;
    if(rc) return rc ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    ctx->stats.structure_ms = (wall_now()-t0)*1e3 ;
    return AMIE_B200_OK ;
}

extern "C" {

int amie_b200_set_values(amie_b200_ctx * ctx, const double * array)
{
    if(!ctx || (!array && ctx->nnzb)) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_set_values(ctx, array) ;
    if(!ctx->have_structure) { ctx->set_error("set_values before set_structure") ; return AMIE_B200_ERR_STATE ; }
    double t0 = wall_now() ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    const int S = ctx->S ;
    const int cl = S+S%2 ;
    if(cl == S && !ctx->block_to)
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->vals, array, ctx->nnzb*S*S*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    else
    {
        // K-Repack: stream the padded array through a device staging buffer, 12 -> 9 doubles per block
        // (with a block map: every stride goes this way and the blocks are scattered to their stored positions)
        const uint64_t chunk_blocks = std::min<uint64_t>(std::max<uint64_t>(ctx->nnzb, 1), (256ull << 20)/(S*cl*8)) ;
        double * stage[2] = {nullptr, nullptr} ;
        for(int i = 0 ; i < 2 ; i++)
        {
            cudaError_t e = cudaMalloc(&stage[i], chunk_blocks*S*cl*sizeof(double)) ;
            if(e != cudaSuccess)
            {
                if(stage[0]) cudaFree(stage[0]) ;
                ctx->set_error(std::string("set_values: staging buffer: ")+cudaGetErrorString(e)) ;
                return AMIE_B200_ERR_CUDA ;
            }
        }
        cudaEvent_t done[2] ;
        for(int i = 0 ; i < 2 ; i++) cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) ;
        int which = 0 ;
        for(uint64_t k = 0 ; k < ctx->nnzb ; k += chunk_blocks, which ^= 1)
        {
            uint64_t nblk = std::min(chunk_blocks, ctx->nnzb-k) ;
            cudaEventSynchronize(done[which]) ;
            cudaMemcpyAsync(stage[which], array+k*S*cl, nblk*S*cl*sizeof(double), cudaMemcpyHostToDevice, ctx->stream) ;
            if(ctx->block_to)
            {
                const int g = vec_grid(ctx, nblk*S*S) ;
                const uint32_t * map = ctx->block_to+k ;
                if(S == 3)      k_repack_scatter<3><<<g, 256, 0, ctx->stream>>>(stage[which], ctx->vals, map, nblk) ;
                else if(S == 2) k_repack_scatter<2><<<g, 256, 0, ctx->stream>>>(stage[which], ctx->vals, map, nblk) ;
                else if(S == 1) k_repack_scatter<1><<<g, 256, 0, ctx->stream>>>(stage[which], ctx->vals, map, nblk) ;
                else if(S == 4) k_repack_scatter<4><<<g, 256, 0, ctx->stream>>>(stage[which], ctx->vals, map, nblk) ;
                else            k_repack_scatter<6><<<g, 256, 0, ctx->stream>>>(stage[which], ctx->vals, map, nblk) ;
            }
            else if(S == 3) k_repack<3><<<vec_grid(ctx, nblk*9), 256, 0, ctx->stream>>>(stage[which], ctx->vals+k*9, nblk) ;
            else            k_repack<1><<<vec_grid(ctx, nblk), 256, 0, ctx->stream>>>(stage[which], ctx->vals+k, nblk) ;
            cudaEventRecord(done[which], ctx->stream) ;
        }
        cudaError_t e = cudaStreamSynchronize(ctx->stream) ;
        for(int i = 0 ; i < 2 ; i++) { cudaFree(stage[i]) ; cudaEventDestroy(done[i]) ; }
        CUDA_TRY(ctx, e) ;
        CUDA_TRY(ctx, cudaGetLastError()) ;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    ctx->have_values = true ;
    ctx->dinv_valid = false ;
    ctx->stats.values_ms = (wall_now()-t0)*1e3 ;
    return AMIE_B200_OK ;
}

int amie_b200_set_block_map(amie_b200_ctx * ctx, const uint32_t * block_to)
{
    if(!ctx) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_set_block_map(ctx, block_to) ;
    if(!ctx->have_structure) { ctx->set_error("set_block_map before set_structure") ; return AMIE_B200_ERR_STATE ; }
    if(ctx->dist) { ctx->set_error("set_block_map: not available on a row-partitioned context") ; return AMIE_B200_ERR_UNSUPPORTED ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    if(!block_to) { dfree(ctx->block_to) ; return AMIE_B200_OK ; }
    std::vector<bool> seen(ctx->nnzb, false) ;
    for(uint64_t k = 0 ; k < ctx->nnzb ; k++)
    {
        if(block_to[k] >= ctx->nnzb || seen[block_to[k]]) { ctx->set_error("set_block_map: not a permutation of the stored blocks") ; return AMIE_B200_ERR_ARG ; }
        seen[block_to[k]] = true ;
    }
    if(!ctx->block_to) CUDA_TRY(ctx, cudaMalloc(&ctx->block_to, std::max<uint64_t>(ctx->nnzb, 1)*sizeof(uint32_t))) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->block_to, block_to, ctx->nnzb*sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    ctx->have_values = false ;            // whatever was uploaded before was in the other order
    ctx->dinv_valid = false ;
    return AMIE_B200_OK ;
}

int amie_b200_upload_rhs(amie_b200_ctx * ctx, const double * b)
{
    if(!ctx || !b) return AMIE_B200_ERR_ARG ;
    if(ctx->group)
    {
        int rc = group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t) { return amie_b200_upload_rhs(c, b+d0) ; }) ;
        if(!rc) ctx->have_rhs = true ;
        return rc ;
    }
    if(!ctx->have_structure) { ctx->set_error("upload_rhs before set_structure") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->b, b, ctx->N*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    ctx->have_rhs = true ;
    return AMIE_B200_OK ;
}

int amie_b200_upload_x0(amie_b200_ctx * ctx, const double * x0, uint64_t nx0)
{
    if(!ctx || (!x0 && nx0)) return AMIE_B200_ERR_ARG ;
    if(ctx->group)
        return group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t nd)
        {
            const uint64_t n0 = nx0 > d0 ? std::min<uint64_t>(nx0-d0, nd) : 0 ;
            return amie_b200_upload_x0(c, n0 ? x0+d0 : nullptr, n0) ;
        }) ;
    if(!ctx->have_structure) { ctx->set_error("upload_x0 before set_structure") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->x, 0, ctx->vec_len*sizeof(double), ctx->stream)) ;
    uint64_t n = std::min(nx0, ctx->N) ;
    if(n) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->x, x0, n*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return AMIE_B200_OK ;
}

int amie_b200_download_x(amie_b200_ctx * ctx, double * x_out)
{
    if(!ctx || !x_out) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t) { return amie_b200_download_x(c, x_out+d0) ; }) ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(x_out, ctx->x, ctx->N*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return AMIE_B200_OK ;
}

int amie_b200_download_rhs(amie_b200_ctx * ctx, double * b_out)
{
    if(!ctx || !b_out) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t) { return amie_b200_download_rhs(c, b_out+d0) ; }) ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(b_out, ctx->b, ctx->N*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return AMIE_B200_OK ;
}

int amie_b200_download_vector(amie_b200_ctx * ctx, int which, double * out)
{
    if(!ctx || !out || which < 0 || which > 3) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t) { return amie_b200_download_vector(c, which, out+d0) ; }) ;
    const double * v[] = { ctx->x, ctx->q, ctx->r, ctx->p } ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(out, v[which], ctx->N*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return AMIE_B200_OK ;
}

int amie_b200_download_matrix(amie_b200_ctx * ctx, uint32_t * row_size_out, uint32_t * column_index_out, double * array_padded_out)
{
    if(!ctx) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_download_matrix(ctx, row_size_out, column_index_out, array_padded_out) ;
    if(!ctx->have_structure) { ctx->set_error("download_matrix: no matrix") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    if(row_size_out)
    {
        std::vector<uint32_t> rp(ctx->nb+1) ;
        CUDA_TRY(ctx, cudaMemcpy(rp.data(), ctx->rowptr, (ctx->nb+1)*sizeof(uint32_t), cudaMemcpyDeviceToHost)) ;
        for(uint64_t i = 0 ; i < ctx->nb ; i++) row_size_out[i] = rp[i+1]-rp[i] ;
    }
    if(column_index_out)
        CUDA_TRY(ctx, cudaMemcpy(column_index_out, ctx->col, ctx->nnzb*sizeof(uint32_t), cudaMemcpyDeviceToHost)) ;
    if(array_padded_out)
    {
        const int S = ctx->S, cl = S+S%2 ;
        std::vector<double> v(ctx->nnzb*S*S) ;
        CUDA_TRY(ctx, cudaMemcpy(v.data(), ctx->vals, v.size()*sizeof(double), cudaMemcpyDeviceToHost)) ;
        for(uint64_t k = 0 ; k < ctx->nnzb ; k++)
            for(int c = 0 ; c < S ; c++)
                for(int r = 0 ; r < cl ; r++)
                    array_padded_out[k*S*cl+c*cl+r] = r < S ? v[k*S*S+c*S+r] : 0. ;
    }
    return AMIE_B200_OK ;
}

int amie_b200_pcg_resident(amie_b200_ctx * ctx, int precond_kind, double eps, int maxit, uint64_t nssor,
                           uint64_t rowstart, uint64_t colstart, uint64_t * nit_out, double * err_out, double * rho_out)
{
    if(!ctx) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_solve(ctx, false, true, nullptr, nullptr, 0, precond_kind, eps, maxit, nssor, rowstart, colstart, nullptr, nit_out, err_out, rho_out) ;
    if(!ctx->have_values || !ctx->have_rhs) { ctx->set_error("pcg: matrix values / rhs not on the device") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    return solve_cg_resident(ctx, precond_kind, eps, maxit, nssor, rowstart, colstart, nit_out, err_out, rho_out) ;
}

int amie_b200_bicgstab_resident(amie_b200_ctx * ctx, int precond_kind, double eps, int maxit, uint64_t * nit_out, double * err_out)
{
    if(!ctx) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_solve(ctx, true, true, nullptr, nullptr, 0, precond_kind, eps, maxit, 0, 0, 0, nullptr, nit_out, err_out, nullptr) ;
    if(!ctx->have_values || !ctx->have_rhs) { ctx->set_error("bicgstab: matrix values / rhs not on the device") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    return solve_bicg_resident(ctx, precond_kind, eps, maxit, nit_out, err_out) ;
}

// host-buffer solver calls: H2D of b and x0, the resident solve, D2H of x -- what the drop-in pays per call
static int host_call_prologue(amie_b200_ctx * ctx, const double * b, const double * x0, uint64_t nx0, bool bicg)
{
    if(!ctx->have_values) { ctx->set_error("solve before set_values") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    cudaEventRecord(ctx->ev_a, ctx->stream) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->b, b, ctx->N*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->x, 0, ctx->vec_len*sizeof(double), ctx->stream)) ;
    // ConjugateGradient copies min(N, x0.size()) entries (conjugategradient.cpp:95-104);
    // BiCGStab only takes x0 when the sizes agree (biconjugategradientstabilized.cpp:21-24)
    uint64_t n = std::min(nx0, ctx->N) ;
    if(bicg && nx0 != ctx->N) n = 0 ;
    if(n) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->x, x0, n*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    cudaEventRecord(ctx->ev_b, ctx->stream) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    float ms = 0.f ;
    cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b) ;
    ctx->stats.h2d_ms = ms ;
    ctx->stats.h2d_bytes = (ctx->N+n)*sizeof(double) ;
    ctx->have_rhs = true ;
    return AMIE_B200_OK ;
}

static int host_call_epilogue(amie_b200_ctx * ctx, double * x_out)
{
    cudaEventRecord(ctx->ev_a, ctx->stream) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(x_out, ctx->x, ctx->N*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    cudaEventRecord(ctx->ev_b, ctx->stream) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    float ms = 0.f ;
    cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b) ;
    ctx->stats.d2h_ms = ms ;
    ctx->stats.d2h_bytes = ctx->N*sizeof(double) ;
    return AMIE_B200_OK ;
}

int amie_b200_pcg(amie_b200_ctx * ctx, const double * b, const double * x0, uint64_t nx0,
                  int precond_kind, double eps, int maxit, uint64_t nssor, uint64_t rowstart, uint64_t colstart,
                  double * x_out, uint64_t * nit_out, double * err_out, double * rho_out)
{
    if(!ctx || !b || !x_out || (!x0 && nx0)) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_solve(ctx, false, false, b, x0, nx0, precond_kind, eps, maxit, nssor, rowstart, colstart, x_out, nit_out, err_out, rho_out) ;
    int rc = host_call_prologue(ctx, b, x0, nx0, false) ;
    if(rc) return rc ;
    int ret = solve_cg_resident(ctx, precond_kind, eps, maxit, nssor, rowstart, colstart, nit_out, err_out, rho_out) ;
    if(ret < 0) return ret ;
    rc = host_call_epilogue(ctx, x_out) ;
    return rc ? rc : ret ;
}

int amie_b200_bicgstab(amie_b200_ctx * ctx, const double * b, const double * x0, uint64_t nx0,
                       int precond_kind, double eps, int maxit, double * x_out, uint64_t * nit_out, double * err_out)
{
    if(!ctx || !b || !x_out || (!x0 && nx0)) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_solve(ctx, true, false, b, x0, nx0, precond_kind, eps, maxit, 0, 0, 0, x_out, nit_out, err_out, nullptr) ;
    int rc = host_call_prologue(ctx, b, x0, nx0, true) ;
    if(rc) return rc ;
    int ret = solve_bicg_resident(ctx, precond_kind, eps, maxit, nit_out, err_out) ;
    if(ret < 0) return ret ;
    rc = host_call_epilogue(ctx, x_out) ;
    return rc ? rc : ret ;
}

int amie_b200_spmv(amie_b200_ctx * ctx, const double * x, const double * b, uint64_t rowstart, uint64_t colstart, double * y_out)
{
    if(!ctx || !x || !y_out) return AMIE_B200_ERR_ARG ;
    if(ctx->group)
        return group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t) { return amie_b200_spmv(c, x+d0, b ? b+d0 : nullptr, rowstart, colstart, y_out+d0) ; }) ;
    if(!ctx->have_values) { ctx->set_error("spmv before set_values") ; return AMIE_B200_ERR_STATE ; }
    if(rowstart%ctx->S || colstart%ctx->S || rowstart > ctx->nb_global*(uint64_t)ctx->S) { ctx->set_error("spmv: rowstart/colstart must be multiples of the stride") ; return AMIE_B200_ERR_ARG ; }
    if(ctx->dist) rowstart = dist_local_rowstart(ctx, rowstart) ;      // rows local, columns global (dist_spmv)
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->p, x, ctx->N*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    if(b) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->z, b, ctx->N*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->q, 0, ctx->N*sizeof(double), ctx->stream)) ;    // rows < rowstart := 0
    SpmvCall c ;
    c.x = ctx->p ; c.b = b ? ctx->z : nullptr ; c.y = ctx->q ; c.minus_b = b != nullptr ;
    c.rowstart = rowstart ; c.colstart = colstart ;
    int rc = launch_spmv(ctx, c) ;
    if(rc) return rc ;
    CUDA_TRY(ctx, cudaGetLastError()) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(y_out, ctx->q, ctx->N*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return AMIE_B200_OK ;
}

int amie_b200_residual(amie_b200_ctx * ctx, const double * u, const double * f, double * r_out, double * norm_out)
{
    if(!ctx || !u || !f) return AMIE_B200_ERR_ARG ;
    if(ctx->group)
    {
        double norms[GROUP_MAX] = {} ;
        int rc = group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t)
        {
            return amie_b200_residual(c, u+d0, f+d0, r_out ? r_out+d0 : nullptr, norms+dist_rank(c)) ;
        }) ;
        if(!rc && norm_out) *norm_out = norms[0] ;
        return rc ;
    }
    if(!ctx->have_values) { ctx->set_error("residual before set_values") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->p, u, ctx->N*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->z, f, ctx->N*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    SpmvCall c ;
    c.x = ctx->p ; c.b = ctx->z ; c.y = ctx->q ; c.minus_b = true ; c.dot = DOT_YY ; c.finalize = FIN_STORE ;
    int rc = launch_spmv(ctx, c) ;
    if(rc) return rc ;
    if((rc = ctx_sync_state(ctx, 2))) return rc ;
    CUDA_TRY(ctx, cudaGetLastError()) ;
    if(norm_out) *norm_out = sqrt(ctx->st_host[2].dot[0]) ;
    if(r_out)
    {
        CUDA_TRY(ctx, cudaMemcpyAsync(r_out, ctx->q, ctx->N*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    }
    return AMIE_B200_OK ;
}

int amie_b200_spmv_resident(amie_b200_ctx * ctx, int reps, int variant, double * ms_out)
{
    if(!ctx || reps < 1) return AMIE_B200_ERR_ARG ;
    if(ctx->group)
    {
        double ms[GROUP_MAX] = {} ;
        int rc = group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t, uint64_t) { return amie_b200_spmv_resident(c, reps, variant, ms+dist_rank(c)) ; }) ;
        if(!rc && ms_out) *ms_out = *std::max_element(ms, ms+GROUP_MAX) ;
        return rc ;
    }
    if(!ctx->have_values) { ctx->set_error("spmv before set_values") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    int saved = ctx->opt_variant, saved_t = ctx->opt_time_spmv ;
    ctx->opt_variant = variant ;
    ctx->opt_time_spmv = 0 ;
    SpmvCall c ;
    c.x = ctx->x ; c.y = ctx->q ;
    bool insolve = false ;
    if(variant >= 100)
    {
        // the in-solve form: q = A p fused with p.q (result stored, no loop control)
        variant -= 100 ;
        ctx->opt_variant = variant ;
        c.dot = DOT_YX ; c.finalize = FIN_STORE ;
        insolve = true ;
    }
    auto one = [&]() { launch_spmv_variant(ctx, c, variant, insolve) ; } ;
    one() ;                                                 // warm-up
    cudaEventRecord(ctx->ev_a, ctx->stream) ;
    for(int i = 0 ; i < reps ; i++) one() ;
    cudaEventRecord(ctx->ev_b, ctx->stream) ;
    ctx->opt_variant = saved ; ctx->opt_time_spmv = saved_t ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    CUDA_TRY(ctx, cudaGetLastError()) ;
    float ms = 0.f ;
    cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b) ;
    if(ms_out) *ms_out = ms/reps ;
    return AMIE_B200_OK ;
}

int amie_b200_set_preconditioner_diagonal(amie_b200_ctx * ctx, const double * d)
{
    if(!ctx || !d) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t) { return amie_b200_set_preconditioner_diagonal(c, d+d0) ; }) ;
    if(!ctx->have_structure) { ctx->set_error("set_preconditioner_diagonal before set_structure") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    if(!ctx->user_diag) CUDA_TRY(ctx, cudaMalloc(&ctx->user_diag, std::max<uint64_t>(ctx->N, 1)*sizeof(double))) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->user_diag, d, ctx->N*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    if(ctx->dinv_kind == AMIE_B200_PRECOND_DIAGONAL) ctx->dinv_valid = false ;
    return AMIE_B200_OK ;
}

int amie_b200_preconditioner_blocks(amie_b200_ctx * ctx, int precond_kind, double * blocks_out)
{
    if(!ctx || !blocks_out || (precond_kind != AMIE_B200_PRECOND_BLOCK2X2 && precond_kind != AMIE_B200_PRECOND_BLOCK3X3)) return AMIE_B200_ERR_ARG ;
    if(ctx->group)
        return group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t) { return amie_b200_preconditioner_blocks(c, precond_kind, blocks_out+d0*(uint64_t)c->S) ; }) ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    int rc = ctx_ensure_dinv(ctx, precond_kind) ;
    if(rc) return rc ;
    CUDA_TRY(ctx, cudaMemcpyAsync(blocks_out, ctx->dinv, ctx->N*(uint64_t)ctx->S*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return AMIE_B200_OK ;
}

int amie_b200_preconditioner_diagonal(amie_b200_ctx * ctx, int precond_kind, double * d_out)
{
    if(!ctx || !d_out || precond_kind == AMIE_B200_PRECOND_NULL || precond_kind > AMIE_B200_PRECOND_DIAGONAL) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t) { return amie_b200_preconditioner_diagonal(c, precond_kind, d_out+d0) ; }) ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    int rc = ctx_ensure_dinv(ctx, precond_kind) ;
    if(rc) return rc ;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_out, ctx->dinv, ctx->N*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return AMIE_B200_OK ;
}

int amie_b200_inverse_diagonal(amie_b200_ctx * ctx, double * d_out)
{
    if(!ctx || !d_out) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t) { return amie_b200_inverse_diagonal(c, d_out+d0) ; }) ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    int rc = ctx_ensure_dinv(ctx) ;
    if(rc) return rc ;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_out, ctx->dinv, ctx->N*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return AMIE_B200_OK ;
}

}


// Values of this context's stored blocks j = 0 .. nnzb-1 taken from blocks src[j] of a host array in the reference's
// padded layout (a multi-device context under a block map, group.cu: the blocks of one device's rows are scattered over
// the caller's array).  Host gather into a staging chunk, then the usual K-Repack; one chunk of host memory.
void amie_b200_gather_blocks(const double * array, const uint32_t * src, uint64_t nblk, uint64_t per_block, double * out)
{
    #pragma omp parallel for schedule(static)
    for(long long j = 0 ; j < (long long)nblk ; j++)
        std::memcpy(out+(uint64_t)j*per_block, array+(uint64_t)src[j]*per_block, per_block*sizeof(double)) ;
}

int ctx_set_values_from(amie_b200_ctx * ctx, const double * array, const uint32_t * src)
{
    if(!ctx->have_structure) { ctx->set_error("set_values before set_structure") ; return AMIE_B200_ERR_STATE ; }
    const double t0 = wall_now() ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    const int S = ctx->S, cl = S+S%2 ;
    const uint64_t per_block = (uint64_t)S*cl ;
    const uint64_t chunk_blocks = std::min<uint64_t>(std::max<uint64_t>(ctx->nnzb, 1), (64ull << 20)/(per_block*8)) ;
    std::vector<double> host(chunk_blocks*per_block) ;
    double * stage = nullptr ;
    if(cl != S) CUDA_TRY(ctx, cudaMalloc(&stage, chunk_blocks*per_block*sizeof(double))) ;
    int rc = AMIE_B200_OK ;
    for(uint64_t k = 0 ; k < ctx->nnzb && !rc ; k += chunk_blocks)
    {
        const uint64_t nblk = std::min(chunk_blocks, ctx->nnzb-k) ;
        amie_b200_gather_blocks(array, src+k, nblk, per_block, host.data()) ;
        cudaError_t e ;
        if(cl == S)
            e = cudaMemcpyAsync(ctx->vals+k*per_block, host.data(), nblk*per_block*sizeof(double), cudaMemcpyHostToDevice, ctx->stream) ;
        else
        {
            e = cudaMemcpyAsync(stage, host.data(), nblk*per_block*sizeof(double), cudaMemcpyHostToDevice, ctx->stream) ;
            if(e == cudaSuccess)
            {
                if(S == 3) k_repack<3><<<vec_grid(ctx, nblk*9), 256, 0, ctx->stream>>>(stage, ctx->vals+k*9, nblk) ;
                else       k_repack<1><<<vec_grid(ctx, nblk), 256, 0, ctx->stream>>>(stage, ctx->vals+k, nblk) ;
                e = cudaGetLastError() ;
            }
        }
        // the host chunk is reused: the copy out of it must be over before the next gather
        if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream) ;
        if(e != cudaSuccess) { ctx->set_error(std::string("set_values (gathered): ")+cudaGetErrorString(e)) ; rc = AMIE_B200_ERR_CUDA ; }
    }
    if(stage) cudaFree(stage) ;
    if(rc) return rc ;
    ctx->have_values = true ;
    ctx->dinv_valid = false ;
    ctx->stats.values_ms = (wall_now()-t0)*1e3 ;
    return AMIE_B200_OK ;
}
