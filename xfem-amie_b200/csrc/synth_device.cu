// synth_device.cu -- the synthetic-mesh generator of synth_recipe.h compiled for the device:
// structure, values (compact layout) and right-hand side are produced directly in HBM, so a
// 50 M-DOF system never needs its 43 GB padded host array.  Built with -fmad=false so that the
// bits equal the host generator's (tests/test_synth.py checks that on the GPU).
#include "launch.cuh"
#include "synth.h"
#include "group.h"
#include <cub/cub.cuh>

namespace {

__global__ void k_synth_count(const SynthRecipe * R, uint64_t row0, uint64_t nrows, uint32_t * sizes)
{
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < nrows ; i += (uint64_t)gridDim.x*blockDim.x)
        sizes[i] = (uint32_t)synth_row_count(*R, row0+i) ;
}

__global__ void k_synth_fill(const SynthRecipe * R, uint64_t row0, uint64_t nrows, const uint32_t * rowptr,
                             uint32_t * col, double * vals, double * b)
{
    const int S = R->stride ;
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < nrows ; i += (uint64_t)gridDim.x*blockDim.x)
    {
        uint32_t cols[27] ;
        double blocks[27*9] ;
        double rhs[3] ;
        const int cnt = synth_row(*R, row0+i, cols, blocks, rhs) ;
        const uint32_t k0 = rowptr[i] ;
        for(int k = 0 ; k < cnt ; k++)
        {
            col[k0+k] = cols[k] ;
            double * dst = vals+(size_t)(k0+k)*S*S ;
            for(int c = 0 ; c < S ; c++)
                for(int r = 0 ; r < S ; r++)
                    dst[c*S+r] = blocks[k*9+c*3+r] ;
        }
        for(int m = 0 ; m < S ; m++) b[i*S+m] = rhs[m] ;
    }
}

}

// rows [row0,row1) of the recipe -> device arrays (global column indices).  Allocates rowptr/col/vals.
int synth_rows_to_device(amie_b200_ctx * ctx, const SynthRecipe & R, uint64_t row0, uint64_t row1,
                         uint32_t ** rowptr_out, uint32_t ** col_out, double ** vals_out, double * b_dev, uint64_t * nnzb_out)
{
    const uint64_t nrows = row1-row0 ;
    const int S = R.stride ;
    SynthRecipe * dR = nullptr ;
    uint32_t * sizes = nullptr, * rowptr = nullptr, * col = nullptr ;
    double * vals = nullptr ;
    void * tmp = nullptr ;
    size_t tmp_bytes = 0 ;
    CUDA_TRY(ctx, cudaMalloc(&dR, sizeof(SynthRecipe))) ;
    CUDA_TRY(ctx, cudaMemcpyAsync(dR, &R, sizeof(SynthRecipe), cudaMemcpyHostToDevice, ctx->stream)) ;
    CUDA_TRY(ctx, cudaMalloc(&sizes, (nrows+1)*sizeof(uint32_t))) ;
    CUDA_TRY(ctx, cudaMalloc(&rowptr, (nrows+1)*sizeof(uint32_t))) ;
    CUDA_TRY(ctx, cudaMemsetAsync(sizes, 0, (nrows+1)*sizeof(uint32_t), ctx->stream)) ;
    const int grid = vec_grid(ctx, nrows) ;
    k_synth_count<<<grid, 256, 0, ctx->stream>>>(dR, row0, nrows, sizes) ;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, sizes, rowptr, nrows+1, ctx->stream) ;
    CUDA_TRY(ctx, cudaMalloc(&tmp, tmp_bytes)) ;
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, sizes, rowptr, nrows+1, ctx->stream) ;
    uint32_t total = 0 ;
    CUDA_TRY(ctx, cudaMemcpyAsync(&total, rowptr+nrows, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    cudaFree(tmp) ; cudaFree(sizes) ;
    CUDA_TRY(ctx, cudaMalloc(&col, std::max<uint64_t>(total, 1)*sizeof(uint32_t)+16)) ;       // +16 B: see api.cu
    CUDA_TRY(ctx, cudaMalloc(&vals, std::max<uint64_t>(total, 1)*S*S*sizeof(double)+16)) ;
    k_synth_fill<<<grid, 128, 0, ctx->stream>>>(dR, row0, nrows, rowptr, col, vals, b_dev) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    CUDA_TRY(ctx, cudaGetLastError()) ;
    cudaFree(dR) ;
    *rowptr_out = rowptr ; *col_out = col ; *vals_out = vals ; *nnzb_out = total ;
    return AMIE_B200_OK ;
}

extern "C" int amie_b200_synth_to_device(amie_b200_ctx * ctx, const amie_b200_synth * s)
{
    if(!ctx || !s) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_synth_to_device(ctx, s) ;
    const SynthRecipe & R = *synth_recipe_of(s) ;
    double t0 = wall_now() ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    const uint64_t nb = synth_num_nodes(R) ;
    if(nb >= 0xffffffffull) { ctx->set_error("synth_to_device: too many nodes") ; return AMIE_B200_ERR_UNSUPPORTED ; }
    // a rough upper bound of the stored blocks must fit uint32 offsets
    if(nb*27 >= 0xffffffffull && R.dim == 3 && R.ntemplates == 1) { ctx->set_error("synth_to_device: more than 2^32-1 blocks") ; return AMIE_B200_ERR_UNSUPPORTED ; }
    ctx_free_matrix(ctx) ;
    ctx->S = R.stride ; ctx->nb = ctx->nb_global = nb ; ctx->row_base = 0 ; ctx->N = nb*R.stride ; ctx->ncols_local = nb ;
    int rc = ctx_alloc_vectors(ctx) ;
    if(rc) return rc ;
    uint64_t nnzb = 0 ;
    rc = synth_rows_to_device(ctx, R, 0, nb, &ctx->rowptr, &ctx->col, &ctx->vals, ctx->b, &nnzb) ;
    if(rc) return rc ;
    ctx->nnzb = nnzb ;
    ctx->have_structure = ctx->have_values = ctx->have_rhs = true ;
    ctx->stats.structure_ms = (wall_now()-t0)*1e3 ;
    ctx->stats.values_ms = 0. ;
    return AMIE_B200_OK ;
}
