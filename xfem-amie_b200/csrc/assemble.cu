// assemble.cu -- device-side value assembly and Dirichlet elimination.
//
// What it replaces: the stiffness scatter loops of Assembly::make_final (solvers/assembly.cpp:657-735 in 2D,
// :1060-1138 in 3D) and Assembly::setBoundaryConditions (:125-330), i.e. the step that produces the `array`
// the Krylov solve consumes.  Repeated re-solves on one topology (damage iterations: thousands of steps on the
// same mesh) then upload only the elementary matrices that changed instead of the whole padded array.
//
// Bit-exactness.  The reference adds the element blocks into each stored block in ELEMENT ORDER with a Kahan
// compensator per entry (:681-697).  A scatter with atomics cannot reproduce that, so the scatter is turned into a
// GATHER: once per topology every stored block gets the list of element blocks that land on it, ascending (one per
// element, so ascending source index == element order); one thread per stored entry then replays the reference's
// compensated sum.  Same additions, same order, no floating-point atomics -> the same bits as the CPU, and
// run-to-run deterministic.  The elimination runs one thread per scalar row, which owns every entry and the
// right-hand-side component the reference touches while it walks that row.
//
// Layout.  update_elements does not keep the elementary matrices in upload order: a placing kernel writes every
// block, multiplied by its element's scale, to ITS position in a stream where the contributions of one stored block
// are contiguous and in element order (the permutation is built once per topology by set_elements).  The gather then
// reads that stream front to back -- no index list, no scale lookup, whole sectors.
//
// Traffic: gather = element blocks once (8 s^2 B each) + 4 B of list offsets per stored block + the stored blocks
// written once; elimination = column indices + a per-node mask byte, values only where a fixed dof is involved.
#include "context.h"
#include "group.h"
#include "kernels_assemble.cuh"
#include <algorithm>
#include <vector>
#include <cub/device/device_scan.cuh>


struct AssemblyMap
{
    uint64_t n_elem = 0 ;
    int npe = 0 ;
    uint64_t nsrc = 0 ;                 // n_elem*npe*npe element blocks
    uint32_t * dest_of_src = nullptr ;  // [nsrc] stored block each element block lands on (NO_DEST: unused node slot)
    uint32_t * cptr = nullptr ;         // [nnzb+1] contribution runs per stored block (positions in `placed`)
    uint32_t * pos_of_src = nullptr ;   // [nsrc] position of every element block in `placed` (NO_DEST: unused node slot)
    uint64_t total = 0 ;                // contributions in all
    double * placed = nullptr ;         // [total*S*S] scale*Ke, the contributions of a stored block contiguous and in element order
    double * stage = nullptr ;          // upload staging (STAGE_DOUBLES doubles) + one scale per staged element
    double * stage_scales = nullptr ;
    unsigned char * dirty = nullptr ;   // [nnzb] stored blocks to re-accumulate at the next assemble
    bool built = false ;                // set_elements done (the struct also carries the BC scratch alone)
    bool all_dirty = true ;
    bool have_ke = false ;
    // boundary-condition scratch (sized on demand)
    unsigned char * fixmask = nullptr ; // [nb] bit n: dof n of the node is eliminated
    unsigned char * forcemask = nullptr ;
    uint64_t mask_nb = 0 ;
} ;

// doubles staged per upload chunk (64 MB): update_elements streams the host array through it
static const uint64_t STAGE_DOUBLES = 8ull << 20 ;

template<typename T> static void afree(T *& p) { if(p) cudaFree(p) ; p = nullptr ; }

void assembly_map_destroy(amie_b200_ctx * ctx)
{
    AssemblyMap * m = ctx->amap ;
    if(!m) return ;
    afree(m->dest_of_src) ; afree(m->cptr) ; afree(m->pos_of_src) ; afree(m->placed) ; afree(m->stage) ; afree(m->stage_scales) ; afree(m->dirty) ;
    afree(m->fixmask) ; afree(m->forcemask) ;
    delete m ;
    ctx->amap = nullptr ;
}

uint64_t assembly_map_bytes(const amie_b200_ctx * ctx)
{
    const AssemblyMap * m = ctx->amap ;
    if(!m) return 0 ;
    uint64_t b = 0 ;
    if(m->built)
    {
        const uint64_t SS = (uint64_t)ctx->S*ctx->S ;
        b += m->nsrc*4+(ctx->nnzb+1)*4+ctx->nnzb+m->total*SS*8 ;                // dest_of_src, cptr, dirty, placed
        b += m->nsrc*4 ;                                                         // pos_of_src
        if(m->stage) b += STAGE_DOUBLES*8+STAGE_DOUBLES/SS*8 ;
    }
    if(m->fixmask) b += 2*m->mask_nb ;
    return b ;
}

// ---------------------------------------------------------------------------------------------------- API

static int require_structure(amie_b200_ctx * ctx, const char * what)
{
    if(!ctx->have_structure) { ctx->set_error(std::string(what)+" before set_structure") ; return AMIE_B200_ERR_STATE ; }
    return AMIE_B200_OK ;
}

// the numbering of the matrix this context holds (kernels_assemble.cuh): the whole matrix, or one part of a partitioned one
static PartMap part_map(const amie_b200_ctx * ctx)
{
    PartMap pm ;
    pm.row_base = (uint32_t)ctx->row_base ;
    pm.nb = (uint32_t)ctx->nb ;
    pm.halo = ctx->halo_glob ;
    pm.nhalo = (uint32_t)(ctx->ncols_local > ctx->nb ? ctx->ncols_local-ctx->nb : 0) ;
    pm.nb_global = (uint32_t)(ctx->nb_global ? ctx->nb_global : ctx->nb) ;
    return pm ;
}

static bool ascending_unique(const uint32_t * ids, uint64_t n, uint64_t limit)
{
    for(uint64_t i = 0 ; i < n ; i++)
        if(ids[i] >= limit || (i && ids[i-1] >= ids[i])) return false ;
    return true ;
}

extern "C" {

int amie_b200_set_elements(amie_b200_ctx * ctx, uint64_t n_elem, int npe, const uint32_t * elem_ids)
{
    if(!ctx || npe < 1 || npe > 64 || (!elem_ids && n_elem)) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_set_elements(ctx, n_elem, npe, elem_ids) ;
    int rc = require_structure(ctx, "set_elements") ;
    if(rc) return rc ;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    assembly_map_destroy(ctx) ;
    const uint64_t nsrc = n_elem*(uint64_t)npe*npe ;
    if(nsrc >= 0xFFFFFFFFull) { ctx->set_error("set_elements: more than 2^32 element blocks") ; return AMIE_B200_ERR_UNSUPPORTED ; }
    const double t0 = wall_now() ;
    AssemblyMap * m = new AssemblyMap ;
    ctx->amap = m ;
    m->n_elem = n_elem ; m->npe = npe ; m->nsrc = nsrc ;
    const int SS = ctx->S*ctx->S ;
    const uint64_t nnzb = ctx->nnzb ;
    uint32_t * ids = nullptr, * count = nullptr, * csrc = nullptr ;
    void * tmp = nullptr ;
    size_t tmp_bytes = 0 ;
    int bad = 0 ;
    auto cleanup = [&]() { afree(ids) ; afree(count) ; afree(csrc) ; if(tmp) cudaFree(tmp) ; tmp = nullptr ; } ;
#define MAP_TRY(expr) do { cudaError_t _e = (expr) ; if(_e != cudaSuccess) { cleanup() ; assembly_map_destroy(ctx) ; \
        ctx->set_error(std::string(#expr)+": "+cudaGetErrorString(_e)) ; return AMIE_B200_ERR_CUDA ; } } while(0)
    MAP_TRY(cudaMalloc(&ids, std::max<uint64_t>(n_elem*npe, 1)*sizeof(uint32_t))) ;
    MAP_TRY(cudaMalloc(&count, (nnzb+1)*sizeof(uint32_t))) ;
    MAP_TRY(cudaMalloc(&m->dest_of_src, std::max<uint64_t>(nsrc, 1)*sizeof(uint32_t))) ;
    MAP_TRY(cudaMalloc(&m->cptr, (nnzb+1)*sizeof(uint32_t))) ;
    MAP_TRY(cudaMalloc(&m->dirty, std::max<uint64_t>(nnzb, 1))) ;
    MAP_TRY(cudaMalloc(&m->pos_of_src, std::max<uint64_t>(nsrc, 1)*sizeof(uint32_t))) ;
    MAP_TRY(cudaMemcpyAsync(ids, elem_ids, n_elem*npe*sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream)) ;
    MAP_TRY(cudaMemsetAsync(count, 0, (nnzb+1)*sizeof(uint32_t), ctx->stream)) ;
    MAP_TRY(cudaMemsetAsync(ctx->flag, 0, sizeof(int), ctx->stream)) ;
    if(nsrc)
        k_map_dest<<<vec_grid(ctx, nsrc), AMIE_VEC_THREADS, 0, ctx->stream>>>(ctx->rowptr, ctx->col, part_map(ctx), ids, nsrc, npe,
                                                                             m->dest_of_src, count, ctx->flag) ;
    MAP_TRY(cudaMemcpyAsync(&bad, ctx->flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)) ;
    MAP_TRY(cudaStreamSynchronize(ctx->stream)) ;
    if(bad)
    {
        cleanup() ; assembly_map_destroy(ctx) ;
        ctx->set_error(bad == 1 ? "set_elements: node id out of range" : "set_elements: an element couples two nodes whose block is not in the sparsity pattern") ;
        return AMIE_B200_ERR_ARG ;
    }
    // exclusive scan of the counts -> list offsets (count[nnzb] == 0, so cptr[nnzb] is the total)
    MAP_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, count, m->cptr, nnzb+1, ctx->stream)) ;
    MAP_TRY(cudaMalloc(&tmp, std::max<size_t>(tmp_bytes, 16))) ;
    MAP_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, count, m->cptr, nnzb+1, ctx->stream)) ;
    uint32_t total = 0 ;
    MAP_TRY(cudaMemcpyAsync(&total, m->cptr+nnzb, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream)) ;
    MAP_TRY(cudaStreamSynchronize(ctx->stream)) ;
    m->total = total ;
    MAP_TRY(cudaMalloc(&csrc, std::max<uint64_t>(total, 1)*sizeof(uint32_t))) ;
    MAP_TRY(cudaMalloc(&m->placed, std::max<uint64_t>((uint64_t)total*SS, 1)*sizeof(double))) ;
    MAP_TRY(cudaMemsetAsync(count, 0, (nnzb+1)*sizeof(uint32_t), ctx->stream)) ;       // reused as the fill cursor
    MAP_TRY(cudaMemsetAsync(m->pos_of_src, 0xFF, std::max<uint64_t>(nsrc, 1)*sizeof(uint32_t), ctx->stream)) ;   // NO_DEST
    if(nsrc)
    {
        // the lists (element blocks per stored block, ascending = element order) exist only to number the stream
        k_map_fill<<<vec_grid(ctx, nsrc), AMIE_VEC_THREADS, 0, ctx->stream>>>(m->dest_of_src, nsrc, m->cptr, count, csrc) ;
        k_map_sort<<<vec_grid(ctx, nnzb), AMIE_VEC_THREADS, 0, ctx->stream>>>(m->cptr, csrc, nnzb) ;
        if(total) k_map_positions<<<vec_grid(ctx, total), AMIE_VEC_THREADS, 0, ctx->stream>>>(csrc, total, m->pos_of_src) ;
    }
    MAP_TRY(cudaGetLastError()) ;
    MAP_TRY(cudaStreamSynchronize(ctx->stream)) ;
#undef MAP_TRY
    cleanup() ;
    m->built = true ;
    m->all_dirty = true ;
    m->have_ke = false ;
    ctx->stats.elements_ms = (wall_now()-t0)*1e3 ;
    ctx->stats.element_blocks = nsrc ;
    return AMIE_B200_OK ;
}

int amie_b200_update_elements(amie_b200_ctx * ctx, uint64_t first, uint64_t count, const double * ke, const double * scales)
{
    if(!ctx || (!ke && count)) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_update_elements(ctx, first, count, ke, scales) ;
    AssemblyMap * m = ctx->amap ;
    if(!m || !m->built) { ctx->set_error("update_elements before set_elements") ; return AMIE_B200_ERR_STATE ; }
    if(first+count > m->n_elem) { ctx->set_error("update_elements: element range out of bounds") ; return AMIE_B200_ERR_ARG ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    const uint64_t pp = (uint64_t)m->npe*m->npe, SS = (uint64_t)ctx->S*ctx->S ;
    if(!m->have_ke)
    {
        // elements never uploaded contribute nothing until they are
        CUDA_TRY(ctx, cudaMemsetAsync(m->placed, 0, std::max<uint64_t>(m->total*SS, 1)*sizeof(double), ctx->stream)) ;
        m->have_ke = true ;
    }
    if(!count) return AMIE_B200_OK ;
    if(!m->stage)
    {
        CUDA_TRY(ctx, cudaMalloc(&m->stage, STAGE_DOUBLES*sizeof(double))) ;
        CUDA_TRY(ctx, cudaMalloc(&m->stage_scales, (STAGE_DOUBLES/SS+1)*sizeof(double))) ;
    }
    // host array -> staging -> its places in the stream, one chunk of whole elements at a time
    const uint64_t per_elem = pp*SS ;
    const uint64_t chunk = std::max<uint64_t>(1, STAGE_DOUBLES/per_elem) ;
    if(per_elem > STAGE_DOUBLES) { ctx->set_error("update_elements: one elementary matrix exceeds the staging buffer") ; return AMIE_B200_ERR_UNSUPPORTED ; }
    for(uint64_t e0 = 0 ; e0 < count ; e0 += chunk)
    {
        const uint64_t ne = std::min(chunk, count-e0) ;
        CUDA_TRY(ctx, cudaMemcpyAsync(m->stage, ke+e0*per_elem, ne*per_elem*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
        if(scales) CUDA_TRY(ctx, cudaMemcpyAsync(m->stage_scales, scales+e0, ne*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
        const double * sc = scales ? m->stage_scales : nullptr ;
        const int g = vec_grid(ctx, ne*per_elem) ;
        const uint64_t src0 = (first+e0)*pp, ns = ne*pp ;
#define PLACE(N) k_place_elements<N><<<g, AMIE_VEC_THREADS, 0, ctx->stream>>>(m->stage, sc, first+e0, m->pos_of_src, src0, ns, (uint32_t)pp, m->placed)
        switch(ctx->S)
        {
            case 1: PLACE(1) ; break ;
            case 2: PLACE(4) ; break ;
            case 3: PLACE(9) ; break ;
            case 4: PLACE(16) ; break ;
            case 6: PLACE(36) ; break ;
            default: ctx->set_error("update_elements: unsupported stride") ; return AMIE_B200_ERR_UNSUPPORTED ;
        }
#undef PLACE
    }
    if(!m->all_dirty)
        k_mark_dirty<<<vec_grid(ctx, count*pp), AMIE_VEC_THREADS, 0, ctx->stream>>>(m->dest_of_src, first*pp, (first+count)*pp, m->dirty) ;
    CUDA_TRY(ctx, cudaGetLastError()) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    ctx->stats.h2d_bytes = count*pp*SS*sizeof(double)+count*sizeof(double) ;
    return AMIE_B200_OK ;
}

int amie_b200_assemble(amie_b200_ctx * ctx)
{
    if(!ctx) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_assemble(ctx) ;
    AssemblyMap * m = ctx->amap ;
    if(!m || !m->built || !m->have_ke) { ctx->set_error("assemble before set_elements + update_elements") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    const uint64_t SS = (uint64_t)ctx->S*ctx->S, nent = ctx->nnzb*SS ;
    const int all = m->all_dirty ? 1 : 0 ;
    const int grid = vec_grid(ctx, nent) ;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, ctx->stream)) ;
    if(nent)
    {
#define GATHER(N) k_assemble_gather<N><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(m->cptr, m->placed, m->dirty, all, ctx->vals, nent)
        switch(ctx->S)
        {
            case 1: GATHER(1) ; break ;
            case 2: GATHER(4) ; break ;
            case 3: GATHER(9) ; break ;
            case 4: GATHER(16) ; break ;
            case 6: GATHER(36) ; break ;
            default: ctx->set_error("assemble: unsupported stride") ; return AMIE_B200_ERR_UNSUPPORTED ;
        }
#undef GATHER
        k_clear_dirty<<<vec_grid(ctx, ctx->nnzb), AMIE_VEC_THREADS, 0, ctx->stream>>>(m->dirty, ctx->nnzb) ;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, ctx->stream)) ;
    CUDA_TRY(ctx, cudaGetLastError()) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    float ms = 0.f ;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b)) ;
    ctx->stats.assemble_ms = ms ;
    m->all_dirty = false ;
    ctx->have_values = true ;
    ctx->dinv_valid = false ;
    return AMIE_B200_OK ;
}

int amie_b200_set_boundary_conditions(amie_b200_ctx * ctx, uint64_t nfix, const uint32_t * fix_ids, const double * fix_values,
                                      uint64_t nforce, const uint32_t * force_ids, const double * force_values,
                                      const double * add_to_forces, double * natural_inout)
{
    if(!ctx || (nfix && (!fix_ids || !fix_values)) || (nforce && (!force_ids || !force_values))) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_set_boundary_conditions(ctx, nfix, fix_ids, fix_values, nforce, force_ids, force_values, add_to_forces, natural_inout) ;
    int rc = require_structure(ctx, "set_boundary_conditions") ;
    if(rc) return rc ;
    if(!ctx->have_values || !ctx->have_rhs)
    { ctx->set_error("set_boundary_conditions needs the matrix values (set_values / assemble) and the force vector (upload_rhs)") ; return AMIE_B200_ERR_STATE ; }
    if(ctx->S > 8) { ctx->set_error("set_boundary_conditions: stride > 8") ; return AMIE_B200_ERR_UNSUPPORTED ; }
    // the id lists are GLOBAL dof ids; on one part of a partitioned matrix add_to_forces / natural_inout are the part's rows
    const PartMap pm = part_map(ctx) ;
    const uint64_t n_global = (uint64_t)pm.nb_global*ctx->S ;
    const uint64_t ncols = std::max<uint64_t>(ctx->ncols_local, ctx->nb) ;
    if(!ascending_unique(fix_ids, nfix, n_global) || !ascending_unique(force_ids, nforce, n_global))
    { ctx->set_error("set_boundary_conditions: dof ids must be ascending, unique and < N (Assembly sorts its multipliers by id)") ; return AMIE_B200_ERR_ARG ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    if(!ctx->amap) ctx->amap = new AssemblyMap ;              // only the mask scratch is used
    AssemblyMap * m = ctx->amap ;
    if(m->mask_nb != ncols)
    {
        afree(m->fixmask) ; afree(m->forcemask) ;
        m->mask_nb = 0 ;
        CUDA_TRY(ctx, cudaMalloc(&m->fixmask, std::max<uint64_t>(ncols, 1))) ;
        CUDA_TRY(ctx, cudaMalloc(&m->forcemask, std::max<uint64_t>(ncols, 1))) ;
        m->mask_nb = ncols ;
    }
    uint32_t * d_ids = nullptr ;
    double * d_vals = nullptr, * d_add = nullptr, * d_nat = nullptr ;
    const uint64_t nm = nfix+nforce ;
    auto cleanup = [&]() { afree(d_ids) ; afree(d_vals) ; afree(d_add) ; afree(d_nat) ; } ;
#define BC_TRY(expr) do { cudaError_t _e = (expr) ; if(_e != cudaSuccess) { cleanup() ; \
        ctx->set_error(std::string(#expr)+": "+cudaGetErrorString(_e)) ; return AMIE_B200_ERR_CUDA ; } } while(0)
    BC_TRY(cudaMalloc(&d_ids, std::max<uint64_t>(nm, 1)*sizeof(uint32_t))) ;
    BC_TRY(cudaMalloc(&d_vals, std::max<uint64_t>(nm, 1)*sizeof(double))) ;
    if(nfix)
    {
        BC_TRY(cudaMemcpyAsync(d_ids, fix_ids, nfix*sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream)) ;
        BC_TRY(cudaMemcpyAsync(d_vals, fix_values, nfix*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
        BC_TRY(cudaMemsetAsync(m->fixmask, 0, ncols, ctx->stream)) ;
        k_bc_mask<<<vec_grid(ctx, nfix), AMIE_VEC_THREADS, 0, ctx->stream>>>(d_ids, nfix, ctx->S, pm, m->fixmask) ;
    }
    if(nforce)
    {
        BC_TRY(cudaMemcpyAsync(d_ids+nfix, force_ids, nforce*sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream)) ;
        BC_TRY(cudaMemcpyAsync(d_vals+nfix, force_values, nforce*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
        BC_TRY(cudaMemsetAsync(m->forcemask, 0, ncols, ctx->stream)) ;
        k_bc_mask<<<vec_grid(ctx, nforce), AMIE_VEC_THREADS, 0, ctx->stream>>>(d_ids+nfix, nforce, ctx->S, pm, m->forcemask) ;
    }
    if(add_to_forces)
    {
        BC_TRY(cudaMalloc(&d_add, ctx->N*sizeof(double))) ;
        BC_TRY(cudaMemcpyAsync(d_add, add_to_forces, ctx->N*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    }
    if(natural_inout)
    {
        BC_TRY(cudaMalloc(&d_nat, ctx->N*sizeof(double))) ;
        BC_TRY(cudaMemcpyAsync(d_nat, natural_inout, ctx->N*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    }
    unsigned char * dirty = (m->built && !m->all_dirty) ? m->dirty : nullptr ;
    BC_TRY(cudaEventRecord(ctx->ev_a, ctx->stream)) ;
    const int grid = vec_grid(ctx, ctx->N) ;
#define DIRICHLET(N) k_dirichlet<N><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(ctx->rowptr, ctx->col, ctx->nb, pm, ctx->vals, ctx->b, d_nat, d_add, \
        m->fixmask, d_ids, d_vals, (uint32_t)nfix, m->forcemask, d_ids+nfix, d_vals+nfix, (uint32_t)nforce, dirty)
    if(ctx->N)
        switch(ctx->S)
        {
            case 1: DIRICHLET(1) ; break ;
            case 2: DIRICHLET(2) ; break ;
            case 3: DIRICHLET(3) ; break ;
            case 4: DIRICHLET(4) ; break ;
            case 6: DIRICHLET(6) ; break ;
            default: cleanup() ; ctx->set_error("set_boundary_conditions: unsupported stride") ; return AMIE_B200_ERR_UNSUPPORTED ;
        }
#undef DIRICHLET
    BC_TRY(cudaEventRecord(ctx->ev_b, ctx->stream)) ;
    BC_TRY(cudaGetLastError()) ;
    if(natural_inout)
        BC_TRY(cudaMemcpyAsync(natural_inout, d_nat, ctx->N*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    BC_TRY(cudaStreamSynchronize(ctx->stream)) ;
    float ms = 0.f ;
    BC_TRY(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b)) ;
#undef BC_TRY
    cleanup() ;
    ctx->stats.bc_ms = ms ;
    ctx->dinv_valid = false ;
    return AMIE_B200_OK ;
}

}
