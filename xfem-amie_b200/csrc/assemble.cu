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
// Traffic: gather = element blocks once (8 s^2 B each) + their 4 B list entries + the stored blocks written once;
// elimination = column indices + a per-node mask byte, values only where a fixed dof is involved.
#include "context.h"
#include "launch.cuh"
#include <algorithm>
#include <vector>
#include <cub/device/device_scan.cuh>

#define NO_NODE 0xFFFFFFFFu
#define NO_DEST 0xFFFFFFFFu

struct AssemblyMap
{
    uint64_t n_elem = 0 ;
    int npe = 0 ;
    uint64_t nsrc = 0 ;                 // n_elem*npe*npe element blocks
    uint32_t * dest_of_src = nullptr ;  // [nsrc] stored block each element block lands on (NO_DEST: unused node slot)
    uint32_t * cptr = nullptr ;         // [nnzb+1] contribution lists per stored block
    uint32_t * csrc = nullptr ;         // [ncontrib] element-block indices, ascending within a list
    double * ke = nullptr ;             // [nsrc*S*S] elementary matrices (blocks column-major)
    double * scales = nullptr ;         // [n_elem]
    unsigned char * dirty = nullptr ;   // [nnzb] stored blocks to re-accumulate at the next assemble
    bool built = false ;                // set_elements done (the struct also carries the BC scratch alone)
    bool all_dirty = true ;
    bool have_ke = false ;
    // boundary-condition scratch (sized on demand)
    unsigned char * fixmask = nullptr ; // [nb] bit n: dof n of the node is eliminated
    unsigned char * forcemask = nullptr ;
    uint64_t mask_nb = 0 ;
} ;

template<typename T> static void afree(T *& p) { if(p) cudaFree(p) ; p = nullptr ; }

void assembly_map_destroy(amie_b200_ctx * ctx)
{
    AssemblyMap * m = ctx->amap ;
    if(!m) return ;
    afree(m->dest_of_src) ; afree(m->cptr) ; afree(m->csrc) ; afree(m->ke) ; afree(m->scales) ; afree(m->dirty) ;
    afree(m->fixmask) ; afree(m->forcemask) ;
    delete m ;
    ctx->amap = nullptr ;
}

// ---------------------------------------------------------------------------------------------------- map build

// element block (e, j, k) -> stored block (ids[j], ids[k]); counts the contributions of every stored block
static __global__ void k_map_dest(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ col, uint32_t nb,
                                  const uint32_t * __restrict__ ids, uint64_t nsrc, int npe,
                                  uint32_t * __restrict__ dest_of_src, uint32_t * __restrict__ count, int * __restrict__ flag)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    const uint32_t pp = (uint32_t)(npe*npe) ;
    for(uint64_t src = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; src < nsrc ; src += stride)
    {
        const uint64_t e = src/pp ;
        const uint32_t jk = (uint32_t)(src-e*pp) ;
        const uint32_t rj = __ldg(ids+e*npe+jk/npe), ck = __ldg(ids+e*npe+jk%npe) ;
        uint32_t d = NO_DEST ;
        if(rj != NO_NODE && ck != NO_NODE)
        {
            if(rj >= nb || ck >= nb) { *flag = 1 ; }
            else
            {
                const uint32_t k1 = __ldg(rowptr+rj+1) ;
                const uint32_t k = row_lower_bound(col, __ldg(rowptr+rj), k1, ck) ;
                if(k < k1 && __ldg(col+k) == ck) { d = k ; atomicAdd(count+k, 1u) ; }
                else *flag = 2 ;
            }
        }
        dest_of_src[src] = d ;
    }
}

static __global__ void k_map_fill(const uint32_t * __restrict__ dest_of_src, uint64_t nsrc, const uint32_t * __restrict__ cptr,
                                  uint32_t * __restrict__ cursor, uint32_t * __restrict__ csrc)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    for(uint64_t src = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; src < nsrc ; src += stride)
    {
        const uint32_t d = dest_of_src[src] ;
        if(d == NO_DEST) continue ;
        const uint32_t pos = atomicAdd(cursor+d, 1u) ;
        csrc[__ldg(cptr+d)+pos] = (uint32_t)src ;
    }
}

// the atomics above fill each list in arbitrary order: sort it (lists are a handful of entries long)
static __global__ void k_map_sort(const uint32_t * __restrict__ cptr, uint32_t * __restrict__ csrc, uint64_t nnzb)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    for(uint64_t d = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; d < nnzb ; d += stride)
    {
        const uint32_t p0 = cptr[d], p1 = cptr[d+1] ;
        for(uint32_t i = p0+1 ; i < p1 ; i++)
        {
            const uint32_t v = csrc[i] ;
            uint32_t j = i ;
            while(j > p0 && csrc[j-1] > v) { csrc[j] = csrc[j-1] ; j-- ; }
            csrc[j] = v ;
        }
    }
}

// ---------------------------------------------------------------------------------------------------- gather

static __global__ void k_mark_dirty(const uint32_t * __restrict__ dest_of_src, uint64_t src0, uint64_t src1,
                                    unsigned char * __restrict__ dirty)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    for(uint64_t src = src0+(uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; src < src1 ; src += stride)
    {
        const uint32_t d = dest_of_src[src] ;
        if(d != NO_DEST) dirty[d] = 1 ;
    }
}

// one thread per stored entry: replay `y = scale*Ke - c ; t = a + y ; c = (t - a) - y ; a = t` over the block's
// contributions in element order (solvers/assembly.cpp:685-690).  Explicit _rn intrinsics: no FMA contraction.
template<int SS>
static __global__ void k_assemble_gather(const uint32_t * __restrict__ cptr, const uint32_t * __restrict__ csrc,
                                         const double * __restrict__ ke, const double * __restrict__ scales,
                                         uint32_t pp, unsigned char * __restrict__ dirty, int all,
                                         double * __restrict__ vals, uint64_t nent)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    for(uint64_t idx = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; idx < nent ; idx += stride)
    {
        const uint64_t d = idx/SS ;
        const uint32_t ent = (uint32_t)(idx-d*SS) ;
        if(!all && !dirty[d]) continue ;
        const uint32_t p0 = __ldg(cptr+d), p1 = __ldg(cptr+d+1) ;
        double a = 0., c = 0. ;
        for(uint32_t p = p0 ; p < p1 ; p++)
        {
            const uint32_t src = __ldg(csrc+p) ;
            const double sc = __ldg(scales+src/pp) ;
            const double y = __dsub_rn(__dmul_rn(sc, ld_stream(ke+(uint64_t)src*SS+ent)), c) ;
            const double t = __dadd_rn(a, y) ;
            c = __dsub_rn(__dsub_rn(t, a), y) ;
            a = t ;
        }
        vals[idx] = a ;
    }
}

static __global__ void k_clear_dirty(unsigned char * __restrict__ dirty, uint64_t n)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < n ; i += stride) dirty[i] = 0 ;
}

// ---------------------------------------------------------------------------------------------------- elimination

// ids ascending and unique: the first thread of every node gathers the node's bits (no atomics)
static __global__ void k_bc_mask(const uint32_t * __restrict__ ids, uint64_t n, int S, unsigned char * __restrict__ mask)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < n ; i += stride)
    {
        const uint32_t node = ids[i]/S ;
        if(i && ids[i-1]/S == node) continue ;
        unsigned int bits = 0 ;
        for(uint64_t j = i ; j < n && ids[j]/S == node ; j++) bits |= 1u << (ids[j]-node*S) ;
        mask[node] = (unsigned char)bits ;
    }
}

__device__ __forceinline__ double bc_value(const uint32_t * __restrict__ ids, const double * __restrict__ values,
                                           uint32_t n, uint32_t id)
{
    const uint32_t k = row_lower_bound(ids, 0, n, id) ;
    return values[k] ;
}

// One thread per scalar row (node k, component m).  It walks the row's blocks in storage order and, inside each
// block, the multipliers of the row's node ("in line", solvers/assembly.cpp:170-207) and then those of the column's
// node ("in block", :210-253), ascending -- the order in which the reference updates externalForces[k*S+m].
template<int S>
static __global__ void k_dirichlet(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ col, uint64_t nb,
                                   double * __restrict__ vals, double * __restrict__ forces, double * __restrict__ natural,
                                   const double * __restrict__ add_to_forces,
                                   const unsigned char * __restrict__ fixmask, const uint32_t * __restrict__ fix_ids,
                                   const double * __restrict__ fix_values, uint32_t nfix,
                                   const unsigned char * __restrict__ forcemask, const uint32_t * __restrict__ force_ids,
                                   const double * __restrict__ force_values, uint32_t nforce,
                                   unsigned char * __restrict__ dirty)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    const uint64_t nrows = nb*S ;
    for(uint64_t row = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; row < nrows ; row += stride)
    {
        const uint32_t k = (uint32_t)(row/S) ;
        const int m = (int)(row-(uint64_t)k*S) ;
        const unsigned int rm = nfix ? fixmask[k] : 0u ;
        double f = forces[row] ;
        double nat = natural ? natural[row] : 0. ;
        const uint32_t k0 = __ldg(rowptr+k), k1 = __ldg(rowptr+k+1) ;
        for(uint32_t l = k0 ; nfix && l < k1 ; l++)
        {
            const uint32_t cb = __ldg(col+l) ;
            const unsigned int cm = fixmask[cb] ;
            if(!(rm | cm)) continue ;
            double * B = vals+(uint64_t)l*(S*S) ;
            if(dirty) dirty[l] = 1 ;
            for(int n0 = 0 ; n0 < S ; n0++)                     // multipliers of the row's node
            {
                if(!((rm >> n0) & 1u)) continue ;
                if(n0 == m)
                {
                    for(int n = 0 ; n < S ; n++) B[n*S+m] = (cb == k && n == m) ? 1. : 0. ;
                }
                else if(cb == k)
                {
                    const double v = bc_value(fix_ids, fix_values, nfix, k*S+n0) ;
                    const double val = B[n0*S+m] ;
                    const double prod = __dmul_rn(v, val) ;
                    f = __dsub_rn(f, prod) ;
                    nat = __dsub_rn(nat, prod) ;
                    B[n0*S+m] = 0. ;
                }
            }
            for(int n0 = 0 ; n0 < S ; n0++)                     // multipliers of the column's node
            {
                if(!((cm >> n0) & 1u)) continue ;
                const double v = bc_value(fix_ids, fix_values, nfix, cb*S+n0) ;
                if(cb == k && n0 == m)
                {
                    f = v ;
                    for(int n = 0 ; n < S ; n++) B[n*S+m] = (n == m) ? 1. : 0. ;
                }
                else
                {
                    const double val = B[n0*S+m] ;
                    const double prod = __dmul_rn(v, val) ;
                    f = __dsub_rn(f, prod) ;
                    nat = __dsub_rn(nat, prod) ;
                    B[n0*S+m] = 0. ;
                }
            }
        }
        if(nforce && ((forcemask[k] >> m) & 1u))                // SET_FORCE_*: externalForces[id] += value (:262-268)
            f = __dadd_rn(f, bc_value(force_ids, force_values, nforce, (uint32_t)row)) ;
        if(add_to_forces)                                       // externalForces += addToExternalForces (:323-324)
            f = __dadd_rn(f, ((rm >> m) & 1u) ? 0. : add_to_forces[row]) ;
        forces[row] = f ;
        if(natural) natural[row] = nat ;
    }
}

// ---------------------------------------------------------------------------------------------------- API

static int require_single(amie_b200_ctx * ctx, const char * what)
{
    if(ctx->dist) { ctx->set_error(std::string(what)+": not available on a row-partitioned context yet") ; return AMIE_B200_ERR_UNSUPPORTED ; }
    if(!ctx->have_structure) { ctx->set_error(std::string(what)+" before set_structure") ; return AMIE_B200_ERR_STATE ; }
    return AMIE_B200_OK ;
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
    int rc = require_single(ctx, "set_elements") ;
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
    uint32_t * ids = nullptr, * count = nullptr ;
    void * tmp = nullptr ;
    size_t tmp_bytes = 0 ;
    int bad = 0 ;
    auto cleanup = [&]() { afree(ids) ; afree(count) ; if(tmp) cudaFree(tmp) ; tmp = nullptr ; } ;
#define MAP_TRY(expr) do { cudaError_t _e = (expr) ; if(_e != cudaSuccess) { cleanup() ; assembly_map_destroy(ctx) ; \
        ctx->set_error(std::string(#expr)+": "+cudaGetErrorString(_e)) ; return AMIE_B200_ERR_CUDA ; } } while(0)
    MAP_TRY(cudaMalloc(&ids, std::max<uint64_t>(n_elem*npe, 1)*sizeof(uint32_t))) ;
    MAP_TRY(cudaMalloc(&count, (nnzb+1)*sizeof(uint32_t))) ;
    MAP_TRY(cudaMalloc(&m->dest_of_src, std::max<uint64_t>(nsrc, 1)*sizeof(uint32_t))) ;
    MAP_TRY(cudaMalloc(&m->cptr, (nnzb+1)*sizeof(uint32_t))) ;
    MAP_TRY(cudaMalloc(&m->dirty, std::max<uint64_t>(nnzb, 1))) ;
    MAP_TRY(cudaMalloc(&m->ke, std::max<uint64_t>(nsrc*SS, 1)*sizeof(double))) ;
    MAP_TRY(cudaMalloc(&m->scales, std::max<uint64_t>(n_elem, 1)*sizeof(double))) ;
    MAP_TRY(cudaMemcpyAsync(ids, elem_ids, n_elem*npe*sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream)) ;
    MAP_TRY(cudaMemsetAsync(count, 0, (nnzb+1)*sizeof(uint32_t), ctx->stream)) ;
    MAP_TRY(cudaMemsetAsync(ctx->flag, 0, sizeof(int), ctx->stream)) ;
    if(nsrc)
        k_map_dest<<<vec_grid(ctx, nsrc), AMIE_VEC_THREADS, 0, ctx->stream>>>(ctx->rowptr, ctx->col, (uint32_t)ctx->nb, ids, nsrc, npe,
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
    MAP_TRY(cudaMalloc(&m->csrc, std::max<uint64_t>(total, 1)*sizeof(uint32_t))) ;
    MAP_TRY(cudaMemsetAsync(count, 0, (nnzb+1)*sizeof(uint32_t), ctx->stream)) ;       // reused as the fill cursor
    if(nsrc)
    {
        k_map_fill<<<vec_grid(ctx, nsrc), AMIE_VEC_THREADS, 0, ctx->stream>>>(m->dest_of_src, nsrc, m->cptr, count, m->csrc) ;
        k_map_sort<<<vec_grid(ctx, nnzb), AMIE_VEC_THREADS, 0, ctx->stream>>>(m->cptr, m->csrc, nnzb) ;
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
    AssemblyMap * m = ctx->amap ;
    if(!m || !m->built) { ctx->set_error("update_elements before set_elements") ; return AMIE_B200_ERR_STATE ; }
    if(first+count > m->n_elem) { ctx->set_error("update_elements: element range out of bounds") ; return AMIE_B200_ERR_ARG ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    const uint64_t pp = (uint64_t)m->npe*m->npe, SS = (uint64_t)ctx->S*ctx->S ;
    if(!m->have_ke)
    {
        // elements never uploaded contribute nothing until they are (scale 1, Ke 0)
        CUDA_TRY(ctx, cudaMemsetAsync(m->ke, 0, std::max<uint64_t>(m->nsrc*SS, 1)*sizeof(double), ctx->stream)) ;
        std::vector<double> ones(m->n_elem, 1.) ;
        CUDA_TRY(ctx, cudaMemcpyAsync(m->scales, ones.data(), m->n_elem*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
        m->have_ke = true ;
    }
    if(!count) return AMIE_B200_OK ;
    CUDA_TRY(ctx, cudaMemcpyAsync(m->ke+first*pp*SS, ke, count*pp*SS*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    if(scales)
        CUDA_TRY(ctx, cudaMemcpyAsync(m->scales+first, scales, count*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    else
    {
        std::vector<double> ones(count, 1.) ;
        CUDA_TRY(ctx, cudaMemcpyAsync(m->scales+first, ones.data(), count*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
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
    AssemblyMap * m = ctx->amap ;
    if(!m || !m->built || !m->have_ke) { ctx->set_error("assemble before set_elements + update_elements") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    const uint64_t SS = (uint64_t)ctx->S*ctx->S, nent = ctx->nnzb*SS ;
    const uint32_t pp = (uint32_t)(m->npe*m->npe) ;
    const int all = m->all_dirty ? 1 : 0 ;
    const int grid = vec_grid(ctx, nent) ;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, ctx->stream)) ;
    if(nent)
    {
#define GATHER(N) k_assemble_gather<N><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(m->cptr, m->csrc, m->ke, m->scales, pp, m->dirty, all, ctx->vals, nent)
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
    int rc = require_single(ctx, "set_boundary_conditions") ;
    if(rc) return rc ;
    if(!ctx->have_values || !ctx->have_rhs)
    { ctx->set_error("set_boundary_conditions needs the matrix values (set_values / assemble) and the force vector (upload_rhs)") ; return AMIE_B200_ERR_STATE ; }
    if(ctx->S > 8) { ctx->set_error("set_boundary_conditions: stride > 8") ; return AMIE_B200_ERR_UNSUPPORTED ; }
    if(!ascending_unique(fix_ids, nfix, ctx->N) || !ascending_unique(force_ids, nforce, ctx->N))
    { ctx->set_error("set_boundary_conditions: dof ids must be ascending, unique and < N (Assembly sorts its multipliers by id)") ; return AMIE_B200_ERR_ARG ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    if(!ctx->amap) ctx->amap = new AssemblyMap ;              // only the mask scratch is used
    AssemblyMap * m = ctx->amap ;
    if(m->mask_nb != ctx->nb)
    {
        afree(m->fixmask) ; afree(m->forcemask) ;
        CUDA_TRY(ctx, cudaMalloc(&m->fixmask, std::max<uint64_t>(ctx->nb, 1))) ;
        CUDA_TRY(ctx, cudaMalloc(&m->forcemask, std::max<uint64_t>(ctx->nb, 1))) ;
        m->mask_nb = ctx->nb ;
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
        BC_TRY(cudaMemsetAsync(m->fixmask, 0, ctx->nb, ctx->stream)) ;
        k_bc_mask<<<vec_grid(ctx, nfix), AMIE_VEC_THREADS, 0, ctx->stream>>>(d_ids, nfix, ctx->S, m->fixmask) ;
    }
    if(nforce)
    {
        BC_TRY(cudaMemcpyAsync(d_ids+nfix, force_ids, nforce*sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream)) ;
        BC_TRY(cudaMemcpyAsync(d_vals+nfix, force_values, nforce*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
        BC_TRY(cudaMemsetAsync(m->forcemask, 0, ctx->nb, ctx->stream)) ;
        k_bc_mask<<<vec_grid(ctx, nforce), AMIE_VEC_THREADS, 0, ctx->stream>>>(d_ids+nfix, nforce, ctx->S, m->forcemask) ;
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
#define DIRICHLET(N) k_dirichlet<N><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(ctx->rowptr, ctx->col, ctx->nb, ctx->vals, ctx->b, d_nat, d_add, \
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
