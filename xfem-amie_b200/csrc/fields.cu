// fields.cu -- per-element field recovery from the displacement field.
//
// What it replaces: the step right after the Krylov solve.  FeatureTree::stepElements hands the solution to every
// element (ElementState::step, elements/integrable_entity.cpp:3607-3667: a gather of the element's dofs) and the
// behaviours / post-processors then ask ElementState::getField for
//   TOTAL_STRAIN_FIELD       :977-1104   (2D :979-1017, 3D :1019-1095)
//   MECHANICAL_STRAIN_FIELD  :964-975    (total strain - imposed strain)
//   REAL_STRESS_FIELD        :1379-1392  (tensor * mechanical strain - imposed stress; utilities/matrixops.h:545-558)
// one element and one virtual call at a time.  Here: one thread per element, the solution read straight from the
// resident x of the solve (no download of u), all three fields in one pass.
//
// Inputs that stay upstream (the reference's polynomial VM and geometry produce them once per topology): the
// shape-function derivatives at the evaluation point and the element's cached inverse Jacobian
// (ElementState::JinvCache).  Enrichment functions are just further slots of an element (:1002-1011, :1049-1068).
//
// Bit-exactness: every product and sum is an explicit _rn intrinsic in the reference's order (the reference is
// built without FMA contraction), so the results equal ElementState::getField's bit for bit
// (tests/test_gpu_recovery.py against tests/golden/AMIE-*-fields.npz, which the unmodified reference produced).
//
// Layout in HBM: per-element operands are stored component-major ([component][element]) so that consecutive
// threads read consecutive addresses; behaviours (tensor + imposed strain/stress) are a table indexed per element
// -- a handful of entries for an undamaged composite (L1/L2 resident), one per element under damage.  Results are
// staged through shared memory and written element-major (the layout the caller wants) with coalesced stores.
// Bound: HBM.  Algorithmic bytes of one launch (dim d, npe slots, nc = 3|6 components, nb nodes):
//   n_elem * (4 npe (ids) + 8 npe d (derivatives) + 8 d^2 (Jinv) + 4 (behaviour index) + 3*8 nc (results))
//   + 8 d nb (the solution, read once: the per-slot gathers hit L2)
//   = 332 B per linear tetrahedron, 152 B per linear triangle, + 8 nc (nc+2) per element with per-element behaviours.
// Measured (profiles/r01b_ncu_element_fields.txt, 5.82 M tetrahedra): 0.43 ms with a 2-entry table = 4.6 TB/s (70 %
// of the measured copy peak), 0.83 ms with per-element behaviours = 5.0 TB/s (77 %); DRAM traffic 2.08 / 4.39 GB
// against 1.96 / 4.19 GB algorithmic; stalls are all long_scoreboard at 50 % occupancy (56 registers).
#include "context.h"
#include "group.h"
#include "kernels_fields.cuh"
#include <algorithm>
#include <vector>


struct FieldMap
{
    uint64_t n_elem = 0 ;
    int npe = 0, dim = 0, nc = 0 ;
    uint32_t * ids = nullptr ;        // [npe][n_elem]
    double * dshape = nullptr ;       // [npe*dim][n_elem]
    double * jinv = nullptr ;         // [dim*dim][n_elem]
    uint64_t n_tensors = 0 ;
    double * tensors = nullptr ;      // [n_tensors][nc*nc] row-major
    double * istrain = nullptr ;      // [n_tensors][nc] (zeros when the caller passed NULL)
    double * istress = nullptr ;
    uint32_t * tensor_of_elem = nullptr ; // [n_elem] or nullptr (identity)
    double * out[3] = {nullptr, nullptr, nullptr} ;   // total strain, mechanical strain, real stress: [n_elem][nc]
    double * principal = nullptr ;    // [n_elem][dim] scratch of element_principal
    bool have_fields = false ;        // element_fields ran since the last kinematics / behaviour change
    double * u_tmp = nullptr ;        // staging for a host-supplied displacement field
    uint64_t u_tmp_len = 0 ;
    bool have_behaviour = false ;
} ;

template<typename T> static void ffree(T *& p) { if(p) cudaFree(p) ; p = nullptr ; }

void field_map_destroy(amie_b200_ctx * ctx)
{
    FieldMap * m = ctx->fmap ;
    if(!m) return ;
    ffree(m->ids) ; ffree(m->dshape) ; ffree(m->jinv) ; ffree(m->tensors) ; ffree(m->istrain) ; ffree(m->istress) ;
    ffree(m->tensor_of_elem) ; ffree(m->u_tmp) ; ffree(m->principal) ;
    for(int i = 0 ; i < 3 ; i++) ffree(m->out[i]) ;
    delete m ;
    ctx->fmap = nullptr ;
}

uint64_t field_map_bytes(const amie_b200_ctx * ctx)
{
    const FieldMap * m = ctx->fmap ;
    if(!m) return 0 ;
    const uint64_t nc = m->nc ;
    uint64_t b = m->n_elem*((uint64_t)m->npe*4+(uint64_t)m->npe*m->dim*8+(uint64_t)m->dim*m->dim*8+3*nc*8) ;
    b += m->n_tensors*(nc*nc+2*nc)*8 ;
    if(m->tensor_of_elem) b += m->n_elem*4 ;
    if(m->principal) b += m->n_elem*m->dim*8 ;
    b += m->u_tmp_len*8 ;
    return b ;
}

static int field_grid(const amie_b200_ctx * ctx, uint64_t n_elem)
{
    const uint64_t tiles = (n_elem+FIELD_THREADS-1)/FIELD_THREADS ;
    const uint64_t cap = (uint64_t)ctx->num_sms*8 ;
    return (int)std::max<uint64_t>(1, std::min(tiles, cap)) ;
}

extern "C" {

int amie_b200_set_element_kinematics(amie_b200_ctx * ctx, uint64_t n_elem, int npe, int dim, const uint32_t * elem_ids,
                                     const double * dshape, const double * jinv)
{
    if(!ctx || npe < 1 || npe > 64 || (n_elem && (!elem_ids || !dshape || !jinv))) return AMIE_B200_ERR_ARG ;
    // A multi-device context splits the ELEMENTS into one contiguous range per device (group.cu); each device then holds
    // a full-length copy of the displacement field, so element ids stay global and any element can be computed anywhere.
    if(ctx->group) return group_set_element_kinematics(ctx, n_elem, npe, dim, elem_ids, dshape, jinv) ;
    if(!ctx->have_structure) { ctx->set_error("set_element_kinematics before set_structure") ; return AMIE_B200_ERR_STATE ; }
    if((dim != 2 && dim != 3) || dim != ctx->S)
    {
        // the reference's strain code exists for 2 dofs per node in 2D and 3 in 3D only (elements/integrable_entity.cpp:979, :1019)
        ctx->set_error("set_element_kinematics: dim must be 2 or 3 and equal the stride of the system") ;
        return AMIE_B200_ERR_UNSUPPORTED ;
    }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    field_map_destroy(ctx) ;
    FieldMap * m = new FieldMap ;
    ctx->fmap = m ;
    m->n_elem = n_elem ; m->npe = npe ; m->dim = dim ; m->nc = dim == 2 ? 3 : 6 ;
    void * tmp = nullptr ;
    const uint64_t one = 1 ;
    const uint64_t tmp_bytes = std::max(one, n_elem*(uint64_t)std::max(npe*dim, dim*dim)*sizeof(double)) ;
#define F_TRY(expr) do { cudaError_t _e = (expr) ; if(_e != cudaSuccess) { if(tmp) cudaFree(tmp) ; field_map_destroy(ctx) ; \
        ctx->set_error(std::string(#expr)+": "+cudaGetErrorString(_e)) ; return AMIE_B200_ERR_CUDA ; } } while(0)
    F_TRY(cudaMalloc(&tmp, tmp_bytes)) ;
    F_TRY(cudaMalloc(&m->ids, std::max(one, n_elem*npe)*sizeof(uint32_t))) ;
    F_TRY(cudaMalloc(&m->dshape, std::max(one, n_elem*npe*dim)*sizeof(double))) ;
    F_TRY(cudaMalloc(&m->jinv, std::max(one, n_elem*dim*dim)*sizeof(double))) ;
    for(int i = 0 ; i < 3 ; i++) F_TRY(cudaMalloc(&m->out[i], std::max(one, n_elem*m->nc)*sizeof(double))) ;
    if(n_elem)
    {
        F_TRY(cudaMemcpyAsync(tmp, elem_ids, n_elem*npe*sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream)) ;
        k_to_component_major<uint32_t><<<vec_grid(ctx, n_elem*npe), AMIE_VEC_THREADS, 0, ctx->stream>>>((const uint32_t *)tmp, m->ids, n_elem, npe) ;
        F_TRY(cudaStreamSynchronize(ctx->stream)) ;
        F_TRY(cudaMemcpyAsync(tmp, dshape, n_elem*npe*dim*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
        k_to_component_major<double><<<vec_grid(ctx, n_elem*npe*dim), AMIE_VEC_THREADS, 0, ctx->stream>>>((const double *)tmp, m->dshape, n_elem, npe*dim) ;
        F_TRY(cudaStreamSynchronize(ctx->stream)) ;
        F_TRY(cudaMemcpyAsync(tmp, jinv, n_elem*dim*dim*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
        k_to_component_major<double><<<vec_grid(ctx, n_elem*dim*dim), AMIE_VEC_THREADS, 0, ctx->stream>>>((const double *)tmp, m->jinv, n_elem, dim*dim) ;
        F_TRY(cudaGetLastError()) ;
        F_TRY(cudaStreamSynchronize(ctx->stream)) ;
    }
#undef F_TRY
    cudaFree(tmp) ;
    return AMIE_B200_OK ;
}

int amie_b200_set_element_behaviour(amie_b200_ctx * ctx, uint64_t n_tensors, const double * tensors,
                                    const double * imposed_strain, const double * imposed_stress,
                                    const uint32_t * tensor_of_elem)
{
    if(!ctx || !n_tensors || !tensors) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_set_element_behaviour(ctx, n_tensors, tensors, imposed_strain, imposed_stress, tensor_of_elem) ;
    FieldMap * m = ctx->fmap ;
    if(!m) { ctx->set_error("set_element_behaviour before set_element_kinematics") ; return AMIE_B200_ERR_STATE ; }
    if(!tensor_of_elem && n_tensors != m->n_elem)
    { ctx->set_error("set_element_behaviour: without tensor_of_elem there must be one behaviour per element") ; return AMIE_B200_ERR_ARG ; }
    if(tensor_of_elem)
        for(uint64_t e = 0 ; e < m->n_elem ; e++)
            if(tensor_of_elem[e] >= n_tensors) { ctx->set_error("set_element_behaviour: tensor_of_elem entry out of range") ; return AMIE_B200_ERR_ARG ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    const uint64_t nc = m->nc ;
    if(m->n_tensors != n_tensors)
    {
        ffree(m->tensors) ; ffree(m->istrain) ; ffree(m->istress) ;
        m->n_tensors = 0 ; m->have_behaviour = false ;
        CUDA_TRY(ctx, cudaMalloc(&m->tensors, n_tensors*nc*nc*sizeof(double))) ;
        CUDA_TRY(ctx, cudaMalloc(&m->istrain, n_tensors*nc*sizeof(double))) ;
        CUDA_TRY(ctx, cudaMalloc(&m->istress, n_tensors*nc*sizeof(double))) ;
        m->n_tensors = n_tensors ;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(m->tensors, tensors, n_tensors*nc*nc*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    // absent imposed terms are zeros: x - 0 == x bit for bit, which is what the reference computes for behaviours
    // without induced forces (it skips the subtraction, :967) and for the zero vector the base class returns (:1392)
    if(imposed_strain) CUDA_TRY(ctx, cudaMemcpyAsync(m->istrain, imposed_strain, n_tensors*nc*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    else CUDA_TRY(ctx, cudaMemsetAsync(m->istrain, 0, n_tensors*nc*sizeof(double), ctx->stream)) ;
    if(imposed_stress) CUDA_TRY(ctx, cudaMemcpyAsync(m->istress, imposed_stress, n_tensors*nc*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
    else CUDA_TRY(ctx, cudaMemsetAsync(m->istress, 0, n_tensors*nc*sizeof(double), ctx->stream)) ;
    if(tensor_of_elem)
    {
        if(!m->tensor_of_elem) CUDA_TRY(ctx, cudaMalloc(&m->tensor_of_elem, std::max<uint64_t>(1, m->n_elem)*sizeof(uint32_t))) ;
        CUDA_TRY(ctx, cudaMemcpyAsync(m->tensor_of_elem, tensor_of_elem, m->n_elem*sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream)) ;
    }
    else ffree(m->tensor_of_elem) ;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    m->have_behaviour = true ;
    m->have_fields = false ;
    return AMIE_B200_OK ;
}

}   // extern "C"

// the device buffer a displacement field of n doubles is staged in (host-supplied fields; on a multi-device context the
// resident solution gathered from every part)
int fields_u_buffer(amie_b200_ctx * ctx, uint64_t n, double ** out)
{
    FieldMap * m = ctx->fmap ;
    if(!m) { ctx->set_error("element_fields before set_element_kinematics") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    if(m->u_tmp_len < n)
    {
        ffree(m->u_tmp) ; m->u_tmp_len = 0 ;
        CUDA_TRY(ctx, cudaMalloc(&m->u_tmp, std::max<uint64_t>(n, 1)*sizeof(double))) ;
        m->u_tmp_len = n ;
    }
    *out = m->u_tmp ;
    return AMIE_B200_OK ;
}

// du: the displacement field on this device, `len` doubles, indexed by the (global) node ids of the elements
int fields_run(amie_b200_ctx * ctx, const double * du, uint64_t len, uint64_t h2d,
               double * total_strain_out, double * mechanical_strain_out, double * real_stress_out)
{
    FieldMap * m = ctx->fmap ;
    if(!m) { ctx->set_error("element_fields before set_element_kinematics") ; return AMIE_B200_ERR_STATE ; }
    if(!m->have_behaviour) { ctx->set_error("element_fields before set_element_behaviour") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, ctx->stream)) ;
    if(m->n_elem)
    {
        const int grid = field_grid(ctx, m->n_elem) ;
#define FIELDS(D, P) k_element_fields<D, P><<<grid, FIELD_THREADS, 0, ctx->stream>>>(m->ids, m->dshape, m->jinv, m->tensors, m->istrain, m->istress, \
                    m->tensor_of_elem, du, len, m->n_elem, m->npe, m->out[0], m->out[1], m->out[2])
        // linear triangles / tetrahedra take the unrolled, phase-split instantiations (kernels_fields.cuh: 18 loads in
        // flight before the first multiply; 71 -> 77 % and 78 -> 80 % of the measured peak, same bits,
        // profiles/r02_notes.md); option "fields_variant" = 0 forces the generic slot loop (tests)
        const bool unrolled = ctx->opt_fields_variant != 0 ;
        if(m->dim == 2) { if(unrolled && m->npe == 3) FIELDS(2, 3) ; else FIELDS(2, 0) ; }
        else            { if(unrolled && m->npe == 4) FIELDS(3, 4) ; else FIELDS(3, 0) ; }
#undef FIELDS
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, ctx->stream)) ;
    CUDA_TRY(ctx, cudaGetLastError()) ;
    const uint64_t bytes = m->n_elem*m->nc*sizeof(double) ;
    double * host[3] = { total_strain_out, mechanical_strain_out, real_stress_out } ;
    uint64_t d2h = 0 ;
    for(int i = 0 ; i < 3 ; i++)
        if(host[i] && bytes)
        {
            CUDA_TRY(ctx, cudaMemcpyAsync(host[i], m->out[i], bytes, cudaMemcpyDeviceToHost, ctx->stream)) ;
            d2h += bytes ;
        }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    float ms = 0.f ;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b)) ;
    ctx->stats.fields_ms = ms ;
    ctx->stats.field_elements = m->n_elem ;
    m->have_fields = true ;
    ctx->stats.h2d_bytes = h2d ;
    ctx->stats.d2h_bytes = d2h ;
    return AMIE_B200_OK ;
}

extern "C" {

int amie_b200_element_fields(amie_b200_ctx * ctx, const double * u, uint64_t n_u,
                             double * total_strain_out, double * mechanical_strain_out, double * real_stress_out)
{
    if(!ctx || (!u && n_u)) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_element_fields(ctx, u, n_u, total_strain_out, mechanical_strain_out, real_stress_out) ;
    FieldMap * m = ctx->fmap ;
    if(!m) { ctx->set_error("element_fields before set_element_kinematics") ; return AMIE_B200_ERR_STATE ; }
    if(!m->have_behaviour) { ctx->set_error("element_fields before set_element_behaviour") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    const double * du = ctx->x ;        // the resident solution of the last solve
    uint64_t len = ctx->N ;
    uint64_t h2d = 0 ;
    if(u)
    {
        double * buf = nullptr ;
        int rc = fields_u_buffer(ctx, n_u, &buf) ;
        if(rc) return rc ;
        CUDA_TRY(ctx, cudaMemcpyAsync(buf, u, n_u*sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) ;
        du = buf ; len = n_u ; h2d = n_u*sizeof(double) ;
    }
    else if(ctx->dist)
    {
        // one rank of a partitioned matrix only holds its own rows of the solution; element ids are global
        ctx->set_error("element_fields on one rank of a row-partitioned matrix: pass the displacement field (all rows)") ;
        return AMIE_B200_ERR_UNSUPPORTED ;
    }
    return fields_run(ctx, du, len, h2d, total_strain_out, mechanical_strain_out, real_stress_out) ;
}

int amie_b200_element_principal(amie_b200_ctx * ctx, int field, double * principal_out)
{
    if(!ctx || !principal_out || field < 0 || field > 2) return AMIE_B200_ERR_ARG ;
    if(ctx->group) return group_element_principal(ctx, field, principal_out) ;
    FieldMap * m = ctx->fmap ;
    if(!m || !m->have_fields) { ctx->set_error("element_principal before element_fields") ; return AMIE_B200_ERR_STATE ; }
    CUDA_TRY(ctx, cudaSetDevice(ctx->device)) ;
    if(!m->principal) CUDA_TRY(ctx, cudaMalloc(&m->principal, std::max<uint64_t>(1, m->n_elem*m->dim)*sizeof(double))) ;
    if(m->n_elem)
    {
        const int grid = vec_grid(ctx, m->n_elem) ;
        const bool strain = field != 2 ;           // strains carry engineering shears: DOUBLE_OFF_DIAGONAL_VALUES
        if(m->dim == 2 && strain)       k_element_principal<2, true><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(m->out[field], m->principal, m->n_elem) ;
        else if(m->dim == 2)            k_element_principal<2, false><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(m->out[field], m->principal, m->n_elem) ;
        else if(strain)                 k_element_principal<3, true><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(m->out[field], m->principal, m->n_elem) ;
        else                            k_element_principal<3, false><<<grid, AMIE_VEC_THREADS, 0, ctx->stream>>>(m->out[field], m->principal, m->n_elem) ;
        CUDA_TRY(ctx, cudaGetLastError()) ;
        CUDA_TRY(ctx, cudaMemcpyAsync(principal_out, m->principal, m->n_elem*m->dim*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) ;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)) ;
    return AMIE_B200_OK ;
}

}
