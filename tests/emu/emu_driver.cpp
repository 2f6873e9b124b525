// emu_driver.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).  C entry points that run the product's kernel sources
// on the host in the launch sequences of the C-ABI functions they belong to (assemble.cu, fields.cu, cgsolve.cu,
// api.cu: ctx_ensure_dinv / set_values).  The grids are small on purpose: every grid-stride loop wraps many times.
#include "cuda_emu.h"
#include "kernels_setup.cuh"
#include "kernels_assemble.cuh"
#include "kernels_fields.cuh"
#include "kernels_history.cuh"
#include <string.h>

static const unsigned GRID = 3, BLOCK = 64 ;

static std::vector<uint32_t> rowptr_of(uint64_t nb, const uint32_t * row_size)
{
    std::vector<uint32_t> rp(nb+1, 0) ;
    for(uint64_t i = 0 ; i < nb ; i++) rp[i+1] = rp[i]+row_size[i] ;
    return rp ;
}

#define BY_STRIDE(S, CALL) switch(S) { case 1: { constexpr int N = 1 ; CALL ; } break ; case 2: { constexpr int N = 2 ; CALL ; } break ; \
    case 3: { constexpr int N = 3 ; CALL ; } break ; case 4: { constexpr int N = 4 ; CALL ; } break ; case 6: { constexpr int N = 6 ; CALL ; } break ; default: return -5 ; }

extern "C" {

// set_values: K-Repack (padded reference layout -> compact), strides 1 and 3; the other strides carry no pad
int emu_repack(int stride, const double * padded, double * compact, uint64_t nblocks)
{
    if(stride == 3)      emu_launch(GRID, BLOCK, [&]() { k_repack<3>(padded, compact, nblocks) ; }) ;
    else if(stride == 1) emu_launch(GRID, BLOCK, [&]() { k_repack<1>(padded, compact, nblocks) ; }) ;
    else memcpy(compact, padded, nblocks*stride*stride*sizeof(double)) ;
    return 0 ;
}

// set_values with a block map (amie_b200_set_block_map): block k of the host array -> stored block block_to[k]
int emu_repack_scatter(int stride, const double * padded, const uint32_t * block_to, uint64_t nblocks, double * compact)
{
    BY_STRIDE(stride, emu_launch(GRID, BLOCK, [&]() { k_repack_scatter<N>(padded, compact, block_to, nblocks) ; }))
    return 0 ;
}

// ctx_ensure_dinv: kind 0 InverseDiagonal, 2 InverseDiagonalSquared, 3 InverseLumpedDiagonal
int emu_precond_diagonal(int kind, int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * col,
                         const double * vals, double * d)
{
    std::vector<uint32_t> rp = rowptr_of(nb, row_size) ;
    if(kind == 0)      { BY_STRIDE(stride, emu_launch(GRID, BLOCK, [&]() { k_inverse_diagonal<N>(rp.data(), col, vals, 0u, nb, d) ; })) }
    else if(kind == 2) { BY_STRIDE(stride, emu_launch(GRID, BLOCK, [&]() { k_inverse_diagonal_squared<N>(rp.data(), col, vals, nb, d) ; })) }
    else if(kind == 3) { BY_STRIDE(stride, emu_launch(GRID, BLOCK, [&]() { k_inverse_lumped_diagonal<N>(rp.data(), col, vals, nb, d) ; })) }
    else return -2 ;
    return 0 ;
}

// set_elements + update_elements + assemble (assemble.cu).  vals (compact) is input and output: with all == 0 only the
// stored blocks touched by elements [mark_first, mark_first+mark_count) are re-accumulated (k_mark_dirty), like an
// incremental damage step.  Returns 0, 1 (node id out of range), 2 (pair outside the pattern).
static int emu_assemble_pm(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * col, uint64_t nnzb, PartMap pm,
                 uint64_t n_elem, int npe, const uint32_t * ids, const double * ke, const double * scales,
                 int all, uint64_t mark_first, uint64_t mark_count, double * vals) ;

int emu_assemble(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * col, uint64_t nnzb,
                 uint64_t n_elem, int npe, const uint32_t * ids, const double * ke, const double * scales,
                 int all, uint64_t mark_first, uint64_t mark_count, double * vals)
{
    const PartMap pm = { 0u, (uint32_t)nb, nullptr, 0u, (uint32_t)nb } ;
    return emu_assemble_pm(stride, nb, row_size, col, nnzb, pm, n_elem, npe, ids, ke, scales, all, mark_first, mark_count, vals) ;
}

// one part of a row-partitioned matrix (dist.cu numbering): nb local rows = global rows [row_base, row_base+nb), col holds
// LOCAL block columns (owned, then nb + position in the sorted halo list); ids stay GLOBAL, every part sees every element
int emu_assemble_part(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * col, uint64_t nnzb,
                      uint64_t row_base, const uint32_t * halo, uint64_t nhalo, uint64_t nb_global,
                      uint64_t n_elem, int npe, const uint32_t * ids, const double * ke, const double * scales,
                      int all, uint64_t mark_first, uint64_t mark_count, double * vals)
{
    const PartMap pm = { (uint32_t)row_base, (uint32_t)nb, halo, (uint32_t)nhalo, (uint32_t)nb_global } ;
    return emu_assemble_pm(stride, nb, row_size, col, nnzb, pm, n_elem, npe, ids, ke, scales, all, mark_first, mark_count, vals) ;
}

static int emu_assemble_pm(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * col, uint64_t nnzb, PartMap pm,
                 uint64_t n_elem, int npe, const uint32_t * ids, const double * ke, const double * scales,
                 int all, uint64_t mark_first, uint64_t mark_count, double * vals)
{
    std::vector<uint32_t> rp = rowptr_of(nb, row_size) ;
    const uint64_t nsrc = n_elem*(uint64_t)npe*npe ;
    const uint32_t pp = (uint32_t)(npe*npe) ;
    std::vector<uint32_t> dest(nsrc ? nsrc : 1), count(nnzb+1, 0), cptr(nnzb+1, 0) ;
    int flag = 0 ;
    emu_launch(GRID, BLOCK, [&]() { k_map_dest(rp.data(), col, pm, ids, nsrc, npe, dest.data(), count.data(), &flag) ; }) ;
    if(flag) return flag ;
    for(uint64_t k = 0 ; k < nnzb ; k++) cptr[k+1] = cptr[k]+count[k] ;          // cub::DeviceScan::ExclusiveSum
    std::vector<uint32_t> csrc(cptr[nnzb] ? cptr[nnzb] : 1) ;
    std::fill(count.begin(), count.end(), 0u) ;
    // the device fills the lists in arbitrary order; scramble the order here too (blocks visited backwards)
    gridDim.x = GRID ; blockDim.x = BLOCK ;
    for(int b = (int)GRID-1 ; b >= 0 ; b--)
        for(unsigned t = 0 ; t < BLOCK ; t++)
        {
            blockIdx.x = (unsigned)b ; threadIdx.x = t ;
            k_map_fill(dest.data(), nsrc, cptr.data(), count.data(), csrc.data()) ;
        }
    emu_launch(GRID, BLOCK, [&]() { k_map_sort(cptr.data(), csrc.data(), nnzb) ; }) ;
    // update_elements: the stream numbering (k_map_positions), then the elementary matrices placed in two uneven chunks
    // of whole elements (the device streams the host array through a staging buffer)
    const uint64_t total = cptr[nnzb] ;
    const uint64_t SSu = (uint64_t)stride*stride ;
    std::vector<uint32_t> pos(nsrc ? nsrc : 1, NO_DEST) ;
    emu_launch(GRID, BLOCK, [&]() { k_map_positions(csrc.data(), total, pos.data()) ; }) ;
    std::vector<double> placed(total*SSu ? total*SSu : 1, 0.) ;
    const uint64_t cut = n_elem/3 ;
    for(int part = 0 ; part < 2 ; part++)
    {
        const uint64_t e0 = part ? cut : 0, ne = part ? n_elem-cut : cut ;
        if(!ne) continue ;
        const double * stage = ke+e0*pp*SSu ;
        const double * sc = scales ? scales+e0 : nullptr ;
        BY_STRIDE(stride, emu_launch(GRID, BLOCK, [&]() { k_place_elements<N*N>(stage, sc, e0, pos.data(), e0*pp, ne*pp, pp, placed.data()) ; }))
    }
    std::vector<unsigned char> dirty(nnzb ? nnzb : 1, 0) ;
    if(!all)
        emu_launch(GRID, BLOCK, [&]() { k_mark_dirty(dest.data(), mark_first*pp, (mark_first+mark_count)*pp, dirty.data()) ; }) ;
    const uint64_t nent = nnzb*(uint64_t)stride*stride ;
    BY_STRIDE(stride, emu_launch(GRID, BLOCK, [&]() { k_assemble_gather<N*N>(cptr.data(), placed.data(), dirty.data(), all, vals, nent) ; }))
    emu_launch(GRID, BLOCK, [&]() { k_clear_dirty(dirty.data(), nnzb) ; }) ;
    for(uint64_t k = 0 ; k < nnzb ; k++) if(dirty[k]) return -1 ;
    return 0 ;
}

// set_boundary_conditions (assemble.cu): masks, then the row-owned elimination; dirty_out (nnzb, may be NULL) receives
// the stored blocks the elimination touched
static int emu_dirichlet_pm(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * col, uint64_t nnzb, PartMap pm,
                  double * vals, double * forces, double * natural, const double * add_to_forces,
                  uint64_t nfix, const uint32_t * fix_ids, const double * fix_values,
                  uint64_t nforce, const uint32_t * force_ids, const double * force_values, unsigned char * dirty_out) ;

int emu_dirichlet(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * col, uint64_t nnzb,
                  double * vals, double * forces, double * natural, const double * add_to_forces,
                  uint64_t nfix, const uint32_t * fix_ids, const double * fix_values,
                  uint64_t nforce, const uint32_t * force_ids, const double * force_values, unsigned char * dirty_out)
{
    const PartMap pm = { 0u, (uint32_t)nb, nullptr, 0u, (uint32_t)nb } ;
    return emu_dirichlet_pm(stride, nb, row_size, col, nnzb, pm, vals, forces, natural, add_to_forces, nfix, fix_ids, fix_values,
                            nforce, force_ids, force_values, dirty_out) ;
}

// one part (see emu_assemble_part): vals / forces / natural / add_to_forces are the part's LOCAL rows, the id lists GLOBAL
int emu_dirichlet_part(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * col, uint64_t nnzb,
                       uint64_t row_base, const uint32_t * halo, uint64_t nhalo, uint64_t nb_global,
                       double * vals, double * forces, double * natural, const double * add_to_forces,
                       uint64_t nfix, const uint32_t * fix_ids, const double * fix_values,
                       uint64_t nforce, const uint32_t * force_ids, const double * force_values, unsigned char * dirty_out)
{
    const PartMap pm = { (uint32_t)row_base, (uint32_t)nb, halo, (uint32_t)nhalo, (uint32_t)nb_global } ;
    return emu_dirichlet_pm(stride, nb, row_size, col, nnzb, pm, vals, forces, natural, add_to_forces, nfix, fix_ids, fix_values,
                            nforce, force_ids, force_values, dirty_out) ;
}

static int emu_dirichlet_pm(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * col, uint64_t nnzb, PartMap pm,
                  double * vals, double * forces, double * natural, const double * add_to_forces,
                  uint64_t nfix, const uint32_t * fix_ids, const double * fix_values,
                  uint64_t nforce, const uint32_t * force_ids, const double * force_values, unsigned char * dirty_out)
{
    std::vector<uint32_t> rp = rowptr_of(nb, row_size) ;
    const uint64_t ncols = nb+pm.nhalo ;
    std::vector<unsigned char> fixmask(ncols ? ncols : 1, 0), forcemask(ncols ? ncols : 1, 0) ;
    (void)nnzb ;
    if(nfix)   emu_launch(GRID, BLOCK, [&]() { k_bc_mask(fix_ids, nfix, stride, pm, fixmask.data()) ; }) ;
    if(nforce) emu_launch(GRID, BLOCK, [&]() { k_bc_mask(force_ids, nforce, stride, pm, forcemask.data()) ; }) ;
    BY_STRIDE(stride, emu_launch(GRID, BLOCK, [&]() { k_dirichlet<N>(rp.data(), col, nb, pm, vals, forces, natural, add_to_forces,
                                                                       fixmask.data(), fix_ids, fix_values, (uint32_t)nfix,
                                                                       forcemask.data(), force_ids, force_values, (uint32_t)nforce, dirty_out) ; }))
    return 0 ;
}

// set_element_kinematics + element_fields (fields.cu)
int emu_element_fields(int dim, uint64_t n_elem, int npe, const uint32_t * ids, const double * dshape, const double * jinv,
                       const double * tensors, const double * istrain, const double * istress, const uint32_t * tensor_of_elem,
                       const double * u, uint64_t n_u, int variant, double * total, double * mech, double * stress)
{
    std::vector<uint32_t> ids_t(std::max<uint64_t>(1, n_elem*npe)) ;
    std::vector<double> ds_t(std::max<uint64_t>(1, n_elem*npe*dim)), ji_t(std::max<uint64_t>(1, n_elem*dim*dim)) ;
    emu_launch(GRID, BLOCK, [&]() { k_to_component_major<uint32_t>(ids, ids_t.data(), n_elem, npe) ; }) ;
    emu_launch(GRID, BLOCK, [&]() { k_to_component_major<double>(dshape, ds_t.data(), n_elem, npe*dim) ; }) ;
    emu_launch(GRID, BLOCK, [&]() { k_to_component_major<double>(jinv, ji_t.data(), n_elem, dim*dim) ; }) ;
    // "fields_variant" = 1: the unrolled, phase-split instantiations for linear triangles / tetrahedra
    if(variant == 1 && dim == 2 && npe == 3)
        emu_launch_sync(2, FIELD_THREADS, [&]() { k_element_fields<2, 3>(ids_t.data(), ds_t.data(), ji_t.data(), tensors, istrain, istress, tensor_of_elem,
                                                                       u, n_u, n_elem, npe, total, mech, stress) ; }) ;
    else if(variant == 1 && dim == 3 && npe == 4)
        emu_launch_sync(2, FIELD_THREADS, [&]() { k_element_fields<3, 4>(ids_t.data(), ds_t.data(), ji_t.data(), tensors, istrain, istress, tensor_of_elem,
                                                                       u, n_u, n_elem, npe, total, mech, stress) ; }) ;
    else if(dim == 2)
        emu_launch_sync(2, FIELD_THREADS, [&]() { k_element_fields<2>(ids_t.data(), ds_t.data(), ji_t.data(), tensors, istrain, istress, tensor_of_elem,
                                                                    u, n_u, n_elem, npe, total, mech, stress) ; }) ;
    else if(dim == 3)
        emu_launch_sync(2, FIELD_THREADS, [&]() { k_element_fields<3>(ids_t.data(), ds_t.data(), ji_t.data(), tensors, istrain, istress, tensor_of_elem,
                                                                    u, n_u, n_elem, npe, total, mech, stress) ; }) ;
    else return -5 ;
    return 0 ;
}

// element_principal (fields.cu)
int emu_element_principal(int dim, uint64_t n_elem, const double * in, int double_offdiag, double * out)
{
    if(dim == 2 && double_offdiag)  emu_launch(GRID, BLOCK, [&]() { k_element_principal<2, true>(in, out, n_elem) ; }) ;
    else if(dim == 2)               emu_launch(GRID, BLOCK, [&]() { k_element_principal<2, false>(in, out, n_elem) ; }) ;
    else if(dim == 3 && double_offdiag) emu_launch(GRID, BLOCK, [&]() { k_element_principal<3, true>(in, out, n_elem) ; }) ;
    else if(dim == 3)               emu_launch(GRID, BLOCK, [&]() { k_element_principal<3, false>(in, out, n_elem) ; }) ;
    else return -5 ;
    return 0 ;
}

// cgsolve.cu: Assembly::extrapolate on a two-vector history, and displacements*0.
int emu_extrapolate(const double * prev, double * back, double * x, uint64_t n, double factor)
{
    emu_launch(GRID, BLOCK, [&]() { k_extrapolate(prev, back, x, n, factor) ; }) ;
    return 0 ;
}

int emu_times_zero(const double * x, double * out, uint64_t n)
{
    emu_launch(GRID, BLOCK, [&]() { k_times_zero(x, out, n) ; }) ;
    return 0 ;
}

}
