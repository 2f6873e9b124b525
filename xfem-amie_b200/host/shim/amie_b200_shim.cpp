// amie_b200_shim.cpp -- per-Assembly device contexts for the drop-in translation units.
#include "amie_b200_shim.h"
#include "solvers/preconditionners.h"
#include "solvers/inversediagonal.h"
#include <map>
#include <vector>
#include <iostream>
#include <cstdlib>

namespace
{
struct Entry
{
    amie_b200_ctx * ctx = nullptr ;
    size_t stride = 0, nb = 0, nnzb = 0 ;
    const unsigned int * colptr = nullptr ;
    uint64_t colhash = 0 ;
    bool renumbered = false ;
    std::vector<uint32_t> perm ;            // perm[old node] = new node while renumbered
} ;
std::map<Amie::Assembly *, Entry> registry ;

uint64_t hash_u32(const unsigned int * p, size_t n)
{
    // FNV-1a over a strided sample: cheap, and any re-numbering changes it
    uint64_t h = 1469598103934665603ull ;
    size_t step = n > 65536 ? n/65536 : 1 ;
    for(size_t i = 0 ; i < n ; i += step) { h ^= p[i] ; h *= 1099511628211ull ; }
    h ^= n ;
    return h ;
}
}

namespace AmieB200Shim
{

amie_b200_ctx * context_for(Amie::Assembly * a)
{
    Amie::CoordinateIndexedSparseMatrix & A = a->getMatrix() ;
    Entry & e = registry[a] ;
    if(!e.ctx)
    {
        e.ctx = amie_b200_create(nullptr, 0) ;
        if(!e.ctx)
        {
            std::cerr << "amie_b200: " << amie_b200_global_error() << " (no CPU fallback)" << std::endl ;
            return nullptr ;
        }
        if(getenv("AMIE_B200_VERBOSE")) amie_b200_set_option(e.ctx, "verbose", 1) ;
    }
    const size_t nb = A.row_size.size(), nnzb = A.column_index.size() ;
    const unsigned int * cp = nnzb ? &A.column_index[0] : nullptr ;
    const uint64_t h = hash_u32(cp, nnzb) ;
    // rowstart / colstart address AMIE's numbering (space-time planes): such assemblies keep it
    const char * env = getenv("AMIE_B200_RENUMBER") ;
    const bool want_renumber = env && atoi(env) != 0 && a->rowstart == 0 && a->colstart == 0 && nb > 0 ;
    if(e.stride != A.stride || e.nb != nb || e.nnzb != nnzb || e.colptr != cp || e.colhash != h || e.renumbered != want_renumber)
    {
        int rc = 0 ;
        const char * what = "set_structure" ;
        if(want_renumber)
        {
            e.perm.assign(nb, 0u) ;
            std::vector<uint32_t> rs2(nb), ci2(nnzb), from(nnzb), to(nnzb) ;
            rc = amie_b200_rcm_order(nb, &A.row_size[0], cp, e.perm.data()) ;
            if(!rc) rc = amie_b200_permute_structure(nb, &A.row_size[0], cp, e.perm.data(), rs2.data(), ci2.data(), from.data()) ;
            if(rc) { std::cerr << "amie_b200: renumbering failed (" << rc << ")" << std::endl ; return nullptr ; }
            for(size_t k = 0 ; k < nnzb ; k++) to[from[k]] = (uint32_t)k ;
            rc = amie_b200_set_structure(e.ctx, (int)A.stride, nb, rs2.data(), ci2.data(), nnzb) ;
            if(!rc) { what = "set_block_map" ; rc = amie_b200_set_block_map(e.ctx, to.data()) ; }
        }
        else
        {
            e.perm.clear() ;
            rc = amie_b200_set_structure(e.ctx, (int)A.stride, nb, &A.row_size[0], cp, nnzb) ;
        }
        if(rc)
        {
            std::cerr << "amie_b200: " << what << ": " << amie_b200_last_error(e.ctx) << std::endl ;
            e.stride = 0 ;                                  // nothing usable is on the device: upload again next time
            return nullptr ;
        }
        e.stride = A.stride ; e.nb = nb ; e.nnzb = nnzb ; e.colptr = cp ; e.colhash = h ; e.renumbered = want_renumber ;
    }
    // values change on every assembly (make_final zeroes and re-scatters them): upload each solve
    if(amie_b200_set_values(e.ctx, &A.array[0]))
    {
        std::cerr << "amie_b200: set_values: " << amie_b200_last_error(e.ctx) << std::endl ;
        return nullptr ;
    }
    return e.ctx ;
}

void release(Amie::Assembly * a)
{
    auto it = registry.find(a) ;
    if(it == registry.end()) return ;
    amie_b200_destroy(it->second.ctx) ;
    registry.erase(it) ;
}

const std::vector<uint32_t> * permutation_for(Amie::Assembly * a)
{
    auto it = registry.find(a) ;
    if(it == registry.end() || !it->second.renumbered) return nullptr ;
    return &it->second.perm ;
}

void to_device_order(const std::vector<uint32_t> & perm, size_t stride, const Vector & in, Vector & out)
{
    out.resize(perm.size()*stride) ;
    out = 0. ;
    for(size_t i = 0 ; i < perm.size() ; i++)
        for(size_t m = 0 ; m < stride ; m++)
            if(i*stride+m < in.size()) out[(size_t)perm[i]*stride+m] = in[i*stride+m] ;
}

void from_device_order(const std::vector<uint32_t> & perm, size_t stride, const Vector & in, Vector & out)
{
    if(out.size() != perm.size()*stride) out.resize(perm.size()*stride) ;
    for(size_t i = 0 ; i < perm.size() ; i++)
        for(size_t m = 0 ; m < stride ; m++)
            out[i*stride+m] = in[(size_t)perm[i]*stride+m] ;
}

int precond_kind(Amie::Preconditionner * p, const Vector ** diagonal_out)
{
    *diagonal_out = nullptr ;
    if(!p) return AMIE_B200_PRECOND_JACOBI ;
    if(dynamic_cast<Amie::NullPreconditionner *>(p)) return AMIE_B200_PRECOND_NULL ;
    if(Amie::InverseDiagonal * d = dynamic_cast<Amie::InverseDiagonal *>(p)) { *diagonal_out = &d->diagonal ; return AMIE_B200_PRECOND_DIAGONAL ; }
    if(Amie::InverseLumpedDiagonal * d = dynamic_cast<Amie::InverseLumpedDiagonal *>(p)) { *diagonal_out = &d->diagonal ; return AMIE_B200_PRECOND_DIAGONAL ; }
    if(Amie::InverseDiagonalSquared * d = dynamic_cast<Amie::InverseDiagonalSquared *>(p)) { *diagonal_out = d->diagonal ; return AMIE_B200_PRECOND_DIAGONAL ; }
    return -1 ;
}

bool upload_diagonal(amie_b200_ctx * ctx, const Vector * diagonal, size_t ndof, Amie::Assembly * a)
{
    if(!diagonal) return true ;
    if(diagonal->size() != ndof)
    {
        std::cerr << "amie_b200: the preconditioner's diagonal has " << diagonal->size() << " entries for " << ndof << " degrees of freedom" << std::endl ;
        return false ;
    }
    Vector renumbered ;
    if(const std::vector<uint32_t> * perm = a ? permutation_for(a) : nullptr)
    {
        to_device_order(*perm, ndof/perm->size(), *diagonal, renumbered) ;
        diagonal = &renumbered ;
    }
    if(amie_b200_set_preconditioner_diagonal(ctx, &(*diagonal)[0]))
    {
        std::cerr << "amie_b200: set_preconditioner_diagonal: " << amie_b200_last_error(ctx) << std::endl ;
        return false ;
    }
    return true ;
}

}
