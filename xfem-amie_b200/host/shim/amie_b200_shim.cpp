// amie_b200_shim.cpp -- per-Assembly device contexts for the drop-in translation units.
#include "amie_b200_shim.h"
#include "solvers/preconditionners.h"
#include "solvers/inversediagonal.h"
#include <map>
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
    if(e.stride != A.stride || e.nb != nb || e.nnzb != nnzb || e.colptr != cp || e.colhash != h)
    {
        int rc = amie_b200_set_structure(e.ctx, (int)A.stride, nb, &A.row_size[0], cp, nnzb) ;
        if(rc)
        {
            std::cerr << "amie_b200: set_structure: " << amie_b200_last_error(e.ctx) << std::endl ;
            return nullptr ;
        }
        e.stride = A.stride ; e.nb = nb ; e.nnzb = nnzb ; e.colptr = cp ; e.colhash = h ;
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

bool upload_diagonal(amie_b200_ctx * ctx, const Vector * diagonal, size_t ndof)
{
    if(!diagonal) return true ;
    if(diagonal->size() != ndof)
    {
        std::cerr << "amie_b200: the preconditioner's diagonal has " << diagonal->size() << " entries for " << ndof << " degrees of freedom" << std::endl ;
        return false ;
    }
    if(amie_b200_set_preconditioner_diagonal(ctx, &(*diagonal)[0]))
    {
        std::cerr << "amie_b200: set_preconditioner_diagonal: " << amie_b200_last_error(ctx) << std::endl ;
        return false ;
    }
    return true ;
}

}
