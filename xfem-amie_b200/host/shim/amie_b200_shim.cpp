// amie_b200_shim.cpp -- per-Assembly device contexts for the drop-in translation units.
#include "amie_b200_shim.h"
#include "solvers/preconditionners.h"
#include "solvers/inversediagonal.h"
#include <map>
#include <vector>
#include <iostream>
#include <cstdlib>
#include <cstring>
#include <algorithm>

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
    bool have_values = false ;
    uint64_t valhash = 0 ;                  // of the array last uploaded
    uint64_t last_use = 0 ;
} ;
std::map<Amie::Assembly *, Entry> registry ;
uint64_t use_clock = 0 ;

// Order-dependent 64-bit hash of EVERY word (multiply-xorshift per word, four independent lanes per chunk, chunks
// in parallel and then folded in order).  It decides whether an upload can be skipped, so nothing is sampled.
uint64_t hash_words(const uint64_t * p, size_t n, uint64_t seed)
{
    const size_t chunk = 1u << 16 ;
    const size_t nchunks = (n+chunk-1)/chunk ;
    std::vector<uint64_t> part(nchunks) ;
    #pragma omp parallel for schedule(static)
    for(long long c = 0 ; c < (long long)nchunks ; c++)
    {
        const size_t a = (size_t)c*chunk, b = std::min(n, a+chunk) ;
        uint64_t h[4] = { seed^0x9e3779b97f4a7c15ull, seed^0xbf58476d1ce4e5b9ull, seed^0x94d049bb133111ebull, seed^0x2545f4914f6cdd1dull } ;
        size_t i = a ;
        for( ; i+4 <= b ; i += 4)
            for(int l = 0 ; l < 4 ; l++) { h[l] = (h[l]^p[i+l])*0x100000001b3ull ; h[l] ^= h[l] >> 29 ; }
        for( ; i < b ; i++) { h[0] = (h[0]^p[i])*0x100000001b3ull ; h[0] ^= h[0] >> 29 ; }
        part[c] = ((h[0]*31+h[1])*31+h[2])*31+h[3] ;
    }
    uint64_t h = seed^n ;
    for(uint64_t v : part) { h = (h^v)*0xff51afd7ed558ccdull ; h ^= h >> 33 ; }
    return h ;
}

uint64_t hash_u32(const unsigned int * p, size_t n, uint64_t seed)
{
    uint64_t h = hash_words(reinterpret_cast<const uint64_t *>(p), n/2, seed) ;      // valarray storage is 8-byte aligned
    if(n & 1) h = (h^p[n-1])*0x100000001b3ull ;
    return h^n ;
}

// Assembly::cgsolve builds a fresh solver per call and nothing tells this file when an Assembly dies, so the registry
// is a small cache: beyond AMIE_B200_MAX_CONTEXTS (default 2) the least recently used context is destroyed and its HBM
// released.  An evicted Assembly that solves again simply uploads again.
void evict_beyond_capacity(Amie::Assembly * keep)
{
    size_t cap = 2 ;
    if(const char * e = getenv("AMIE_B200_MAX_CONTEXTS")) cap = (size_t)std::max(1, atoi(e)) ;
    while(registry.size() > cap)
    {
        auto victim = registry.end() ;
        for(auto it = registry.begin() ; it != registry.end() ; ++it)
            if(it->first != keep && (victim == registry.end() || it->second.last_use < victim->second.last_use)) victim = it ;
        if(victim == registry.end()) break ;
        amie_b200_destroy(victim->second.ctx) ;
        registry.erase(victim) ;
    }
}

struct AtExit { ~AtExit() { for(auto & kv : registry) amie_b200_destroy(kv.second.ctx) ; registry.clear() ; } } at_exit ;
}

namespace AmieB200Shim
{

amie_b200_ctx * context_for(Amie::Assembly * a)
{
    Amie::CoordinateIndexedSparseMatrix & A = a->getMatrix() ;
    Entry & e = registry[a] ;
    e.last_use = ++use_clock ;
    if(!e.ctx)
    {
        evict_beyond_capacity(a) ;
        // one device (AMIE_B200_DEVICE) or several behind the same calls (AMIE_B200_DEVICES=0,1,...: the block rows are
        // partitioned inside the library, csrc/group.cu)
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
    const uint64_t h = hash_u32(cp, nnzb, 1)^hash_u32(nb ? &A.row_size[0] : nullptr, nb, 2) ;
    // Renumbering (reverse Cuthill-McKee, once per topology): AMIE's mesher numbering has no locality, and the x gather
    // of the SpMV pays for it -- measured on a 2.65 M-DOF tetrahedral system with a numbering without locality:
    // 2.8 TB/s as it comes, 5.3 TB/s renumbered, PCG 2 126 -> 3 622 it/s (profiles/r02_notes.md).  On by default from
    // 20 000 nodes up (smaller solves are launch-bound); AMIE_B200_RENUMBER=0 / 1 forces it off / on.
    // rowstart / colstart address AMIE's numbering (space-time planes): such assemblies keep it.  A multi-device
    // context (AMIE_B200_DEVICES) keeps it by default too -- it partitions the rows as they come -- and renumbers when
    // AMIE_B200_RENUMBER=1 asks for it (contiguous row ranges of a Cuthill-McKee numbering have small halos).
    const char * env = getenv("AMIE_B200_RENUMBER") ;
    const char * devs = getenv("AMIE_B200_DEVICES") ;
    const bool multi = devs && strchr(devs, ',') ;
    const bool asked = env ? atoi(env) != 0 : (nb >= 20000 && !multi) ;
    const bool want_renumber = asked && a->rowstart == 0 && a->colstart == 0 && nb > 0 ;
    if(e.stride != A.stride || e.nb != nb || e.nnzb != nnzb || e.colptr != cp || e.colhash != h || e.renumbered != want_renumber)
    {
        int rc = 0 ;
        const char * what = "set_structure" ;
        if(want_renumber)
        {
            e.perm.assign(nb, 0u) ;
            std::vector<uint32_t> rs2(nb), ci2(nnzb), from(nnzb), to(nnzb) ;
            rc = amie_b200_rcm_order(nb, &A.row_size[0], cp, e.perm.data()) ;
            // AMIE_B200_RENUMBER=2 (opt-in): rows of similar length next to one another inside windows of 80 nodes of the
            // Cuthill-McKee numbering (a tile of the SpMV costs its longest row; csrc/reorder.cpp)
            if(!rc && env && atoi(env) == 2) rc = amie_b200_group_rows_by_length(nb, &A.row_size[0], 80, e.perm.data()) ;
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
        e.have_values = false ;
    }
    // Values change whenever make_final re-scatters them (every damage step) -- but not between the CG, CG, BiCGStab
    // solves of one FeatureTree::step, nor between load steps of an elastic run.  A full hash of the array (host memory
    // speed, all cores) costs far less than pushing it over PCIe again.  AMIE_B200_ALWAYS_UPLOAD=1 turns the check off.
    const char * always = getenv("AMIE_B200_ALWAYS_UPLOAD") ;
    const bool check = !(always && atoi(always) != 0) ;
    const uint64_t vh = check ? hash_words(reinterpret_cast<const uint64_t *>(&A.array[0]), A.array.size(), 3) : 0 ;
    if(!check || !e.have_values || vh != e.valhash)
    {
        e.have_values = false ;
        if(amie_b200_set_values(e.ctx, &A.array[0]))
        {
            std::cerr << "amie_b200: set_values: " << amie_b200_last_error(e.ctx) << std::endl ;
            return nullptr ;
        }
        e.have_values = check ;
        e.valhash = vh ;
        if(getenv("AMIE_B200_SHIM_TRACE"))
        {
            amie_b200_stats st ;
            if(amie_b200_get_stats(e.ctx, &st) == 0)
                std::cerr << "amie_b200: matrix upload: structure " << st.structure_ms << " ms, values " << st.values_ms << " ms" << std::endl ;
        }
    }
    else if(getenv("AMIE_B200_SHIM_TRACE"))
        std::cerr << "amie_b200: matrix unchanged since the last solve: no upload" << std::endl ;
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
    // Inverse2x2Diagonal (solvers/inversediagonal.cpp:84-133): built on the device from the matrix being solved -- the
    // usual `Inverse2x2Diagonal P(A)` on the assembly's own matrix (an object built from ANOTHER matrix is not told apart)
    if(dynamic_cast<Amie::Inverse2x2Diagonal *>(p)) return AMIE_B200_PRECOND_BLOCK2X2 ;
    if(!p)
    {
        // opt-in: node-block Jacobi where the caller passes nullptr (Assembly::cgsolve always does).  Other iteration
        // counts than the reference's InverseDiagonal: AMIE_B200_BLOCK_JACOBI=1 says the host accepts that.
        const char * e = getenv("AMIE_B200_BLOCK_JACOBI") ;
        if(e && atoi(e) != 0) return -2 ;          // resolved by stride in the solver shim
        return AMIE_B200_PRECOND_JACOBI ;
    }
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
