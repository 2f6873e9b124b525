// kernels_assemble.cuh -- the kernels of the device-side value assembly and Dirichlet elimination (assemble.cu holds
// the description, the host side and the C-ABI).  Depends on device_utils.cuh only (see there).
#pragma once
#include "device_utils.cuh"

#define NO_NODE 0xFFFFFFFFu
#define NO_DEST 0xFFFFFFFFu

// ---------------------------------------------------------------------------------------------------- numbering

// The numbering of the matrix a context holds.  Single device: local == global ({0, nb, nullptr, 0}).  One part of a
// row-partitioned matrix (dist.cu): local block rows [0, nb) are the global rows row_base .. row_base+nb, and the local
// block column l >= nb is halo[l-nb] (the sorted global ids of the foreign columns the part's rows touch).  The blocks of
// a row stay in GLOBAL ascending order in storage (dist.cu renumbers the column indices in place), which is the order
// the reference walks them in -- so the kernels below search and walk rows through global_of().
struct PartMap
{
    uint32_t row_base ;
    uint32_t nb ;                 // owned block rows
    const uint32_t * halo ;
    uint32_t nhalo ;
    uint32_t nb_global ;

    __host__ __device__ uint32_t global_of(uint32_t l) const { return l < nb ? l+row_base : halo[l-nb] ; }
    __host__ __device__ bool owns(uint32_t g) const { return g >= row_base && g-row_base < nb ; }
    // local block column of global node g; NO_NODE when the part neither owns it nor has it in its halo
    __host__ __device__ uint32_t local_of(uint32_t g) const
    {
        if(owns(g)) return g-row_base ;
        uint32_t lo = 0, hi = nhalo ;
        while(lo < hi)
        {
            const uint32_t mid = lo+((hi-lo) >> 1) ;
            if(halo[mid] < g) lo = mid+1 ; else hi = mid ;
        }
        return (lo < nhalo && halo[lo] == g) ? nb+lo : 0xFFFFFFFFu ;
    }
} ;

// ---------------------------------------------------------------------------------------------------- map build

// element block (e, j, k) -> stored block (ids[j], ids[k]); counts the contributions of every stored block.
// Node ids are GLOBAL; blocks whose row belongs to another part of a partitioned matrix get NO_DEST here (that part
// builds the same map over the same element list and keeps them).
static __global__ void k_map_dest(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ col, PartMap pm,
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
            if(rj >= pm.nb_global || ck >= pm.nb_global) { *flag = 1 ; }
            else if(pm.owns(rj))
            {
                const uint32_t lr = rj-pm.row_base ;
                uint32_t k0 = __ldg(rowptr+lr) ;
                const uint32_t k1 = __ldg(rowptr+lr+1) ;
                uint32_t hi = k1 ;
                while(k0 < hi)                      // the row's blocks ascend in GLOBAL column order
                {
                    const uint32_t mid = k0+((hi-k0) >> 1) ;
                    if(pm.global_of(__ldg(col+mid)) < ck) k0 = mid+1 ; else hi = mid ;
                }
                if(k0 < k1 && pm.global_of(__ldg(col+k0)) == ck) { d = k0 ; atomicAdd(count+k0, 1u) ; }
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

// position of every element block in the PLACED stream: the contributions of stored block d occupy positions
// [cptr[d], cptr[d+1]) in list order (ascending element-block index = element order)
static __global__ void k_map_positions(const uint32_t * __restrict__ csrc, uint64_t total, uint32_t * __restrict__ pos_of_src)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    for(uint64_t p = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; p < total ; p += stride)
        pos_of_src[csrc[p]] = (uint32_t)p ;
}

// update_elements: the uploaded elementary matrices go to their places in the stream, already multiplied by the
// element's scale -- `scale*Ke` rounds the same whenever it is formed, so the sums below keep the reference's bits.
// stage holds the blocks [src0, src0+nsrc) in upload order; scales (NULL = 1) one value per element from elem0 on.
template<int SS>
static __global__ void k_place_elements(const double * __restrict__ stage, const double * __restrict__ scales, uint64_t elem0,
                                        const uint32_t * __restrict__ pos_of_src, uint64_t src0, uint64_t nsrc, uint32_t pp,
                                        double * __restrict__ placed)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    for(uint64_t idx = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; idx < nsrc*SS ; idx += stride)
    {
        const uint64_t s = idx/SS ;
        const uint32_t ent = (uint32_t)(idx-s*SS) ;
        const uint64_t src = src0+s ;
        const uint32_t pos = __ldg(pos_of_src+src) ;
        if(pos == NO_DEST) continue ;
        const double sc = scales ? __ldg(scales+(src/pp-elem0)) : 1. ;
        placed[(uint64_t)pos*SS+ent] = __dmul_rn(sc, ld_stream(stage+idx)) ;
    }
}

// one thread per stored entry: replay `y = scale*Ke - c ; t = a + y ; c = (t - a) - y ; a = t` over the block's
// contributions in element order (solvers/assembly.cpp:685-690).  Explicit _rn intrinsics: no FMA contraction.
// The contributions are one contiguous run of the placed stream: no index list, no scale load, and the 9 threads of a
// stored block (and the blocks next to it in the warp) walk neighbouring 72-byte pieces -- every sector that is
// fetched is used whole (the list-indexed version fetched 1.26x its bytes: profiles/r01c_ncu_assembly.txt).
template<int SS>
static __global__ void k_assemble_gather(const uint32_t * __restrict__ cptr, const double * __restrict__ placed,
                                         unsigned char * __restrict__ dirty, int all,
                                         double * __restrict__ vals, uint64_t nent)
{
    // ncu on the first form of this kernel (profiles/r02_ncu_assemble_gather.txt): DRAM traffic = the algorithmic bytes,
    // 97 % occupancy, and every stall a long_scoreboard -- two DEPENDENT latencies per entry (run offsets, then the
    // run).  So the offsets of the NEXT entry of the grid-stride loop are fetched while the current run is summed, and
    // the run is read two contributions at a time before the (sequential, order-preserving) compensated additions.
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    uint64_t idx = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ;
    if(idx >= nent) return ;
    uint64_t d = idx/SS ;
    uint32_t p0 = __ldg(cptr+d), p1 = __ldg(cptr+d+1) ;
    bool todo = all || dirty[d] ;
    for( ; idx < nent ; )
    {
        const uint64_t nidx = idx+stride ;
        uint64_t nd = 0 ;
        uint32_t np0 = 0, np1 = 0 ;
        bool ntodo = false ;
        if(nidx < nent)
        {
            nd = nidx/SS ;
            np0 = __ldg(cptr+nd) ; np1 = __ldg(cptr+nd+1) ;
            ntodo = all || dirty[nd] ;
        }
        if(todo)
        {
            const uint32_t ent = (uint32_t)(idx-d*SS) ;
            const double * v = placed+(uint64_t)p0*SS+ent ;
            double a = 0., c = 0. ;
            uint32_t p = p0 ;
            for( ; p+2 <= p1 ; p += 2, v += 2*SS)
            {
                // plain cached loads: the 72-byte pieces of a run share sectors from one step to the next, and with
                // the neighbouring stored blocks of the warp -- L1 serves the second touch
                const double v0 = __ldg(v), v1 = __ldg(v+SS) ;
                double y = __dsub_rn(v0, c) ;
                double t = __dadd_rn(a, y) ;
                c = __dsub_rn(__dsub_rn(t, a), y) ;
                a = t ;
                y = __dsub_rn(v1, c) ;
                t = __dadd_rn(a, y) ;
                c = __dsub_rn(__dsub_rn(t, a), y) ;
                a = t ;
            }
            if(p < p1)
            {
                const double y = __dsub_rn(__ldg(v), c) ;
                const double t = __dadd_rn(a, y) ;
                c = __dsub_rn(__dsub_rn(t, a), y) ;
                a = t ;
            }
            vals[idx] = a ;
        }
        idx = nidx ; d = nd ; p0 = np0 ; p1 = np1 ; todo = ntodo ;
    }
}

static __global__ void k_clear_dirty(unsigned char * __restrict__ dirty, uint64_t n)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < n ; i += stride) dirty[i] = 0 ;
}

// ---------------------------------------------------------------------------------------------------- elimination

// ids (GLOBAL dof ids) ascending and unique: the first thread of every node gathers the node's bits (no atomics).
// mask is indexed by LOCAL block column (owned rows, then the halo); nodes the part does not see are skipped.
static __global__ void k_bc_mask(const uint32_t * __restrict__ ids, uint64_t n, int S, PartMap pm, unsigned char * __restrict__ mask)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < n ; i += stride)
    {
        const uint32_t node = ids[i]/S ;
        if(i && ids[i-1]/S == node) continue ;
        const uint32_t local = pm.local_of(node) ;
        if(local == NO_NODE) continue ;
        unsigned int bits = 0 ;
        for(uint64_t j = i ; j < n && ids[j]/S == node ; j++) bits |= 1u << (ids[j]-node*S) ;
        mask[local] = (unsigned char)bits ;
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
// (A per-node offset into the id list instead of the binary search below was measured: same time, profiles/r02_notes.md.)
// On one part of a partitioned matrix the rows, `forces`, `natural`, `add_to_forces` and the masks are LOCAL, the id
// lists GLOBAL; a row's blocks are stored in global column order, so the updates of f keep the reference's order.
template<int S>
static __global__ void k_dirichlet(const uint32_t * __restrict__ rowptr, const uint32_t * __restrict__ col, uint64_t nb, PartMap pm,
                                   double * __restrict__ vals, double * __restrict__ forces, double * __restrict__ natural,
                                   const double * __restrict__ add_to_forces,
                                   const unsigned char * __restrict__ fixmask, const uint32_t * __restrict__ fix_ids,
                                   const double * __restrict__ fix_values, uint32_t nfix,
                                   const unsigned char * __restrict__ forcemask, const uint32_t * __restrict__ force_ids,
                                   const double * __restrict__ force_values, uint32_t nforce,
                                   unsigned char * __restrict__ dirty)
{
    // the imposed value of dof (node, n), whose bit is set in the node's mask
    auto fixed_value = [&](uint32_t node, unsigned int, int n) { return bc_value(fix_ids, fix_values, nfix, pm.global_of(node)*S+n) ; } ;
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
                    const double v = fixed_value(k, rm, n0) ;
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
                const double v = fixed_value(cb, cm, n0) ;
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
            f = __dadd_rn(f, bc_value(force_ids, force_values, nforce, (uint32_t)row+pm.row_base*S)) ;
        if(add_to_forces)                                       // externalForces += addToExternalForces (:323-324)
            f = __dadd_rn(f, ((rm >> m) & 1u) ? 0. : add_to_forces[row]) ;
        forces[row] = f ;
        if(natural) natural[row] = nat ;
    }
}
