/* amie_oracle_assembly.c -- CPU restatement of the reference's value assembly and Dirichlet elimination
 * (the step right before the Krylov solve).
 *
 * TEST INFRASTRUCTURE ONLY (see amie_oracle.c).  Never on the product path.
 *
 * Parity status: PINNED.  tests/test_oracle_assembly.py checks
 *   - amie_oracle_set_boundary_conditions against the real Assembly::setBoundaryConditions
 *     (oracle/_ref/libamie_ref_oracle.so: ref_set_boundary_conditions) on random systems, bit for bit;
 *   - amie_oracle_assemble + the elimination against the matrices the unmodified FeatureTree assembled
 *     (tests/golden/AMIE-*-elements.npz, dumped by oracle/e2e_harness.cpp), bit for bit.
 *
 * Element matrices are passed as npe x npe grids of column-major s x s blocks:
 *   Ke[((e*npe + j)*npe + k)*s*s + m*s + n] == element e, getCachedElementaryMatrix()[j][k][n][m]
 * (row n, column m of the block coupling node slot j to node slot k).  A node slot whose id is
 * 0xFFFFFFFF is unused (elements with fewer nodes than npe).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NO_NODE 0xFFFFFFFFu

/* CoordinateIndexedSparseMatrix::operator[](row).getPointer(col): binary search of the block
 * (sparse/sparse_vector.h:864-870) */
static int64_t find_block(const uint64_t * acc, const uint32_t * column_index, uint64_t brow, uint32_t bcol)
{
    uint64_t lo = acc[brow], hi = acc[brow+1] ;
    while(lo < hi)
    {
        uint64_t mid = lo+(hi-lo)/2 ;
        if(column_index[mid] < bcol) lo = mid+1 ; else hi = mid ;
    }
    if(lo < acc[brow+1] && column_index[lo] == bcol) return (int64_t)lo ;
    return -1 ;
}

/* one Kahan-compensated accumulation of a scaled element block into a stored block
 * (solvers/assembly.cpp:681-697 and :705-733; 3D twin :1084-1136) */
static void add_block(double * dst, double * comp, const double * ke, double scale, int s, int cl)
{
    for(int m = 0 ; m < s ; m++)
        for(int n = 0 ; n < s ; n++)
        {
            double * a = dst+m*cl+n ;
            double * c = comp+m*cl+n ;
            double y = scale*ke[m*s+n] - *c ;
            double t = *a + y ;
            *c = (t-*a)-y ;
            *a = t ;
        }
}

/* Assembly::make_final, stiffness scatter loop: solvers/assembly.cpp:657-735 (2D), :1060-1138 (3D).
 * `array` (nnzb*s*cl doubles, reference layout) is zeroed first (`coordinateIndexedMatrix->array = 0`, :650).
 * Returns 0, or -1 when an element couples two nodes whose block is not in the sparsity pattern. */
int amie_oracle_assemble(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, uint64_t nnzb,
                         uint64_t n_elem, int npe, const uint32_t * elem_ids, const double * elem_blocks,
                         const double * scales, double * array)
{
    const int s = stride, cl = s+s%2 ;
    const uint64_t bs = (uint64_t)s*cl ;
    uint64_t * acc = (uint64_t *)malloc((nb+1)*sizeof(uint64_t)) ;
    acc[0] = 0 ;
    for(uint64_t i = 0 ; i < nb ; i++) acc[i+1] = acc[i]+row_size[i] ;
    double * c = (double *)calloc(nnzb*bs+1, sizeof(double)) ;          /* Vector c(0., array.size()) :656 */
    memset(array, 0, nnzb*bs*sizeof(double)) ;
    int rc = 0 ;
    for(uint64_t e = 0 ; e < n_elem && !rc ; e++)
    {
        const uint32_t * ids = elem_ids+e*npe ;
        const double * ke = elem_blocks+e*(uint64_t)npe*npe*s*s ;
        const double scale = scales ? scales[e] : 1. ;
        for(int j = 0 ; j < npe && !rc ; j++)
        {
            if(ids[j] == NO_NODE) continue ;
            int64_t d = find_block(acc, column_index, ids[j], ids[j]) ;
            if(d < 0) { rc = -1 ; break ; }
            add_block(array+d*bs, c+d*bs, ke+((uint64_t)j*npe+j)*s*s, scale, s, cl) ;
            for(int k = j+1 ; k < npe ; k++)
            {
                if(ids[k] == NO_NODE) continue ;
                int64_t d0 = find_block(acc, column_index, ids[j], ids[k]) ;
                int64_t d1 = find_block(acc, column_index, ids[k], ids[j]) ;
                if(d0 < 0 || d1 < 0) { rc = -1 ; break ; }
                /* the reference interleaves the two blocks entry by entry; they are distinct blocks, so
                 * doing one after the other gives the same bits */
                add_block(array+d0*bs, c+d0*bs, ke+((uint64_t)j*npe+k)*s*s, scale, s, cl) ;
                add_block(array+d1*bs, c+d1*bs, ke+((uint64_t)k*npe+j)*s*s, scale, s, cl) ;
            }
        }
    }
    free(c) ; free(acc) ;
    return rc ;
}

/* std::lower_bound / std::upper_bound over the sorted multiplier ids (int in the reference) */
static int64_t lb(const int64_t * a, int64_t lo, int64_t hi, int64_t key)
{
    while(lo < hi) { int64_t mid = lo+(hi-lo)/2 ; if(a[mid] < key) lo = mid+1 ; else hi = mid ; }
    return lo ;
}
static int64_t ub(const int64_t * a, int64_t lo, int64_t hi, int64_t key)
{
    while(lo < hi) { int64_t mid = lo+(hi-lo)/2 ; if(a[mid] <= key) lo = mid+1 ; else hi = mid ; }
    return lo ;
}

/* Assembly::setBoundaryConditions, solvers/assembly.cpp:125-330, for the multiplier kinds the solve path sees in
 * the elastic / damage drivers:
 *   - displacement-type multipliers (SET_ALONG_*, anything that is not SET_FORCE_*, SET_PROPORTIONAL_DISPLACEMENT
 *     or GENERAL): row/column elimination :137-254;
 *   - SET_FORCE_* multipliers: externalForces[id] += value :262-268;
 *   - externalForces += addToExternalForces :323-324 (entries of fixed dofs zeroed first, :177,:218).
 * `fix_ids` ascending and unique (the reference sorts its multipliers by id, :428, and erases duplicates when they are
 * added); the two lists together stand for the id-sorted multiplier vector, which is why both are searched with one
 * merged id array below.  GENERAL / PROPORTIONAL multipliers are outside this restatement.
 * `natural` (nullable) receives the same subtractions as `forces` (naturalBoundaryConditionForces). */
void amie_oracle_set_boundary_conditions(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * column_index,
                                         uint64_t nnzb, double * array, double * forces, double * natural,
                                         double * add_to_forces,
                                         uint64_t nfix, const uint32_t * fix_ids, const double * fix_values,
                                         uint64_t nforce, const uint32_t * force_ids, const double * force_values)
{
    (void)nnzb ;
    const int s = stride, cl = s+s%2 ;
    /* merged, id-sorted multiplier list (stable: a displacement multiplier before a force on the same id) */
    const uint64_t nm = nfix+nforce ;
    int64_t * mid = (int64_t *)malloc((nm+1)*sizeof(int64_t)) ;
    double * mval = (double *)malloc((nm+1)*sizeof(double)) ;
    int * mforce = (int *)malloc((nm+1)*sizeof(int)) ;
    {
        uint64_t a = 0, b = 0, o = 0 ;
        while(a < nfix || b < nforce)
        {
            int take_fix = (b >= nforce) || (a < nfix && fix_ids[a] <= force_ids[b]) ;
            if(take_fix) { mid[o] = fix_ids[a] ; mval[o] = fix_values[a] ; mforce[o] = 0 ; a++ ; }
            else         { mid[o] = force_ids[b] ; mval[o] = force_values[b] ; mforce[o] = 1 ; b++ ; }
            o++ ;
        }
    }
    uint64_t * acc = (uint64_t *)malloc((nb+1)*sizeof(uint64_t)) ;
    acc[0] = 0 ;
    for(uint64_t i = 0 ; i < nb ; i++) acc[i+1] = acc[i]+row_size[i] ;

    for(uint64_t k = 0 ; k < nb ; k++)
    {
        if(!row_size[k]) continue ;
        const int64_t line = (int64_t)k ;
        int64_t start = lb(mid, 0, (int64_t)nm, (int64_t)column_index[acc[k]]*s) ;                         /* :147 */
        int64_t end = ub(mid, start, (int64_t)nm, (int64_t)column_index[acc[k]+row_size[k]-1]*s+s-1) ;      /* :148 */
        int64_t line0 = lb(mid, start, end, line*s) ;                                                       /* :149 */
        int64_t line1 = ub(mid, start, end, line*s+s-1) ;                                                   /* :150 */
        for(uint64_t l = 0 ; l < row_size[k] ; l++)
        {
            const int64_t colb = column_index[acc[k]+l] ;
            int64_t blk0 = lb(mid, start, end, colb*s) ;                                                    /* :159 */
            int64_t blk1 = ub(mid, start, end, colb*s+s-1) ;                                                /* :160 */
            double * blockstart = array+(acc[k]+l)*(uint64_t)s*cl ;
            for(int64_t p = line0 ; p != line1 ; p++)                                                       /* :170-207 */
            {
                if(mforce[p]) continue ;
                const int64_t id = mid[p] ;
                if(add_to_forces) add_to_forces[id] = 0. ;
                for(int m = 0 ; m < s ; m++)
                {
                    if(id != line*s+m)
                    {
                        for(int n = 0 ; n < s ; n++)
                            if(id == colb*s+n)
                            {
                                double val = blockstart[cl*n+m] ;
                                forces[line*s+m] -= mval[p]*val ;
                                if(natural) natural[line*s+m] -= mval[p]*val ;
                                blockstart[cl*n+m] = 0 ;
                            }
                    }
                    else
                        for(int n = 0 ; n < s ; n++)
                            blockstart[cl*n+m] = (colb*s+n == id) ? 1 : 0 ;
                }
            }
            for(int64_t p = blk0 ; p != blk1 ; p++)                                                         /* :210-253 */
            {
                if(mforce[p]) continue ;
                const int64_t id = mid[p] ;
                if(add_to_forces) add_to_forces[id] = 0. ;
                for(int m = 0 ; m < s ; m++)
                {
                    if(id != line*s+m)
                    {
                        for(int n = 0 ; n < s ; n++)
                            if(id == colb*s+n)
                            {
                                double val = blockstart[cl*n+m] ;
                                forces[line*s+m] -= mval[p]*val ;
                                if(natural) natural[line*s+m] -= mval[p]*val ;
                                blockstart[cl*n+m] = 0 ;
                            }
                    }
                    else
                    {
                        forces[id] = mval[p] ;
                        for(int n = 0 ; n < s ; n++)
                            blockstart[cl*n+m] = (colb*s+n == id) ? 1 : 0 ;
                    }
                }
            }
        }
    }
    for(uint64_t i = 0 ; i < nm ; i++)                                                                      /* :258-268 */
        if(mforce[i]) forces[mid[i]] += mval[i] ;
    if(add_to_forces)                                                                                       /* :323-324 */
        for(uint64_t i = 0 ; i < nb*(uint64_t)s ; i++) forces[i] += add_to_forces[i] ;
    free(acc) ; free(mid) ; free(mval) ; free(mforce) ;
}
