// reorder.cpp -- node renumbering of a block-sparse structure (host only).
//
// The numbering AMIE's mesher hands to Assembly has no locality (profiles/r01_notes.md: mean |column - row| is a
// quarter of the matrix on a FeatureTree-assembled system), which costs the SpMV its x-gather reuse window and makes
// neighbouring rows very unequal.  These two functions are the host half of a renumbering applied once per topology:
//   amie_b200_rcm_order         reverse Cuthill-McKee on the block graph (every connected component, started from a
//                               node of minimum degree, neighbours by ascending degree);
//   amie_b200_group_rows_by_length  (opt-in) rows of similar length next to one another inside windows of that numbering;
//   amie_b200_permute_structure the structure in the new numbering (columns ascending inside each row again, as
//                               CoordinateIndexedSparseMatrix requires -- sparse/sparse_vector.h:864-870 binary-searches
//                               them) plus, for every stored block, where it came from, so that values can follow by
//                               a gather.
// A solve is invariant under such a renumbering up to the rounding of its dot products.
#include "../../include/amie_b200.h"
#include <vector>
#include <algorithm>
#include <numeric>

extern "C" {

int amie_b200_rcm_order(uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, uint32_t * perm_out)
{
    if((nb && (!row_size || !perm_out)) || nb >= 0xffffffffull) return AMIE_B200_ERR_ARG ;
    std::vector<uint64_t> acc(nb+1, 0) ;
    for(uint64_t i = 0 ; i < nb ; i++) acc[i+1] = acc[i]+row_size[i] ;
    if(acc[nb] && !column_index) return AMIE_B200_ERR_ARG ;
    for(uint64_t k = 0 ; k < acc[nb] ; k++) if(column_index[k] >= nb) return AMIE_B200_ERR_ARG ;
    auto degree = [&](uint32_t i) { return row_size[i] ; } ;
    std::vector<uint32_t> order ;                    // Cuthill-McKee visiting order (old ids)
    order.reserve(nb) ;
    std::vector<char> seen(nb, 0) ;
    // start nodes: ascending degree, ties by id, so the result does not depend on anything but the structure
    std::vector<uint32_t> starts(nb) ;
    std::iota(starts.begin(), starts.end(), 0u) ;
    std::stable_sort(starts.begin(), starts.end(), [&](uint32_t a, uint32_t b) { return degree(a) < degree(b) ; }) ;
    std::vector<uint32_t> nbrs ;
    for(uint32_t s : starts)
    {
        if(seen[s]) continue ;
        seen[s] = 1 ;
        size_t head = order.size() ;
        order.push_back(s) ;
        while(head < order.size())
        {
            const uint32_t v = order[head++] ;
            nbrs.clear() ;
            for(uint64_t k = acc[v] ; k < acc[v+1] ; k++)
            {
                const uint32_t c = column_index[k] ;
                if(!seen[c]) { seen[c] = 1 ; nbrs.push_back(c) ; }
            }
            std::stable_sort(nbrs.begin(), nbrs.end(), [&](uint32_t a, uint32_t b) { return degree(a) < degree(b) ; }) ;
            order.insert(order.end(), nbrs.begin(), nbrs.end()) ;
        }
    }
    for(uint64_t i = 0 ; i < nb ; i++) perm_out[order[nb-1-i]] = (uint32_t)i ;       // reversed
    return AMIE_B200_OK ;
}

// Rows of similar length next to one another, without giving the locality of `perm` away: inside every window of
// `window` consecutive nodes of the numbering `perm`, the nodes are re-ordered by row length (longest first, ties in
// the order they had).  The row-thread SpMV walks a tile of 10 (16 in 2D) consecutive block rows with one lane per
// scalar row, so a tile costs its LONGEST row; on a FeatureTree-assembled 3D system (26 088 unknowns) the sum of
// 10 x max length over the tiles is 1.86x the stored blocks as numbered by the mesher, 1.71x after Cuthill-McKee,
// 1.16x with window 80 on top of it -- paid for by the x lines a tile touches (27.6 -> 44.5; profiles/r02_notes.md
// section 11).  Opt-in until that trade is measured on the device.
int amie_b200_group_rows_by_length(uint64_t nb, const uint32_t * row_size, uint64_t window, uint32_t * perm_inout)
{
    if((nb && (!row_size || !perm_inout)) || nb >= 0xffffffffull || window == 0) return AMIE_B200_ERR_ARG ;
    std::vector<uint32_t> inv(nb, 0xffffffffu) ;            // inv[new] = old
    for(uint64_t i = 0 ; i < nb ; i++)
    {
        if(perm_inout[i] >= nb || inv[perm_inout[i]] != 0xffffffffu) return AMIE_B200_ERR_ARG ;
        inv[perm_inout[i]] = (uint32_t)i ;
    }
    for(uint64_t w0 = 0 ; w0 < nb ; w0 += window)
    {
        const uint64_t w1 = std::min<uint64_t>(nb, w0+window) ;
        std::stable_sort(inv.begin()+w0, inv.begin()+w1, [&](uint32_t a, uint32_t b) { return row_size[a] > row_size[b] ; }) ;
    }
    for(uint64_t i = 0 ; i < nb ; i++) perm_inout[inv[i]] = (uint32_t)i ;
    return AMIE_B200_OK ;
}

int amie_b200_permute_structure(uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, const uint32_t * perm,
                                uint32_t * row_size_out, uint32_t * column_index_out, uint32_t * block_from_out)
{
    if(nb && (!row_size || !perm || !row_size_out)) return AMIE_B200_ERR_ARG ;
    std::vector<uint64_t> acc(nb+1, 0) ;
    for(uint64_t i = 0 ; i < nb ; i++) acc[i+1] = acc[i]+row_size[i] ;
    if(acc[nb] >= 0xffffffffull || (acc[nb] && (!column_index || !column_index_out || !block_from_out))) return AMIE_B200_ERR_ARG ;
    // perm must be a permutation of 0 .. nb-1
    std::vector<uint32_t> inv(nb, 0xffffffffu) ;
    for(uint64_t i = 0 ; i < nb ; i++)
    {
        if(perm[i] >= nb || inv[perm[i]] != 0xffffffffu) return AMIE_B200_ERR_ARG ;
        inv[perm[i]] = (uint32_t)i ;
    }
    for(uint64_t k = 0 ; k < acc[nb] ; k++) if(column_index[k] >= nb) return AMIE_B200_ERR_ARG ;
    // offsets of the new rows, then every row on its own (rows are independent: OpenMP over them)
    std::vector<uint64_t> accn(nb+1, 0) ;
    for(uint64_t r = 0 ; r < nb ; r++)
    {
        row_size_out[r] = row_size[inv[r]] ;
        accn[r+1] = accn[r]+row_size_out[r] ;
    }
    #pragma omp parallel
    {
        std::vector<std::pair<uint32_t, uint32_t> > row ;      // (new column, old block)
        #pragma omp for schedule(dynamic, 4096)
        for(int64_t r = 0 ; r < (int64_t)nb ; r++)
        {
            const uint32_t old = inv[r] ;
            row.clear() ;
            for(uint64_t k = acc[old] ; k < acc[old+1] ; k++)
                row.push_back(std::make_pair(perm[column_index[k]], (uint32_t)k)) ;
            std::sort(row.begin(), row.end()) ;
            uint64_t pos = accn[r] ;
            for(const auto & e : row)
            {
                column_index_out[pos] = e.first ;
                block_from_out[pos] = e.second ;
                pos++ ;
            }
        }
    }
    return AMIE_B200_OK ;
}

}
