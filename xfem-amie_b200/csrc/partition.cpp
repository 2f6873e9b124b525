// partition.cpp -- contiguous block-row partition and halo lists (host only).
// Rank r owns block rows [bounds[r], bounds[r+1]), balanced by stored blocks; its halo is the set
// of distinct off-range block columns its rows reference (what it must receive before a SpMV).
#include "../../include/amie_b200.h"
#include <vector>
#include <algorithm>

extern "C" {

int amie_b200_partition_rows(uint64_t nb, const uint32_t * row_size, int nparts, uint64_t * bounds_out)
{
    if(!row_size || !bounds_out || nparts < 1) return AMIE_B200_ERR_ARG ;
    uint64_t total = 0 ;
    for(uint64_t i = 0 ; i < nb ; i++) total += row_size[i] ;
    bounds_out[0] = 0 ;
    uint64_t acc = 0, row = 0 ;
    for(int p = 1 ; p < nparts ; p++)
    {
        // first row index at which the running block count reaches p/nparts of the total
        const uint64_t target = (total*(uint64_t)p)/(uint64_t)nparts ;
        while(row < nb && acc+row_size[row] <= target)
        {
            acc += row_size[row] ;
            row++ ;
        }
        bounds_out[p] = row ;
    }
    bounds_out[nparts] = nb ;
    for(int p = 1 ; p <= nparts ; p++)
        if(bounds_out[p] < bounds_out[p-1]) bounds_out[p] = bounds_out[p-1] ;
    return AMIE_B200_OK ;
}

int amie_b200_partition_halo(uint64_t r0, uint64_t r1, const uint32_t * row_size_local,
                             const uint32_t * column_index_local, uint32_t * halo_out, uint64_t * nhalo_out)
{
    if(!row_size_local || !nhalo_out || r1 < r0) return AMIE_B200_ERR_ARG ;
    uint64_t nnz = 0 ;
    for(uint64_t i = 0 ; i < r1-r0 ; i++) nnz += row_size_local[i] ;
    std::vector<uint32_t> h ;
    for(uint64_t k = 0 ; k < nnz ; k++)
    {
        const uint32_t c = column_index_local[k] ;
        if(c < r0 || c >= r1) h.push_back(c) ;
    }
    std::sort(h.begin(), h.end()) ;
    h.erase(std::unique(h.begin(), h.end()), h.end()) ;
    *nhalo_out = h.size() ;
    if(halo_out) std::copy(h.begin(), h.end(), halo_out) ;
    return AMIE_B200_OK ;
}

}
