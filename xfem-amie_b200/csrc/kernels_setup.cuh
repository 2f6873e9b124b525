// kernels_setup.cuh -- per-matrix set-up kernels: the diagonals of the reference's diagonal preconditioners and
// K-Repack.  Depends on device_utils.cuh only (see there).
#pragma once
#include "device_utils.cuh"

// CoordinateIndexedSparseMatrix::inverseDiagonal (sparse/sparse_matrix.cpp:216-231):
// d_i = 1/A_ii if |A_ii| > 1e-12 else 0 ; a missing diagonal block reads as 0.
template<int S>
__global__ void k_inverse_diagonal(const uint32_t * rowptr, const uint32_t * col, const double * vals,
                                   uint32_t row_base, uint64_t nrows, double * d)
{
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < nrows*S ; i += (uint64_t)gridDim.x*blockDim.x)
    {
        const uint32_t row = (uint32_t)(i/S) ;
        const int m = (int)(i-(uint64_t)row*S) ;
        uint32_t k0 = rowptr[row], k1 = rowptr[row+1] ;
        const uint32_t key = row_base+row ;
        while(k0 < k1)
        {
            const uint32_t mid = k0+((k1-k0) >> 1) ;
            if(col[mid] < key) k0 = mid+1 ; else k1 = mid ;
        }
        double v = 0. ;
        if(k0 < rowptr[row+1] && col[k0] == key) v = vals[(size_t)k0*S*S+m*S+m] ;
        d[i] = fabs(v) > 1e-12 ? 1./v : 0. ;
    }
}

// The reference's other diagonal preconditioners (solvers/inversediagonal.cpp), same storage walk:
// InverseDiagonalSquared: 1/(A_ii*A_ii), no threshold (CoordinateIndexedSparseMatrix::inverseDiagonalSquared,
// sparse/sparse_matrix.cpp:233-244).
template<int S>
__global__ void k_inverse_diagonal_squared(const uint32_t * rowptr, const uint32_t * col, const double * vals,
                                           uint64_t nrows, double * d)
{
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < nrows*S ; i += (uint64_t)gridDim.x*blockDim.x)
    {
        const uint32_t row = (uint32_t)(i/S) ;
        const int m = (int)(i-(uint64_t)row*S) ;
        uint32_t k0 = rowptr[row], k1 = rowptr[row+1] ;
        while(k0 < k1)
        {
            const uint32_t mid = k0+((k1-k0) >> 1) ;
            if(col[mid] < row) k0 = mid+1 ; else k1 = mid ;
        }
        double v = 0. ;
        if(k0 < rowptr[row+1] && col[k0] == row) v = vals[(size_t)k0*S*S+m*S+m] ;
        d[i] = 1./__dmul_rn(v, v) ;
    }
}

// InverseLumpedDiagonal (solvers/inversediagonal.cpp:19-42): the sum of scalar row i over all columns ascending
// (= the stored entries of the row in storage order; absent entries add 0), then 1/v if |v| > 1e-8, else +-1.
template<int S>
__global__ void k_inverse_lumped_diagonal(const uint32_t * rowptr, const uint32_t * col, const double * vals,
                                          uint64_t nrows, double * d)
{
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < nrows*S ; i += (uint64_t)gridDim.x*blockDim.x)
    {
        const uint32_t row = (uint32_t)(i/S) ;
        const int m = (int)(i-(uint64_t)row*S) ;
        const uint32_t k0 = rowptr[row], k1 = rowptr[row+1] ;
        double v = 0. ;
        for(uint32_t k = k0 ; k < k1 ; k++)
            #pragma unroll
            for(int c = 0 ; c < S ; c++)
                v = __dadd_rn(v, vals[(size_t)k*S*S+c*S+m]) ;
        d[i] = fabs(v) > 1e-8 ? 1./v : (v > 0. ? 1. : -1.) ;
    }
}

// K-Repack: reference padded column-major blocks (cl = S + S%2) -> compact S*S blocks
template<int S>
__global__ void k_repack(const double * padded, double * compact, uint64_t nblocks)
{
    constexpr int CL = S+S%2 ;
    constexpr int SS = S*S ;
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < nblocks*SS ; i += (uint64_t)gridDim.x*blockDim.x)
    {
        const uint64_t k = i/SS ;
        const int e = (int)(i-k*SS) ;
        const int c = e/S, r = e-c*S ;
        compact[i] = padded[k*(S*CL)+c*CL+r] ;
    }
}

// K-Repack through a block map (amie_b200_set_block_map): block k of the chunk goes to stored block block_to[k] --
// the host array is in another block order than the structure on the device (a renumbered matrix, csrc/reorder.cpp)
template<int S>
__global__ void k_repack_scatter(const double * padded, double * compact, const uint32_t * block_to, uint64_t nblocks)
{
    constexpr int CL = S+S%2 ;
    constexpr int SS = S*S ;
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < nblocks*SS ; i += (uint64_t)gridDim.x*blockDim.x)
    {
        const uint64_t k = i/SS ;
        const int e = (int)(i-k*SS) ;
        const int c = e/S, r = e-c*S ;
        compact[(uint64_t)block_to[k]*SS+e] = padded[k*(S*CL)+c*CL+r] ;
    }
}

static __global__ void k_rowptr_from_sizes_check(const uint32_t * col, const uint32_t * rowptr, uint64_t nb, uint32_t ncols, int * bad)
{
    for(uint64_t r = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; r < nb ; r += (uint64_t)gridDim.x*blockDim.x)
    {
        for(uint32_t k = rowptr[r] ; k < rowptr[r+1] ; k++)
        {
            if(col[k] >= ncols) *bad = 1 ;
            if(k > rowptr[r] && col[k] <= col[k-1]) *bad = 2 ;
        }
    }
}
