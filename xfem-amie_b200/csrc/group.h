// group.h -- several GPUs behind ONE context, driven from ONE caller thread (amie_b200_create(devices, ndev > 1)).
//
// The reference calls the solver from a single thread of a single process (Assembly::cgsolve,
// solvers/assembly.cpp:1841-1850).  A group context keeps that shape: the caller hands GLOBAL host arrays to the
// usual entry points; the context partitions the block rows (amie_b200_partition_rows), owns one child context per
// device and one worker thread per child.  Every child runs exactly the per-rank code of dist.cu -- same kernels,
// same NVLink peer-memory halo pushes and mailbox reductions -- but the set-up collectives (halo lists, pointer
// exchange, agreement on scalars) go through this struct instead of NCCL / cudaIpc: all ranks share an address space
// and cudaDeviceEnablePeerAccess makes every child's memory addressable from every device.
#pragma once
#include "context.h"
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define GROUP_MAX 8

struct LocalGroup
{
    int world = 0 ;
    std::vector<amie_b200_ctx *> child ;
    std::vector<uint64_t> bounds ;            // block-row partition, world+1
    std::vector<uint64_t> blk_off ;           // stored blocks in front of each part, world+1
    bool broken = false ;                     // a collective was abandoned: the children are out of step for good
    std::vector<uint64_t> elem_bounds ;       // field recovery: elements [elem_bounds[r], elem_bounds[r+1]) live on device r
    int field_nc = 0, field_dim = 0 ;
    bool field_shared_behaviours = false ;    // set_element_behaviour came with tensor_of_elem (one table on every device)
    std::vector<uint32_t> block_from ;        // set_block_map: stored block j holds block block_from[j] of the caller's array (empty: identity)

    // ---- exchange slots: rank r writes [r], barrier, everybody reads, barrier
    void * ptr[GROUP_MAX] = {} ;
    double dbl[GROUP_MAX] = {} ;
    const void * cptr[GROUP_MAX] = {} ;
    std::vector<long long> ll[GROUP_MAX] ;

    // ---- barrier between the worker threads; abort() releases everybody with `false`
    bool barrier() ;
    void abort() ;

    // ---- run f(rank) on every worker; returns the first non-zero result in rank order (negative codes first)
    int run(const std::function<int(int)> & f) ;
    int results[GROUP_MAX] = {} ;

    void start(int world) ;
    void stop() ;

private:
    std::mutex bm ;
    std::condition_variable bcv ;
    int waiting = 0 ;
    uint64_t generation = 0 ;
    bool aborted = false ;

    std::mutex jm ;
    std::condition_variable jcv, dcv ;
    const std::function<int(int)> * job = nullptr ;
    uint64_t job_gen = 0 ;
    int job_left = 0 ;
    bool quit = false ;
    std::vector<std::thread> workers ;
    void worker(int rank) ;
} ;

// group.cu: the C-ABI entry points on a group context (ctx->group != nullptr)
amie_b200_ctx * group_create(const int * devices, int ndev, std::string & err) ;
void group_destroy(amie_b200_ctx * ctx) ;
int group_set_structure(amie_b200_ctx * ctx, int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, uint64_t nnzb) ;
int group_set_values(amie_b200_ctx * ctx, const double * array) ;
int group_synth_to_device(amie_b200_ctx * ctx, const struct amie_b200_synth * s) ;
int group_solve(amie_b200_ctx * ctx, bool bicg, bool resident, const double * b, const double * x0, uint64_t nx0,
                int precond_kind, double eps, int maxit, uint64_t nssor, uint64_t rowstart, uint64_t colstart,
                double * x_out, uint64_t * nit_out, double * err_out, double * rho_out) ;
// a per-child call on the slices of up to three host vectors of length N (nullptr = not passed)
int group_sliced(amie_b200_ctx * ctx, const std::function<int(amie_b200_ctx *, uint64_t dof0, uint64_t ndof)> & f) ;
int group_get_stats(const amie_b200_ctx * ctx, amie_b200_stats * out) ;
int group_set_option(amie_b200_ctx * ctx, const char * key, int64_t value) ;
int group_unsupported(amie_b200_ctx * ctx, const char * what) ;
int group_set_block_map(amie_b200_ctx * ctx, const uint32_t * block_to) ;
// api.cu: the values of ctx's stored blocks from blocks src[j] of a host array (host gather + K-Repack)
int ctx_set_values_from(amie_b200_ctx * ctx, const double * array, const uint32_t * src) ;
extern "C" void amie_b200_gather_blocks(const double * array, const uint32_t * src, uint64_t nblk, uint64_t per_block, double * out) ;
int group_download_matrix(amie_b200_ctx * ctx, uint32_t * row_size_out, uint32_t * column_index_out, double * array_padded_out) ;
// the rows next to the solve.  Value assembly: every device sees the whole element list (global node ids) and keeps the
// contributions to the block rows it owns.  Field recovery: the ELEMENTS are split, every device holds the whole field.
int group_set_elements(amie_b200_ctx * ctx, uint64_t n_elem, int npe, const uint32_t * elem_ids) ;
int group_update_elements(amie_b200_ctx * ctx, uint64_t first, uint64_t count, const double * ke, const double * scales) ;
int group_assemble(amie_b200_ctx * ctx) ;
int group_set_boundary_conditions(amie_b200_ctx * ctx, uint64_t nfix, const uint32_t * fix_ids, const double * fix_values,
                                  uint64_t nforce, const uint32_t * force_ids, const double * force_values,
                                  const double * add_to_forces, double * natural_inout) ;
int group_set_element_kinematics(amie_b200_ctx * ctx, uint64_t n_elem, int npe, int dim, const uint32_t * elem_ids,
                                 const double * dshape, const double * jinv) ;
int group_set_element_behaviour(amie_b200_ctx * ctx, uint64_t n_tensors, const double * tensors, const double * imposed_strain,
                                const double * imposed_stress, const uint32_t * tensor_of_elem) ;
int group_element_fields(amie_b200_ctx * ctx, const double * u, uint64_t n_u, double * total_strain_out,
                         double * mechanical_strain_out, double * real_stress_out) ;
int group_element_principal(amie_b200_ctx * ctx, int field, double * principal_out) ;
// fields.cu
int fields_u_buffer(amie_b200_ctx * ctx, uint64_t n, double ** out) ;
int fields_run(amie_b200_ctx * ctx, const double * du, uint64_t len, uint64_t h2d,
               double * total_strain_out, double * mechanical_strain_out, double * real_stress_out) ;

// dist.cu: make `ctx` rank `rank` of `g` (the in-process counterpart of amie_b200_dist_init)
int dist_init_local(amie_b200_ctx * ctx, int rank, LocalGroup * g) ;
int dist_set_structure_local(amie_b200_ctx * ctx, int stride, uint64_t nb_global, const uint32_t * row_size_local,
                             const uint32_t * column_index_local, uint64_t nnzb_local) ;
