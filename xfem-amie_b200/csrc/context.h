// context.h -- the device context behind amie_b200_ctx (one per Assembly).
#pragma once
#include "../../include/amie_b200.h"
#include "common.cuh"
#include <vector>
#include <string>
#include <chrono>

struct DistState ;   // dist.cu
struct AssemblyMap ; // assemble.cu
struct FieldMap ;    // fields.cu
struct LocalGroup ;  // group.h

struct amie_b200_ctx
{
    int device = 0 ;
    int num_sms = 148 ;
    cudaStream_t stream = nullptr ;
    std::string err ;

    // ---- matrix (device)
    int S = 0 ;
    uint64_t nb = 0 ;          // local block rows
    uint64_t nb_global = 0 ;   // = nb on a single device
    uint64_t row_base = 0 ;    // global index of local block row 0 (distributed)
    uint64_t nnzb = 0 ;
    uint64_t N = 0 ;           // local DOF
    uint64_t ncols_local = 0 ; // block columns addressable by local col indices (nb + halo)
    uint32_t * halo_glob = nullptr ; // row-partitioned context: the sorted GLOBAL block columns of the halo (ncols_local - nb entries, device)
    uint32_t * rowptr = nullptr ;
    uint32_t * col = nullptr ;
    double * vals = nullptr ;
    double * dinv = nullptr ;
    bool have_structure = false, have_values = false, dinv_valid = false ;
    int dinv_kind = AMIE_B200_PRECOND_JACOBI ;   // which preconditioner `dinv` holds while dinv_valid (a diagonal, or s x s blocks for kinds 5 / 6)
    uint64_t dinv_len = 0 ;                      // doubles allocated for it
    double * user_diag = nullptr ;               // AMIE_B200_PRECOND_DIAGONAL: the caller's diagonal (N doubles)
    uint32_t * block_to = nullptr ;              // amie_b200_set_block_map: block k of the host array -> stored block (nnzb)
    bool have_rhs = false ;

    // ---- vectors (device); x-like vectors that are SpMV inputs have room for the halo tail
    uint64_t vec_len = 0 ;     // allocated doubles per vector
    double * b = nullptr, * x = nullptr, * r = nullptr, * z = nullptr, * p = nullptr, * q = nullptr ;
    double * xc = nullptr, * rc = nullptr, * xmin = nullptr ;
    double * w[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr} ;   // BiCGStab work (lazy)

    // Assembly::displacementHistory in HBM (cgsolve.cu): [0] older, [1] newest; count is 0 or 2 like the reference's
    double * hist[2] = {nullptr, nullptr} ;
    int hist_count = 0 ;
    uint64_t hist_n = 0 ;

    KrylovState * st = nullptr ;         // device
    KrylovState * st_host = nullptr ;    // pinned, 4 slots
    double * partials = nullptr ;        // device, 4*AMIE_MAX_PARTIALS
    double * partials_host = nullptr ;   // pinned
    int * flag = nullptr ;               // device scratch

    // ---- options
    int opt_time_spmv = 0 ;
    int opt_variant = 0 ;
    int opt_verbose = 0 ;
    int opt_batch = 0 ;         // iterations per speculative batch (0 = auto)
    int opt_graph = -1 ;        // -1 auto, 0 off, 1 on
    int opt_fields_variant = 1 ;    // 1: slot loop of k_element_fields unrolled and phase-split for linear triangles / tetrahedra; 0: generic loop

    // ---- stats
    amie_b200_stats stats {} ;
    std::vector<cudaEvent_t> ev_pool ;
    size_t ev_used = 0 ;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_poll[2] = {nullptr, nullptr} ;

    DistState * dist = nullptr ;
    LocalGroup * group = nullptr ;     // != nullptr: this is the caller-facing context of several devices (group.cu); it holds no device memory itself
    AssemblyMap * amap = nullptr ;     // element -> stored-block gather lists (device-side value assembly)
    FieldMap * fmap = nullptr ;        // element kinematics + behaviours (field recovery after the solve)

    // ---- CUDA graphs of iteration batches (small systems: launch-bound inner loops)
    struct GraphSlot
    {
        cudaGraphExec_t exec = nullptr ;
        int batch = 0, precond = 0, variant = 0 ;
        uint64_t rowstart = 0, colstart = 0, alloc_gen = 0 ;
    } ;
    GraphSlot graph_cg, graph_bicg ;
    uint64_t alloc_gen = 0 ;           // bumped whenever device arrays are (re)allocated

    void set_error(const std::string & e) { err = e ; }
} ;

inline double wall_now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() ;
}

// internal entry points shared between translation units
void assembly_map_destroy(amie_b200_ctx * ctx) ;          // assemble.cu
void field_map_destroy(amie_b200_ctx * ctx) ;             // fields.cu
uint64_t assembly_map_bytes(const amie_b200_ctx * ctx) ;  // HBM held by the gather lists, element matrices and masks
uint64_t field_map_bytes(const amie_b200_ctx * ctx) ;     // ... by the element kinematics, behaviours and results
void history_destroy(amie_b200_ctx * ctx) ;               // cgsolve.cu
void ctx_free_matrix(amie_b200_ctx * ctx) ;               // api.cu: matrix arrays + everything tied to the topology
int ctx_alloc_vectors(amie_b200_ctx * ctx) ;
int ctx_ensure_bicg_vectors(amie_b200_ctx * ctx) ;
int ctx_ensure_dinv(amie_b200_ctx * ctx, int kind = AMIE_B200_PRECOND_JACOBI) ;   // dinv = the diagonal of preconditioner `kind`
int ctx_sync_state(amie_b200_ctx * ctx, int slot) ;       // D2H of KrylovState into st_host[slot] + sync
int ctx_push_state(amie_b200_ctx * ctx, const KrylovState & s) ;
void ctx_reset_solve_stats(amie_b200_ctx * ctx) ;
void ctx_collect_spmv_times(amie_b200_ctx * ctx) ;
int ctx_max(amie_b200_ctx * ctx, const double * v, uint64_t n, int mode, double * out) ;
int solve_cg_resident(amie_b200_ctx * ctx, int precond_kind, double eps, int maxit, uint64_t nssor,
                      uint64_t rowstart, uint64_t colstart, uint64_t * nit_out, double * err_out, double * rho_out) ;
int solve_bicg_resident(amie_b200_ctx * ctx, int precond_kind, double eps, int maxit,
                        uint64_t * nit_out, double * err_out) ;

// persistent grids: a multiple of the SM count, never more blocks than there is work, and
// never more than the fused reductions can hold.
static inline int vec_grid(const amie_b200_ctx * ctx, uint64_t n)
{
    uint64_t want = (n+AMIE_VEC_THREADS-1)/AMIE_VEC_THREADS ;
    uint64_t cap = (uint64_t)ctx->num_sms*8 ;
    if(cap > AMIE_MAX_PARTIALS) cap = AMIE_MAX_PARTIALS ;
    uint64_t g = want < cap ? want : cap ;
    return (int)(g ? g : 1) ;
}
