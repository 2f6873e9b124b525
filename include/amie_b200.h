/* amie_b200.h -- C-ABI of the B200-native block-sparse Krylov solve for AMIE.
 *
 * This is the drop-in boundary for ONE path of the reference (CyrilleDunant/xfem-amie):
 *   Assembly::cgsolve -> ConjugateGradient::solve / BiConjugateGradientStabilized::solve
 *   (solvers/assembly.cpp:1829-1953, solvers/conjugategradient.cpp:69-318,
 *    solvers/biconjugategradientstabilized.cpp:12-148)
 * over Assembly's CoordinateIndexedSparseMatrix (sparse/sparse_matrix.h:129-136) with the
 * InverseDiagonal preconditioner (solvers/inversediagonal.cpp:48-67).
 *
 * All pointers are HOST memory owned by the caller unless a function says "device".  The
 * context owns the device memory.  One context per Assembly; calls are serialised by the
 * caller (the reference solver is single-caller, not re-entrant).  There is no CPU fallback:
 * every compute entry point fails with AMIE_B200_ERR_CUDA when no sm_100 device is usable.
 *
 * Return convention: solver entry points return 1 (converged) / 0 (not converged) exactly like
 * the reference's `bool solve(...)`, or a negative AMIE_B200_ERR_* code.  Other entry points
 * return 0 on success or a negative code.
 */
#ifndef AMIE_B200_H
#define AMIE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMIE_B200_OK              0
#define AMIE_B200_ERR_CUDA       -1   /* CUDA runtime / no device / kernel failure            */
#define AMIE_B200_ERR_ARG        -2   /* bad argument                                         */
#define AMIE_B200_ERR_STATE      -3   /* call order (e.g. pcg before set_values)              */
#define AMIE_B200_ERR_NAN        -4   /* NaN initial residual: the reference prints the       */
                                      /* assembly and exit(0)s (conjugategradient.cpp:160-165) */
#define AMIE_B200_ERR_UNSUPPORTED -5  /* stride / preconditioner kind not on the device path   */
#define AMIE_B200_ERR_NCCL       -6

/* precond_kind: what the caller passed as `Preconditionner * precond` */
#define AMIE_B200_PRECOND_JACOBI 0    /* nullptr -> InverseDiagonal (conjugategradient.cpp:80-84)           */
#define AMIE_B200_PRECOND_NULL   1    /* NullPreconditionner: precondition() is a no-op (preconditionners.cpp) */
/* the reference's other diagonal preconditioners (solvers/inversediagonal.cpp; precondition() is t = v .* d):      */
#define AMIE_B200_PRECOND_DIAGONAL_SQUARED 2  /* InverseDiagonalSquared: d = 1/(A_ii*A_ii)      (:69-82)             */
#define AMIE_B200_PRECOND_LUMPED 3    /* InverseLumpedDiagonal: d = 1/(row sum), +-1 if tiny    (:19-48)             */
#define AMIE_B200_PRECOND_DIAGONAL 4  /* any Preconditionner of that form: the caller supplies d with                */
                                      /* amie_b200_set_preconditioner_diagonal                                       */
/* block preconditioners, PCG only (built on the device from the resident values):     */
#define AMIE_B200_PRECOND_BLOCK2X2 5  /* Inverse2x2Diagonal (solvers/inversediagonal.cpp:84-133) on a stride-2 matrix: */
                                      /* the inverse of every node's 2x2 diagonal block; same bits as the class      */
#define AMIE_B200_PRECOND_BLOCK3X3 6  /* the same construction on the 3x3 node blocks of a stride-3 matrix, with the */
                                      /* reference's det() / invert3x3Matrix (utilities/matrixops.cpp:838-847,       */
                                      /* :681-702).  No reference class: opt-in block-Jacobi, other iteration counts */

typedef struct amie_b200_ctx amie_b200_ctx ;

/* ------------------------------------------------------------------ context */

/* devices: CUDA ordinals.
 *   ndev == 1 (or NULL/0 -> device 0, env AMIE_B200_DEVICE): one context drives one device.
 *   ndev  > 1 (or NULL/0 with env AMIE_B200_DEVICES="0,1,..."): ONE context over several GPUs of a box, still driven
 *     by one caller thread with GLOBAL host arrays -- the shape Assembly::cgsolve has (solvers/assembly.cpp:1841-1850).
 *     The block rows are partitioned inside the library (amie_b200_partition_rows), one worker thread and one child
 *     context per device, halo and reductions over NVLink peer memory (csrc/group.cu, csrc/dist.cu).  Needs a
 *     peer-to-peer path between the devices and CUDA_MODULE_LOADING=EAGER (set automatically when AMIE_B200_DEVICES
 *     lists several devices at load time).  Strides 2 and 3.  An ordinal may be listed more than once (several parts on
 *     one GPU: how the path is tested on a one-GPU box).  The value-assembly rows (set_elements / update_elements /
 *     assemble / set_boundary_conditions) send the element list and the id lists to every device, which keeps what lands
 *     on the block rows it owns; the field-recovery rows split the ELEMENTS over the devices, each holding the whole
 *     displacement field (from the host, or gathered from the parts over the peer links).  Element and dof ids stay the
 *     caller's global ones throughout.  set_block_map works too (each device's values are gathered on the host from
 *     the caller's array through the inverse map).
 *   One process per GPU (torchrun, MPI) uses amie_b200_dist_init (below) instead.
 * Returns NULL on failure (see amie_b200_global_error). */
amie_b200_ctx * amie_b200_create(const int * devices, int ndev) ;
void            amie_b200_destroy(amie_b200_ctx * ctx) ;
const char *    amie_b200_last_error(const amie_b200_ctx * ctx) ;
const char *    amie_b200_global_error(void) ;
const char *    amie_b200_version(void) ;

/* ------------------------------------------------------------------ matrix
 * Replaces the reads the reference solver does through assembly->getMatrix()
 * (CoordinateIndexedSparseMatrix: stride, row_size, column_index, array;
 *  sparse/sparse_matrix.h:129-136, ctor sparse/sparse_matrix.cpp:59-66).           */

/* Once per topology change (Assembly::clear(), solvers/assembly.cpp:1380-1391).
 * row_size[nb], column_index[nnzb] sorted ascending inside each block row.           */
int amie_b200_set_structure(amie_b200_ctx * ctx, int stride, uint64_t nb,
                            const uint32_t * row_size, const uint32_t * column_index, uint64_t nnzb) ;

/* Every solve whose values changed.  `array` is the reference layout: block k at
 * array[k*s*cl], element (r,c) at + c*cl + r, cl = s + s%2 (pad slots ignored).      */
int amie_b200_set_values(amie_b200_ctx * ctx, const double * array_padded_colmajor) ;

/* The `array` given to set_values is in ANOTHER block order than the structure given to set_structure: block k of the
 * array is stored block block_to[k] (a permutation of 0 .. nnzb-1; NULL removes the map).  This is how a host keeps its
 * own numbering -- AMIE's mesher numbering has no locality -- while the device works on a renumbered matrix
 * (amie_b200_rcm_order / amie_b200_permute_structure below: block_to is the inverse of block_from); the host then
 * permutes b, x0 and x at the boundary (host/shim does, env AMIE_B200_RENUMBER=1).                                 */
int amie_b200_set_block_map(amie_b200_ctx * ctx, const uint32_t * block_to) ;

/* ------------------------------------------------------------------ solvers */

/* ConjugateGradient::solve(x0, precond, eps, maxit) with members nssor/rowstart/colstart
 * (solvers/conjugategradient.h:22-43).  b = assembly->getForces().
 * x0 may be NULL / shorter than N (conjugategradient.cpp:95-104).
 * x_out[N]; nit_out = cg.nit; err_out = the "Error :" value of the final cerr line;
 * rho_out = "last rho".  Returns 1/0 or <0.                                              */
int amie_b200_pcg(amie_b200_ctx * ctx, const double * b, const double * x0, uint64_t nx0,
                  int precond_kind, double eps, int maxit, uint64_t nssor,
                  uint64_t rowstart, uint64_t colstart,
                  double * x_out, uint64_t * nit_out, double * err_out, double * rho_out) ;

/* BiConjugateGradientStabilized::solve (solvers/biconjugategradientstabilized.cpp:12-148).
 * Ignores rowstart/colstart like the reference.                                          */
int amie_b200_bicgstab(amie_b200_ctx * ctx, const double * b, const double * x0, uint64_t nx0,
                       int precond_kind, double eps, int maxit,
                       double * x_out, uint64_t * nit_out, double * err_out) ;

/* assign(y, A*x [- b], rowstart, colstart)  (sparse/sparse_matrix.cpp:462-547):
 * rows < rowstart are 0, block columns < colstart/stride are skipped.  b may be NULL.   */
int amie_b200_spmv(amie_b200_ctx * ctx, const double * x, const double * b,
                   uint64_t rowstart, uint64_t colstart, double * y_out) ;

/* r = K u - f and its Euclidean norm: what FeatureTree::solve computes right after cgsolve
 * (features/features.cpp:4766-4768, there through the serial operator path).  r_out may be NULL.  */
int amie_b200_residual(amie_b200_ctx * ctx, const double * u, const double * f, double * r_out, double * norm_out) ;

/* CoordinateIndexedSparseMatrix::inverseDiagonal (sparse/sparse_matrix.cpp:216-231)      */
int amie_b200_inverse_diagonal(amie_b200_ctx * ctx, double * d_out) ;
/* The diagonal of preconditioner `precond_kind` (0, 2, 3: built on the device from the resident values; 4: the one
 * set below), as the corresponding reference class holds it in its `diagonal` member.                              */
int amie_b200_preconditioner_diagonal(amie_b200_ctx * ctx, int precond_kind, double * d_out) ;
/* The s x s blocks of precond_kind 5 | 6, row-major per node (blocks_out[nb*s*s]): Inverse2x2Diagonal::blocks.     */
int amie_b200_preconditioner_blocks(amie_b200_ctx * ctx, int precond_kind, double * blocks_out) ;
/* d[N] for AMIE_B200_PRECOND_DIAGONAL: what a user-written Preconditionner with precondition(v, t) { t = v*d } holds.
 * Kept until replaced or until the structure changes.                                                              */
int amie_b200_set_preconditioner_diagonal(amie_b200_ctx * ctx, const double * d) ;

/* ------------------------------------------------------------------ device-resident variants
 * Same algorithms with b / x0 / x kept in HBM (no host<->device copies in the call): used to
 * separate kernel throughput from PCIe time, and by callers that keep x on the device
 * between the CG, CG, BiCGStab triple of one FeatureTree::step.   */
int amie_b200_upload_rhs(amie_b200_ctx * ctx, const double * b) ;                 /* b  -> device */
int amie_b200_upload_x0(amie_b200_ctx * ctx, const double * x0, uint64_t nx0) ;   /* x0 -> device (zero-filled) */
int amie_b200_download_x(amie_b200_ctx * ctx, double * x_out) ;
int amie_b200_download_rhs(amie_b200_ctx * ctx, double * b_out) ;
/* work vectors, for tests: which = 0 x, 1 q (result of amie_b200_spmv_resident), 2 r, 3 p */
int amie_b200_download_vector(amie_b200_ctx * ctx, int which, double * out) ;
/* the matrix as held on the device, converted back to the reference layout (tests; any pointer may be NULL) */
int amie_b200_download_matrix(amie_b200_ctx * ctx, uint32_t * row_size_out, uint32_t * column_index_out,
                              double * array_padded_out) ;
int amie_b200_pcg_resident(amie_b200_ctx * ctx, int precond_kind, double eps, int maxit, uint64_t nssor,
                           uint64_t rowstart, uint64_t colstart,
                           uint64_t * nit_out, double * err_out, double * rho_out) ;
int amie_b200_bicgstab_resident(amie_b200_ctx * ctx, int precond_kind, double eps, int maxit,
                                uint64_t * nit_out, double * err_out) ;
/* `reps` launches of the block-row SpMV y = A*x on resident vectors; average device
 * milliseconds per launch (CUDA events on the launch stream) in *ms_out.                  */
int amie_b200_spmv_resident(amie_b200_ctx * ctx, int reps, int variant, double * ms_out) ;

/* ------------------------------------------------------------------ Assembly::cgsolve with the history in HBM
 * cgsolve_resident = the solver part of Assembly::cgsolve (solvers/assembly.cpp:1841-1868) without host vectors:
 *   x0 = Assembly::extrapolate(factor) (:1772-1814) from the two last solutions kept on the device
 *        (fewer than two: the resident x as it is -- "displacements"; size changed: history cleared, x0 = 0);
 *   ConjugateGradient::solve(x0, precond, eps, -1) with nssor / rowstart / colstart;
 *   displacementHistory update (:1859-1868).
 * The forces must be resident (upload_rhs / set_boundary_conditions); the solution stays in x (download_x,
 * element_fields).  Returns 1/0 like the solver, or <0.  The reference passes factor = 1.
 * extrapolate / push_history expose the two halves (tests; hosts that run their own solver sequence):
 * extrapolate writes x0 into the resident x (and x0_out if not NULL), *case_out = 0 no history (x untouched),
 * 1 extrapolated, 2 size mismatch (history cleared, x zeroed).                                                   */
int amie_b200_cgsolve_resident(amie_b200_ctx * ctx, int precond_kind, double eps, uint64_t nssor,
                               uint64_t rowstart, uint64_t colstart, double factor,
                               uint64_t * nit_out, double * err_out, double * rho_out) ;
int amie_b200_extrapolate(amie_b200_ctx * ctx, double factor, double * x0_out, int * case_out) ;
int amie_b200_push_history(amie_b200_ctx * ctx) ;
int amie_b200_reset_history(amie_b200_ctx * ctx) ;

/* ------------------------------------------------------------------ device-side value assembly
 * The step that PRODUCES the array the solve consumes, for repeated re-solves on one topology (damage steps):
 * upload the elementary matrices that changed instead of the whole padded array.  Results are bit-identical to
 * the reference's (element-order Kahan sums replayed per stored entry; see csrc/assemble.cu).
 *
 * set_elements -- once per topology, after set_structure: element e couples the nodes (block rows)
 *   elem_ids[e*npe .. e*npe+npe) ; 0xFFFFFFFF marks an unused slot (elements with fewer nodes).  Element ORDER is the
 *   order Assembly::element2d / element3d holds them in (it fixes the summation order).  Builds, on the device, the
 *   list of elementary blocks landing on every stored block.  AMIE_B200_ERR_ARG if an element couples two nodes
 *   whose block is not in the sparsity pattern.
 * update_elements -- elementary matrices of elements [first, first+count):
 *   ke[((e-first)*npe + j)*npe + k)*s*s + m*s + n] = getCachedElementaryMatrix()[j][k][n][m]  (blocks column-major),
 *   scales (NULL = 1) = Assembly::scales.  Marks the stored blocks those elements touch for re-accumulation.
 *   Elements never uploaded contribute zero.
 * assemble -- re-accumulates the marked stored blocks (all of them the first time) in element order with the
 *   reference's per-entry compensation: Assembly::make_final scatter loops, solvers/assembly.cpp:657-735 (2D),
 *   :1060-1138 (3D).  Viscous/space-time second matrices are not handled.  Afterwards the ctx holds values.       */
int amie_b200_set_elements(amie_b200_ctx * ctx, uint64_t n_elem, int npe, const uint32_t * elem_ids) ;
int amie_b200_update_elements(amie_b200_ctx * ctx, uint64_t first, uint64_t count, const double * ke, const double * scales) ;
int amie_b200_assemble(amie_b200_ctx * ctx) ;

/* Assembly::setBoundaryConditions (solvers/assembly.cpp:125-330) on the resident matrix and force vector
 * (upload_rhs first; read the result back with download_rhs or solve with the *_resident calls):
 *   - fix_ids/fix_values: displacement-type multipliers (SET_ALONG_*, FIX_ALONG_*; everything the reference
 *     eliminates at :170-253): column folded into the forces, row replaced by the identity row, forces[id] = value;
 *   - force_ids/force_values: SET_FORCE_* multipliers, forces[id] += value (:262-268);
 *   - add_to_forces (nullable, N): forces += addToExternalForces with the entries of fixed dofs taken as 0 (:177, :323);
 *   - natural_inout (nullable, N): naturalBoundaryConditionForces, receives the same subtractions as the forces.
 * Both id lists ascending and unique (Assembly sorts its multipliers by id, :428).  GENERAL and
 * SET_PROPORTIONAL_DISPLACEMENT multipliers are not handled (AMIE_B200_ERR_ARG is NOT raised for them: the caller
 * simply must not route such assemblies here).  Stored blocks touched by the elimination are marked, so the next
 * assemble() restores them from the elements, as the reference's mask does (:539-560).                          */
int amie_b200_set_boundary_conditions(amie_b200_ctx * ctx, uint64_t nfix, const uint32_t * fix_ids, const double * fix_values,
                                      uint64_t nforce, const uint32_t * force_ids, const double * force_values,
                                      const double * add_to_forces, double * natural_inout) ;

/* ------------------------------------------------------------------ field recovery after the solve
 * What FeatureTree::stepElements / ElementState::getField do with the solution, one element and one virtual call at a
 * time (elements/integrable_entity.cpp:3607-3667, :964-1104, :1379-1392): gather the element's dofs, total strain from
 * the shape-function derivatives and the element's cached inverse Jacobian, mechanical strain = total - imposed
 * strain, real stress = tensor * mechanical strain - imposed stress.  Here: one pass on the device, reading the
 * RESIDENT solution of the last solve (no download of u).  Results equal ElementState::getField's bit for bit.
 *
 * set_element_kinematics -- once per topology, after set_structure.  dim = 2|3 and must equal the stride.
 *   elem_ids[e*npe + j]: node (block-row) id of slot j, shape functions first, then enrichment functions
 *   (IntegrableEntity::getDofIds order); 0xFFFFFFFF = unused slot.  A dof beyond the solution vector reads as 0 (:3641-3648).
 *   dshape[(e*npe + j)*dim + d] = vm.deval(function j, XI|ETA|ZETA, p) at the evaluation point p of the element;
 *   jinv[(e*dim + a)*dim + b]   = (*ElementState::JinvCache)[a][b].
 * set_element_behaviour -- whenever behaviours change (damage): a table of n_tensors entries,
 *   tensors[(t*nc + i)*nc + k] = getBehaviour()->getTensor(p)[i][k], nc = 3 (2D) | 6 (3D);
 *   imposed_strain / imposed_stress [t*nc + i] (NULL = none: getImposedStrain / getImposedStress);
 *   tensor_of_elem[e] = table entry of element e (NULL: n_tensors == n_elem, entry e).
 * element_fields -- u == NULL: use the resident x (the solution of the last *_resident / host-buffer solve);
 *   otherwise u[n_u] is uploaded first.  Outputs [e*nc + i], any of them may be NULL:
 *   TOTAL_STRAIN_FIELD, MECHANICAL_STRAIN_FIELD, REAL_STRESS_FIELD.                                                  */
int amie_b200_set_element_kinematics(amie_b200_ctx * ctx, uint64_t n_elem, int npe, int dim, const uint32_t * elem_ids,
                                     const double * dshape, const double * jinv) ;
int amie_b200_set_element_behaviour(amie_b200_ctx * ctx, uint64_t n_tensors, const double * tensors,
                                    const double * imposed_strain, const double * imposed_stress,
                                    const uint32_t * tensor_of_elem) ;
int amie_b200_element_fields(amie_b200_ctx * ctx, const double * u, uint64_t n_u,
                             double * total_strain_out, double * mechanical_strain_out, double * real_stress_out) ;
/* Principal values of a field element_fields left on the device: field 0 total strain, 1 mechanical strain (both with
 * engineering shears: toPrincipal(..., DOUBLE_OFF_DIAGONAL_VALUES)), 2 real stress (SINGLE_OFF_DIAGONAL_VALUES) --
 * getField(PRINCIPAL_TOTAL_STRAIN_FIELD | PRINCIPAL_MECHANICAL_STRAIN_FIELD | PRINCIPAL_REAL_STRESS_FIELD),
 * elements/integrable_entity.cpp:475-596, :1236-1262, :1505-1509.  principal_out[e*dim + i].  2D: the bits of the
 * reference; 3D: pow / atan2 / cos / sin of the device math library (last-place differences from glibc).            */
int amie_b200_element_principal(amie_b200_ctx * ctx, int field, double * principal_out) ;

/* ------------------------------------------------------------------ statistics */
typedef struct amie_b200_stats
{
    uint64_t stride, nb, nnzb, ndof ;
    uint64_t spmv_launches ;         /* block-row SpMV launches in the last solve (incl. smoothing/residual) */
    uint64_t kernel_launches ;       /* all kernels launched by the last solve                               */
    uint64_t smoothing_spmv ;        /* of which pre/post-smoothing + residual SpMVs                          */
    uint64_t iterations ;            /* inner iterations (the reference's nit)                                */
    uint64_t restarts ;
    double   spmv_ms_total ;         /* sum of CUDA-event durations of the timed SpMV launches                */
    uint64_t spmv_timed ;            /* number of SpMV launches that carried events                           */
    double   solve_ms ;              /* device time of the last solve (events)                                */
    double   h2d_ms, d2h_ms ;        /* copies done by the last host-buffer call                              */
    uint64_t h2d_bytes, d2h_bytes ;
    double   structure_ms, values_ms ; /* last set_structure / set_values (wall)                              */
    uint64_t spmv_algorithmic_bytes ;  /* nnzb*(8 s^2 + 4) + 4 (nb+1) + 16 N                 */
    uint64_t device_bytes ;          /* HBM held by the context                                               */
    double   elements_ms ;           /* last set_elements (wall: gather-list build)                           */
    double   assemble_ms ;           /* last assemble (device time, CUDA events)                              */
    double   bc_ms ;                 /* last set_boundary_conditions kernel (device time)                     */
    uint64_t element_blocks ;        /* n_elem * npe^2 elementary blocks held on the device                   */
    double   fields_ms ;             /* last element_fields kernel (device time)                              */
    uint64_t field_elements ;        /* elements held for field recovery                                      */
    uint64_t early_return ;          /* 1: the last solve returned before its loop -- where the reference prints    */
                                     /* its "homogeneous" line (conjugategradient.cpp:74-78) or nothing at all      */
                                     /* (biconjugategradientstabilized.cpp:43-44, :57-63) instead of "converged"    */
} amie_b200_stats ;
int amie_b200_get_stats(const amie_b200_ctx * ctx, amie_b200_stats * out) ;

/* option keys (anything else: AMIE_B200_ERR_ARG):
 *   "time_spmv"       0/1: record a CUDA-event pair around every SpMV launch (stats spmv_ms_total / spmv_timed);
 *   "spmv_variant"    kernel selection of the block-row SpMV, 0 = the shipped choice;
 *   "verbose"         0/1: print the reference's cerr lines from inside the library;
 *   "iters_per_batch" iterations queued per poll of the device-side loop state (0 = automatic);
 *   "graph"           -1 automatic, 0 / 1: replay the iteration batches as a CUDA graph (small, launch-bound systems);
 *   "fields_variant"  1 (default) the unrolled field-recovery kernel for linear triangles / tetrahedra, 0 the generic
 *                     slot loop (same bits, csrc/kernels_fields.cuh).                                          */
int amie_b200_set_option(amie_b200_ctx * ctx, const char * key, int64_t value) ;

/* ------------------------------------------------------------------ synthetic problems
 * Structured 2D/3D elastic systems in the reference's storage layout:
 * preset in {"S3-hex","S3-tet","S2-tri","ASR-hex"}; n = nodes per side.  Host-only
 * (no GPU needed) except amie_b200_synth_to_device.                                       */
typedef struct amie_b200_synth amie_b200_synth ;
amie_b200_synth * amie_b200_synth_create(const char * preset, int n, uint64_t seed) ;
void amie_b200_synth_destroy(amie_b200_synth * s) ;
int  amie_b200_synth_sizes(const amie_b200_synth * s, int * stride, uint64_t * nb, uint64_t * nnzb) ;
/* rows [row0,row1): row_size[row1-row0]; nnzb of the range in *nnzb_out                   */
int  amie_b200_synth_count(const amie_b200_synth * s, uint64_t row0, uint64_t row1,
                           uint32_t * row_size, uint64_t * nnzb_out) ;
/* fills column_index (global block columns), array (reference padded layout) and b for
 * rows [row0,row1); any of the three may be NULL.                                         */
int  amie_b200_synth_fill(const amie_b200_synth * s, uint64_t row0, uint64_t row1,
                          uint32_t * column_index, double * array_padded, double * b) ;
/* Generates structure + values + rhs of the whole system directly in HBM (same generator
 * compiled for the device; no host arrays): for sizes where the padded host array (43 GB at
 * 50 M DOF) is impractical.  Afterwards the ctx is as after set_structure+set_values+upload_rhs. */
int  amie_b200_synth_to_device(amie_b200_ctx * ctx, const amie_b200_synth * s) ;

/* ------------------------------------------------------------------ row partition (multi-GPU, host-only)
 * Contiguous block-row ranges balanced by stored blocks.
 * bounds_out[nparts+1].                                                                    */
int amie_b200_partition_rows(uint64_t nb, const uint32_t * row_size, int nparts, uint64_t * bounds_out) ;
/* For part [r0,r1): the distinct off-range block columns its rows reference, ascending
 * (= the halo it must receive).  Pass halo_out == NULL to get the count.                  */
int amie_b200_partition_halo(uint64_t r0, uint64_t r1, const uint32_t * row_size_local,
                             const uint32_t * column_index_local, uint32_t * halo_out, uint64_t * nhalo_out) ;

/* ------------------------------------------------------------------ node renumbering (host-only)
 * The mesher's numbering has no locality; these are the host half of a renumbering applied once per topology
 * (csrc/reorder.cpp; the device half is amie_b200_set_block_map: values scattered through block_to inside K-Repack,
 * the host permutes the vectors).  perm_out[old node] = new node: reverse Cuthill-McKee on the block graph.      */
int amie_b200_rcm_order(uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, uint32_t * perm_out) ;
/* Opt-in refinement of a numbering (perm_inout[old node] = new node, e.g. the one above): inside every window of
 * `window` consecutive nodes of it, the nodes are re-ordered by row length, longest first.  A tile of the row-thread
 * SpMV costs its longest row; FeatureTree-assembled 3D systems carry a 1.7-1.9x imbalance per tile (csrc/reorder.cpp).
 * row_size is in the ORIGINAL numbering.  AMIE_B200_ERR_ARG if perm_inout is not a permutation or window == 0.       */
int amie_b200_group_rows_by_length(uint64_t nb, const uint32_t * row_size, uint64_t window, uint32_t * perm_inout) ;
/* The structure in the new numbering (columns ascending inside each row) and, for every stored block of it, the
 * stored block of the old structure it is (block_from_out[new k] = old k): array_new block k = array_old block
 * block_from[k].  AMIE_B200_ERR_ARG if perm is not a permutation.                                                  */
int amie_b200_permute_structure(uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, const uint32_t * perm,
                                uint32_t * row_size_out, uint32_t * column_index_out, uint32_t * block_from_out) ;

/* ------------------------------------------------------------------ distributed context (one process per GPU)
 * Rank r owns block rows [bounds[r], bounds[r+1]).  The NCCL communicator is created inside the
 * library from a 128-byte ncclUniqueId the caller broadcasts (e.g. with torch.distributed).   */
int amie_b200_nccl_unique_id(void * id128_out) ;
int amie_b200_dist_init(amie_b200_ctx * ctx, int rank, int world, const void * id128,
                        const uint64_t * bounds /* world+1 */) ;
/* local rows of this rank, GLOBAL block column indices                                      */
int amie_b200_dist_set_structure(amie_b200_ctx * ctx, int stride, uint64_t nb_global,
                                 const uint32_t * row_size_local, const uint32_t * column_index_local,
                                 uint64_t nnzb_local) ;
int amie_b200_dist_synth_to_device(amie_b200_ctx * ctx, const amie_b200_synth * s) ;
/* halo block columns received / owned block columns sent per SpMV, interior block rows, number of peers */
/* 1: halo + reductions go over NVLink peer memory (cudaIpc mappings, device-side flags); 0: NCCL calls.
 * Env AMIE_B200_TRANSPORT=nccl forces the NCCL path.                                              */
int amie_b200_dist_transport(const amie_b200_ctx * ctx) ;
int amie_b200_dist_info(const amie_b200_ctx * ctx, uint64_t * nhalo_out, uint64_t * nsend_out,
                        uint64_t * interior_rows_out, int * npeers_out) ;

#ifdef __cplusplus
}
#endif
#endif
