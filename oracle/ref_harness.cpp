// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never on the product path).
//
// A C-ABI around the UNMODIFIED reference solver, compiled by oracle/build_ref.py against the
// sources under /root/reference into oracle/_ref/libamie_ref_oracle.so.  It drives the reference
// exactly the way SURVEY.md §8(c) "primary oracle" describes: a hand-filled Amie::Assembly
// (all members public, solvers/assembly.h:230-264; getForces() skips make_final when the matrix
// pointer is set, solvers/assembly.cpp:90-96), then
//   Amie::ConjugateGradient::solve            (solvers/conjugategradient.cpp:69-318)
//   Amie::BiConjugateGradientStabilized::solve (solvers/biconjugategradientstabilized.cpp:12-148)
//   Amie::assign(y, A*x[-b], rowstart, colstart) (sparse/sparse_matrix.cpp:462-547)
//   CoordinateIndexedSparseMatrix::inverseDiagonal (sparse/sparse_matrix.cpp:216-231)
//   Amie::Assembly::setBoundaryConditions     (solvers/assembly.cpp:125-383)
// No reference source is copied here; only its public headers are included.

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <sstream>
#include <chrono>
#include <string>
#ifdef HAVE_OPENMP
#include <omp.h>
#endif

#include "solvers/assembly.h"
#include "solvers/conjugategradient.h"
#include "solvers/biconjugategradientstabilized.h"
#include "solvers/inversediagonal.h"
#include "solvers/preconditionners.h"
#include "sparse/sparse_matrix.h"
#include "utilities/matrixops.h"

namespace {

struct CerrCapture
{
    std::streambuf * old ;
    std::ostringstream sink ;
    CerrCapture() : old(std::cerr.rdbuf(sink.rdbuf())) {}
    ~CerrCapture() { std::cerr.rdbuf(old) ; }
} ;

double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() ;
}

void fill_assembly(Amie::Assembly & a, int stride, uint64_t nb, const uint32_t * row_size,
                   const uint32_t * column_index, uint64_t nnzb, const double * array_padded,
                   const double * b)
{
    std::valarray<unsigned int> rs(row_size, nb) ;
    std::valarray<unsigned int> ci(column_index, nnzb) ;
    a.coordinateIndexedMatrix = new Amie::CoordinateIndexedSparseMatrix(rs, ci, (size_t)stride) ;
    Vector & arr = a.coordinateIndexedMatrix->array ;
    std::memcpy(&arr[0], array_padded, arr.size()*sizeof(double)) ;
    a.externalForces.resize(nb*stride) ;
    if(b)
        std::memcpy(&a.externalForces[0], b, nb*stride*sizeof(double)) ;
    else
        a.externalForces = 0. ;
    a.displacements.resize(nb*stride) ;
    a.displacements = 0. ;
}

// a user-written preconditioner, as a caller of LinearSolver::solve(x0, precond, ...) may pass one:
// precondition() is t = v .* d for a diagonal the user chose
struct UserDiagonal : public Amie::Preconditionner
{
    Vector d ;
    UserDiagonal(const double * p, size_t n) : d(p, n) { }
    virtual ~UserDiagonal() { }
    virtual void precondition(const Vector & v, Vector & t) { for(size_t i = 0 ; i < v.size() ; i++) t[i] = v[i]*d[i] ; }
} ;

const double * g_user_diag = nullptr ;
uint64_t g_user_diag_n = 0 ;

// precond_kind: 0 nullptr (the solver builds its InverseDiagonal), 1 NullPreconditionner, 2 InverseDiagonalSquared,
// 3 InverseLumpedDiagonal, 4 UserDiagonal over the vector given to amie_ref_set_user_diagonal.  Caller deletes.
Amie::Preconditionner * make_precond(int kind, Amie::Assembly & a)
{
    switch(kind)
    {
        case 1: return new Amie::NullPreconditionner() ;
        case 2: return new Amie::InverseDiagonalSquared(a.getMatrix()) ;
        case 3: return new Amie::InverseLumpedDiagonal(a.getMatrix()) ;
        case 4: return new UserDiagonal(g_user_diag, g_user_diag_n) ;
        case 5: return new Amie::Inverse2x2Diagonal(a.getMatrix()) ;
        default: return nullptr ;
    }
}

void copy_log(const std::string & s, char * log, uint64_t logcap)
{
    if(!log || !logcap) return ;
    size_t n = std::min<size_t>(s.size(), logcap-1) ;
    // keep the tail: the "converged after" line is printed last
    std::memcpy(log, s.data()+(s.size()-n), n) ;
    log[n] = 0 ;
}

}

extern "C" {

// the diagonal amie_ref_cg / amie_ref_bicgstab hand to the solver as a UserDiagonal when precond_kind == 4
void amie_ref_set_user_diagonal(const double * d, uint64_t n) { g_user_diag = d ; g_user_diag_n = n ; }

// the diagonal the reference's own preconditioner classes build from the matrix: kind 0 InverseDiagonal,
// 2 InverseDiagonalSquared, 3 InverseLumpedDiagonal (solvers/inversediagonal.cpp:19-42, :50, :69-73)
int amie_ref_precond_diagonal(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, uint64_t nnzb,
                              const double * array_padded, int kind, double * d_out)
{
    Amie::Assembly a ;
    fill_assembly(a, stride, nb, row_size, column_index, nnzb, array_padded, nullptr) ;
    const size_t n = nb*stride ;
    if(kind == 0) { Amie::InverseDiagonal P(a.getMatrix()) ; std::memcpy(d_out, &P.diagonal[0], n*sizeof(double)) ; }
    else if(kind == 2) { Amie::InverseDiagonalSquared P(a.getMatrix()) ; std::memcpy(d_out, &(*P.diagonal)[0], n*sizeof(double)) ; }
    else if(kind == 3) { Amie::InverseLumpedDiagonal P(a.getMatrix()) ; std::memcpy(d_out, &P.diagonal[0], n*sizeof(double)) ; }
    else return -1 ;
    return 0 ;
}

// the 2x2 blocks Inverse2x2Diagonal holds (solvers/inversediagonal.cpp:84-119), row-major, one per dof pair
int amie_ref_precond_blocks2(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, uint64_t nnzb,
                             const double * array_padded, double * blocks_out)
{
    Amie::Assembly a ;
    fill_assembly(a, stride, nb, row_size, column_index, nnzb, array_padded, nullptr) ;
    Amie::Inverse2x2Diagonal P(a.getMatrix()) ;
    for(size_t i = 0 ; i < P.blocks.size() ; i++)
        for(int r = 0 ; r < 2 ; r++)
            for(int c = 0 ; c < 2 ; c++)
                blocks_out[i*4+r*2+c] = P.blocks[i][r][c] ;
    return (int)P.blocks.size() ;
}

// Amie::det and Amie::invert3x3Matrix (utilities/matrixops.cpp:826-847, :681-702) on n row-major 3x3 matrices
void amie_ref_det_invert3x3(const double * m_in, uint64_t n, double * det_out, double * inv_out)
{
    for(uint64_t k = 0 ; k < n ; k++)
    {
        Amie::Matrix M(3, 3) ;
        for(int r = 0 ; r < 3 ; r++) for(int c = 0 ; c < 3 ; c++) M[r][c] = m_in[k*9+r*3+c] ;
        det_out[k] = Amie::det(M) ;
        Amie::invert3x3Matrix(M) ;
        for(int r = 0 ; r < 3 ; r++) for(int c = 0 ; c < 3 ; c++) inv_out[k*9+r*3+c] = M[r][c] ;
    }
}

// Assembly::extrapolate(factor) (solvers/assembly.cpp:1772-1814) on a hand-filled displacementHistory {prev, back}
// (nhist = 2) or {back} alone (nhist = 1), with `displacements` = disp (ndisp entries).  Returns the size of the
// vector the reference returned (copied to out, which holds max(n, ndisp) doubles); back_out receives the newest
// history vector afterwards (the reference scrubs NaNs in it) when the history survived.
uint64_t amie_ref_extrapolate(const double * prev, const double * back, uint64_t n, int nhist, const double * disp, uint64_t ndisp,
                              double factor, double * out, double * back_out, uint64_t * hist_size_out)
{
    Amie::Assembly a ;
    if(nhist == 2) a.displacementHistory.push_back(Vector(prev, n)) ;
    if(nhist >= 1) a.displacementHistory.push_back(Vector(back, n)) ;
    a.displacements.resize(ndisp) ;
    if(ndisp) std::memcpy(&a.displacements[0], disp, ndisp*sizeof(double)) ;
    Vector r = a.extrapolate(factor) ;
    if(r.size()) std::memcpy(out, &r[0], r.size()*sizeof(double)) ;
    if(hist_size_out) *hist_size_out = a.displacementHistory.size() ;
    if(back_out && a.displacementHistory.size() && a.displacementHistory.back().size() == n && n)
        std::memcpy(back_out, &a.displacementHistory.back()[0], n*sizeof(double)) ;
    return r.size() ;
}

int amie_ref_max_threads()
{
#ifdef HAVE_OPENMP
    return omp_get_max_threads() ;
#else
    return 1 ;
#endif
}

// returns 1 converged / 0 not converged.  precond_kind: see make_precond
int amie_ref_cg(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, uint64_t nnzb,
                const double * array_padded, const double * b, const double * x0, uint64_t nx0,
                int precond_kind, double eps, int maxit, uint64_t nssor, uint64_t rowstart, uint64_t colstart,
                int nthreads, double * x_out, uint64_t * nit_out, double * wall_s_out, char * log, uint64_t logcap)
{
#ifdef HAVE_OPENMP
    if(nthreads > 0) omp_set_num_threads(nthreads) ;
#endif
    Amie::Assembly a ;
    fill_assembly(a, stride, nb, row_size, column_index, nnzb, array_padded, b) ;
    CerrCapture cap ;
    Amie::ConjugateGradient cg(&a) ;
    cg.nssor = nssor ;
    cg.rowstart = rowstart ;
    cg.colstart = colstart ;
    Vector vx0(0., nx0) ;
    if(nx0) std::memcpy(&vx0[0], x0, nx0*sizeof(double)) ;
    Amie::Preconditionner * P = make_precond(precond_kind, a) ;
    double t0 = now() ;
    bool ok = cg.solve(vx0, P, eps, maxit, false) ;
    double t1 = now() ;
    delete P ;                       // the solver does not own a preconditioner it was given (cleanup == false)
    std::memcpy(x_out, &cg.x[0], cg.x.size()*sizeof(double)) ;
    if(nit_out) *nit_out = cg.nit ;
    if(wall_s_out) *wall_s_out = t1-t0 ;
    copy_log(cap.sink.str(), log, logcap) ;
    return ok ? 1 : 0 ;
}

// The same solve on a system the caller generates IN PLACE: `fill(user, column_index, array_padded, forces)` writes
// straight into the reference's own storage, so a benchmark-size matrix (43 GB of padded values at 50 M DOF) exists
// once in host memory, not twice.  spmv_reps > 0 additionally times assign(y, A*b) on that matrix (seconds per call).
typedef int (*amie_ref_fill_fn)(void * user, uint32_t * column_index, double * array_padded, double * forces) ;
int amie_ref_cg_fill_x0(int stride, uint64_t nb, const uint32_t * row_size, uint64_t nnzb, amie_ref_fill_fn fill, void * user,
                        const double * x0, uint64_t nx0,
                        double eps, int maxit, uint64_t nssor, int nthreads, int spmv_reps,
                        double * x_out, uint64_t * nit_out, double * wall_s_out, double * spmv_s_out, char * log, uint64_t logcap) ;

int amie_ref_cg_fill(int stride, uint64_t nb, const uint32_t * row_size, uint64_t nnzb, amie_ref_fill_fn fill, void * user,
                     double eps, int maxit, uint64_t nssor, int nthreads, int spmv_reps,
                     double * x_out, uint64_t * nit_out, double * wall_s_out, double * spmv_s_out, char * log, uint64_t logcap)
{
    return amie_ref_cg_fill_x0(stride, nb, row_size, nnzb, fill, user, nullptr, 0, eps, maxit, nssor, nthreads, spmv_reps,
                               x_out, nit_out, wall_s_out, spmv_s_out, log, logcap) ;
}

// the same from a starting vector (ConjugateGradient::solve(x0, ...), conjugategradient.cpp:95-104): a bounded sample of a
// long solve is "start where a shorter solve stopped, tighten the tolerance a little"
int amie_ref_cg_fill_x0(int stride, uint64_t nb, const uint32_t * row_size, uint64_t nnzb, amie_ref_fill_fn fill, void * user,
                        const double * x0, uint64_t nx0,
                        double eps, int maxit, uint64_t nssor, int nthreads, int spmv_reps,
                        double * x_out, uint64_t * nit_out, double * wall_s_out, double * spmv_s_out, char * log, uint64_t logcap)
{
#ifdef HAVE_OPENMP
    if(nthreads > 0) omp_set_num_threads(nthreads) ;
#endif
    Amie::Assembly a ;
    {
        std::valarray<unsigned int> rs(row_size, nb) ;
        std::valarray<unsigned int> ci(0u, nnzb) ;
        a.coordinateIndexedMatrix = new Amie::CoordinateIndexedSparseMatrix(rs, ci, (size_t)stride) ;
    }
    a.externalForces.resize(nb*stride) ;
    a.externalForces = 0. ;
    a.displacements.resize(nb*stride) ;
    a.displacements = 0. ;
    static_assert(sizeof(unsigned int) == sizeof(uint32_t), "column_index element type") ;
    if(fill(user, reinterpret_cast<uint32_t *>(&a.coordinateIndexedMatrix->column_index[0]), &a.coordinateIndexedMatrix->array[0], &a.externalForces[0]))
        return -1 ;
    CerrCapture cap ;
    if(spmv_reps > 0 && spmv_s_out)
    {
        Vector y(0., nb*stride) ;
        Amie::assign(y, (*a.coordinateIndexedMatrix)*a.externalForces, 0, 0) ;          // first touch
        double t0 = now() ;
        for(int r = 0 ; r < spmv_reps ; r++) Amie::assign(y, (*a.coordinateIndexedMatrix)*a.externalForces, 0, 0) ;
        *spmv_s_out = (now()-t0)/spmv_reps ;
    }
    Amie::ConjugateGradient cg(&a) ;
    cg.nssor = nssor ;
    Vector vx0(0., x0 ? nx0 : 0) ;
    if(x0 && nx0) std::memcpy(&vx0[0], x0, nx0*sizeof(double)) ;
    double t0 = now() ;
    bool ok = cg.solve(vx0, nullptr, eps, maxit, false) ;
    double t1 = now() ;
    if(x_out) std::memcpy(x_out, &cg.x[0], cg.x.size()*sizeof(double)) ;
    if(nit_out) *nit_out = cg.nit ;
    if(wall_s_out) *wall_s_out = t1-t0 ;
    copy_log(cap.sink.str(), log, logcap) ;
    return ok ? 1 : 0 ;
}

int amie_ref_bicgstab(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, uint64_t nnzb,
                      const double * array_padded, const double * b, const double * x0, uint64_t nx0,
                      int precond_kind, double eps, int maxit, int nthreads,
                      double * x_out, uint64_t * nit_out, double * wall_s_out, char * log, uint64_t logcap)
{
#ifdef HAVE_OPENMP
    if(nthreads > 0) omp_set_num_threads(nthreads) ;
#endif
    Amie::Assembly a ;
    fill_assembly(a, stride, nb, row_size, column_index, nnzb, array_padded, b) ;
    CerrCapture cap ;
    Amie::BiConjugateGradientStabilized cg(&a) ;
    Vector vx0(0., nx0) ;
    if(nx0) std::memcpy(&vx0[0], x0, nx0*sizeof(double)) ;
    Amie::Preconditionner * P = make_precond(precond_kind, a) ;
    double t0 = now() ;
    bool ok = cg.solve(vx0, P, eps, maxit, true) ;
    double t1 = now() ;
    delete P ;
    std::memcpy(x_out, &cg.x[0], cg.x.size()*sizeof(double)) ;
    if(wall_s_out) *wall_s_out = t1-t0 ;
    std::string s = cap.sink.str() ;
    if(nit_out)
    {
        // the reference keeps nit local; it only appears in the verbose cerr line
        // " BiCGStab <n> converged after <nit> iterations" / "did not converge after <nit>"
        *nit_out = 0 ;
        size_t pos = s.rfind(" after ") ;
        if(pos != std::string::npos)
            *nit_out = std::strtoull(s.c_str()+pos+7, nullptr, 10) ;
    }
    copy_log(s, log, logcap) ;
    return ok ? 1 : 0 ;
}

// mode 0: assign(y, A*x, rowstart, colstart)          (Kahan, OpenMP tasks)
// mode 1: assign(y, A*x - b, rowstart, colstart)
// mode 2: y = A*x      via operator Vector()           (serial, non-Kahan)
// mode 3: y = A*x - b  via operator const Vector()     (serial, non-Kahan)
int amie_ref_spmv(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, uint64_t nnzb,
                  const double * array_padded, const double * x, const double * b, int mode,
                  uint64_t rowstart, uint64_t colstart, int nthreads, int reps, double * y_out, double * wall_s_out)
{
#ifdef HAVE_OPENMP
    if(nthreads > 0) omp_set_num_threads(nthreads) ;
#endif
    Amie::Assembly a ;
    fill_assembly(a, stride, nb, row_size, column_index, nnzb, array_padded, b) ;
    const Amie::CoordinateIndexedSparseMatrix & A = *a.coordinateIndexedMatrix ;
    Vector vx(0., nb*stride) ;
    std::memcpy(&vx[0], x, nb*stride*sizeof(double)) ;
    Vector y(0., nb*stride) ;
    if(reps < 1) reps = 1 ;
    double t0 = now() ;
    for(int r = 0 ; r < reps ; r++)
    {
        switch(mode)
        {
        case 0: Amie::assign(y, A*vx, (int)rowstart, (int)colstart) ; break ;
        case 1: Amie::assign(y, A*vx-a.externalForces, (int)rowstart, (int)colstart) ; break ;
        case 2: y = (Vector)(A*vx) ; break ;
        case 3: y = (Vector)(A*vx-a.externalForces) ; break ;
        default: return -1 ;
        }
    }
    double t1 = now() ;
    std::memcpy(y_out, &y[0], nb*stride*sizeof(double)) ;
    if(wall_s_out) *wall_s_out = (t1-t0)/reps ;
    return 0 ;
}

int amie_ref_inverse_diagonal(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, uint64_t nnzb,
                              const double * array_padded, double * d_out)
{
    Amie::Assembly a ;
    fill_assembly(a, stride, nb, row_size, column_index, nnzb, array_padded, nullptr) ;
    Vector d = a.coordinateIndexedMatrix->inverseDiagonal() ;
    std::memcpy(d_out, &d[0], d.size()*sizeof(double)) ;
    return 0 ;
}

// The unmodified Assembly::setBoundaryConditions on a hand-filled Assembly.  `dim` is set to
// SPACE_ONE_DIMENSIONAL so that the space-time rowstart block at the end of the function
// (solvers/assembly.cpp:326-360), which dereferences element2d[0]/element3d[0], is not entered:
// the elimination itself (:137-324) does not look at `dim`.
// fix_type: 0 = SET_ALONG_XI, 1 = SET_ALONG_ETA, 2 = SET_ALONG_ZETA, 3 = SET_ALONG_INDEXED_AXIS (all eliminated alike).
int amie_ref_set_boundary_conditions(int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * column_index,
                                     uint64_t nnzb, double * array_padded, double * forces, double * natural,
                                     double * add_to_forces,
                                     uint64_t nfix, const uint32_t * fix_ids, const double * fix_values,
                                     uint64_t nforce, const uint32_t * force_ids, const double * force_values)
{
    CerrCapture quiet ;
    Amie::Assembly a ;
    fill_assembly(a, stride, nb, row_size, column_index, nnzb, array_padded, forces) ;
    const size_t n = nb*stride ;
    a.naturalBoundaryConditionForces.resize(n) ;
    a.naturalBoundaryConditionForces = 0. ;
    if(natural) std::memcpy(&a.naturalBoundaryConditionForces[0], natural, n*sizeof(double)) ;
    if(add_to_forces)
    {
        a.addToExternalForces.resize(n) ;
        std::memcpy(&a.addToExternalForces[0], add_to_forces, n*sizeof(double)) ;
    }
    else
        a.addToExternalForces.resize(0) ;
    a.dim = Amie::SPACE_ONE_DIMENSIONAL ;
    a.ndof = stride ;
    const Amie::LagrangeMultiplierType kinds[3] = { Amie::SET_ALONG_XI, Amie::SET_ALONG_ETA, Amie::SET_ALONG_ZETA } ;
    const Amie::LagrangeMultiplierType fkinds[3] = { Amie::SET_FORCE_XI, Amie::SET_FORCE_ETA, Amie::SET_FORCE_ZETA } ;
    for(uint64_t i = 0 ; i < nfix ; i++)
    {
        Amie::LagrangeMultiplier m(std::valarray<unsigned int>(), Vector(), fix_values[i], (int)fix_ids[i]) ;
        m.type = kinds[fix_ids[i]%stride%3] ;
        a.multipliers.push_back(m) ;
    }
    for(uint64_t i = 0 ; i < nforce ; i++)
    {
        Amie::LagrangeMultiplier m(std::valarray<unsigned int>(), Vector(), force_values[i], (int)force_ids[i]) ;
        m.type = fkinds[force_ids[i]%stride%3] ;
        a.multipliers.push_back(m) ;
    }
    std::stable_sort(a.multipliers.begin(), a.multipliers.end()) ;      // make_final sorts them (assembly.cpp:428)
    a.setBoundaryConditions(false) ;
    std::memcpy(array_padded, &a.coordinateIndexedMatrix->array[0], a.coordinateIndexedMatrix->array.size()*sizeof(double)) ;
    std::memcpy(forces, &a.externalForces[0], n*sizeof(double)) ;
    if(natural) std::memcpy(natural, &a.naturalBoundaryConditionForces[0], n*sizeof(double)) ;
    if(add_to_forces) std::memcpy(add_to_forces, &a.addToExternalForces[0], n*sizeof(double)) ;
    return 0 ;
}

}
