// amie_b200.hpp -- header-only C++ mirror of the reference's solver interface over the C-ABI.
//
// Same names, argument meaning and error behaviour as
//   Amie::CoordinateIndexedSparseMatrix   sparse/sparse_matrix.h:129-136
//   Amie::Assembly (getMatrix/getForces)  solvers/assembly.h:228-450, cgsolve solvers/assembly.cpp:1829-1858
//   Amie::LinearSolver                    solvers/solver.h:27-52
//   Amie::ConjugateGradient               solvers/conjugategradient.h:22-43
//   Amie::BiConjugateGradientStabilized   solvers/biconjugategradientstabilized.h:19-24
// for C++ hosts that do not link AMIE itself (AMIE proper uses host/shim/*.cpp instead).
#pragma once
#include <valarray>
#include <stdexcept>
#include <string>
#include <cstdint>
#include "../../include/amie_b200.h"

namespace AmieB200
{
typedef std::valarray<double> Vector ;
const double default_solver_precision = 1e-10 ;          // polynomial/variable.h:14

struct CoordinateIndexedSparseMatrix
{
    size_t stride ;
    Vector array ;
    std::valarray<unsigned int> column_index ;
    std::valarray<unsigned int> row_size ;
    std::valarray<unsigned int> accumulated_row_size ;
    CoordinateIndexedSparseMatrix(const std::valarray<unsigned int> & rs, const std::valarray<unsigned int> & ci, size_t s)
        : stride(s), array(0., ci.size()*s*(s+s%2)), column_index(ci), row_size(rs), accumulated_row_size(rs.size())
    {
        for(size_t i = 1 ; i < accumulated_row_size.size() ; i++)
            accumulated_row_size[i] = accumulated_row_size[i-1]+row_size[i-1] ;
    }
} ;

// solvers/preconditionners.h:19-32, solvers/inversediagonal.h:23-47.  kind() is what the C-ABI is told; diagonal()
// (kind AMIE_B200_PRECOND_DIAGONAL) is uploaded before the solve.
struct Preconditionner
{
    virtual ~Preconditionner() { }
    virtual int kind() const = 0 ;
    virtual const Vector * diagonal() const { return nullptr ; }
} ;
struct NullPreconditionner : public Preconditionner { int kind() const override { return AMIE_B200_PRECOND_NULL ; } } ;
struct InverseDiagonalSquared : public Preconditionner { int kind() const override { return AMIE_B200_PRECOND_DIAGONAL_SQUARED ; } } ;
struct InverseLumpedDiagonal : public Preconditionner { int kind() const override { return AMIE_B200_PRECOND_LUMPED ; } } ;
// a user-written preconditioner whose precondition(v, t) is t = v * d
// node-block preconditioners, PCG only: solvers/inversediagonal.h:49-55 on stride-2 systems, and its 3x3 counterpart
// (no reference class; opt-in, other iteration counts) on stride-3 systems -- built on the device from the matrix
struct Inverse2x2Diagonal : public Preconditionner { int kind() const override { return AMIE_B200_PRECOND_BLOCK2X2 ; } } ;
struct BlockJacobi3x3 : public Preconditionner { int kind() const override { return AMIE_B200_PRECOND_BLOCK3X3 ; } } ;
struct DiagonalPreconditionner : public Preconditionner
{
    Vector d ;
    explicit DiagonalPreconditionner(const Vector & diag) : d(diag) { }
    int kind() const override { return AMIE_B200_PRECOND_DIAGONAL ; }
    const Vector * diagonal() const override { return &d ; }
} ;

class Assembly
{
public:
    CoordinateIndexedSparseMatrix * coordinateIndexedMatrix = nullptr ;
    Vector externalForces ;
    Vector displacements ;
    size_t nssor = 32, rowstart = 0, colstart = 0 ;
    double epsilon = default_solver_precision ;

    explicit Assembly(int device = 0)
    {
        ctx = amie_b200_create(&device, 1) ;
        if(!ctx) throw std::runtime_error(std::string("amie_b200: ")+amie_b200_global_error()+" (no CPU fallback)") ;
    }
    // ONE context over several GPUs (block rows partitioned inside the library; the same calls with the same global
    // arrays).  Needs CUDA_MODULE_LOADING=EAGER in the environment before CUDA initialises (see include/amie_b200.h).
    Assembly(const int * devices, int ndev)
    {
        ctx = amie_b200_create(devices, ndev) ;
        if(!ctx) throw std::runtime_error(std::string("amie_b200: ")+amie_b200_global_error()+" (no CPU fallback)") ;
    }
    ~Assembly() { amie_b200_destroy(ctx) ; }
    Assembly(const Assembly &) = delete ;
    CoordinateIndexedSparseMatrix & getMatrix() { return *coordinateIndexedMatrix ; }
    Vector & getForces() { return externalForces ; }
    void setEpsilon(double e) { epsilon = e ; }
    // structure once per topology, values whenever they were re-assembled
    void structureChanged() { structure_ok = false ; }
    void valuesChanged() { values_ok = false ; }
    amie_b200_ctx * device()
    {
        CoordinateIndexedSparseMatrix & A = getMatrix() ;
        if(!structure_ok)
        {
            check(amie_b200_set_structure(ctx, (int)A.stride, A.row_size.size(), &A.row_size[0], &A.column_index[0], A.column_index.size())) ;
            structure_ok = true ; values_ok = false ;
        }
        if(!values_ok) { check(amie_b200_set_values(ctx, &A.array[0])) ; values_ok = true ; }
        return ctx ;
    }
    int check(int rc) const
    {
        if(rc < 0) throw std::runtime_error(std::string("amie_b200: ")+amie_b200_last_error(ctx)) ;
        return rc ;
    }
    // what the caller passed as Preconditionner* -> precond_kind (uploading a user diagonal first)
    int precondKind(Preconditionner * p)
    {
        if(!p) return AMIE_B200_PRECOND_JACOBI ;
        if(const Vector * d = p->diagonal())
        {
            if(d->size() != getForces().size()) throw std::runtime_error("amie_b200: DiagonalPreconditionner needs one entry per degree of freedom") ;
            check(amie_b200_set_preconditioner_diagonal(device(), &(*d)[0])) ;
        }
        return p->kind() ;
    }
    bool cgsolve(int maxit = -1, bool verbose = true) ;      // solvers/assembly.cpp:1829
private:
    amie_b200_ctx * ctx = nullptr ;
    bool structure_ok = false, values_ok = false ;
} ;

struct LinearSolver
{
    unsigned long colstart = 0 ;
    unsigned long rowstart = 0 ;
    Vector x ;
    Assembly * assembly ;
    explicit LinearSolver(Assembly * a) : x(0., a->getForces().size()), assembly(a) { }
    virtual ~LinearSolver() { }
    virtual bool solve(const Vector & x0, Preconditionner * precond = nullptr, const double eps = default_solver_precision,
                       const int maxit = -1, bool verbose = false) = 0 ;
} ;

struct ConjugateGradient : public LinearSolver
{
    size_t nit = 0 ;
    size_t nssor = 128 ;
    double last_error = 0, last_rho = 0 ;
    explicit ConjugateGradient(Assembly * a) : LinearSolver(a) { }
    bool solve(const Vector & x0, Preconditionner * precond = nullptr, const double eps = default_solver_precision,
               const int maxit = -1, bool verbose = false) override
    {
        amie_b200_ctx * c = assembly->device() ;
        const Vector & b = assembly->getForces() ;
        if(x.size() != b.size()) x.resize(b.size(), 0.) ;
        uint64_t n = 0 ;
        int ret = assembly->check(amie_b200_pcg(c, &b[0], x0.size() ? &x0[0] : nullptr, x0.size(),
                                                assembly->precondKind(precond), eps, maxit, nssor,
                                                rowstart, colstart, &x[0], &n, &last_error, &last_rho)) ;
        nit = n ;
        (void)verbose ;
        return ret == 1 ;
    }
} ;

struct BiConjugateGradientStabilized : public LinearSolver
{
    size_t nit = 0 ;
    double last_error = 0 ;
    explicit BiConjugateGradientStabilized(Assembly * a) : LinearSolver(a) { }
    bool solve(const Vector & x0, Preconditionner * precond = nullptr, const double eps = default_solver_precision,
               const int maxit = -1, bool verbose = false) override
    {
        amie_b200_ctx * c = assembly->device() ;
        const Vector & b = assembly->getForces() ;
        x.resize(b.size(), 0.) ;
        uint64_t n = 0 ;
        int ret = assembly->check(amie_b200_bicgstab(c, &b[0], x0.size() ? &x0[0] : nullptr, x0.size(),
                                                     assembly->precondKind(precond), eps, maxit, &x[0], &n, &last_error)) ;
        nit = n ;
        (void)verbose ;
        return ret == 1 ;
    }
} ;

inline bool Assembly::cgsolve(int, bool verbose)
{
    ConjugateGradient cg(this) ;
    cg.nssor = nssor ;
    if(rowstart > 0 || colstart > 0) { cg.rowstart = rowstart ; cg.colstart = colstart ; }
    bool ret = cg.solve(displacements, nullptr, epsilon, -1, verbose) ;
    displacements.resize(cg.x.size()) ;
    displacements = cg.x ;
    return ret ;
}

}
