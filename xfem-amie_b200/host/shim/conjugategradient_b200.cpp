// conjugategradient_b200.cpp -- drop-in replacement of solvers/conjugategradient.cpp.
//
// Same class (solvers/conjugategradient.h:22-43), same constructor and solve() signature; the body
// of solve() forwards to the C-ABI (include/amie_b200.h).  Link this object INSTEAD of the
// reference's conjugategradient.o; nothing else of AMIE changes.
#include "solvers/conjugategradient.h"
#include "amie_b200_shim.h"
#include <iostream>
#include <cstdlib>
#include <algorithm>

namespace Amie {

ConjugateGradient::ConjugateGradient( Assembly* a ) :LinearSolver(a), r(0), z(0), p(0), q(0), xmin(0), cleanup(false), P(nullptr), nit(0)
{
    // the work vectors r, z, p, q, xmin of the reference live in HBM here
}

bool ConjugateGradient::solve(const Vector &x0, Preconditionner * precond, const double eps, const int maxit, bool verbose)
{
    const Vector * diagonal = nullptr ;
    int kind = AmieB200Shim::precond_kind(precond, &diagonal) ;
    if(kind == -2)
    {
        const size_t stride = assembly->getMatrix().stride ;
        kind = stride == 2 ? AMIE_B200_PRECOND_BLOCK2X2 : (stride == 3 ? AMIE_B200_PRECOND_BLOCK3X3 : AMIE_B200_PRECOND_JACOBI) ;
    }
    if(kind < 0)
    {
        std::cerr << "amie_b200: this Preconditionner type is not available on the device (nullptr, NullPreconditionner and the diagonal classes of solvers/inversediagonal.h are)" << std::endl ;
        return false ;
    }
    amie_b200_ctx * ctx = AmieB200Shim::context_for(assembly) ;
    if(!ctx)
        return false ;
    const Vector & b = assembly->getForces() ;
    if(!AmieB200Shim::upload_diagonal(ctx, diagonal, b.size(), assembly))
        return false ;
    if(x.size() != b.size())
        x.resize(b.size(), 0.) ;
    uint64_t n = 0 ;
    double err = 0, rho = 0 ;
    int ret ;
    if(const std::vector<uint32_t> * perm = AmieB200Shim::permutation_for(assembly))
    {
        // the device holds the renumbered matrix (AMIE_B200_RENUMBER): b, x0 in, x out across the permutation
        const size_t stride = b.size()/perm->size() ;
        Vector bp, x0p, xp(0., b.size()) ;
        AmieB200Shim::to_device_order(*perm, stride, b, bp) ;
        if(x0.size()) AmieB200Shim::to_device_order(*perm, stride, x0, x0p) ;
        ret = amie_b200_pcg(ctx, &bp[0], x0p.size() ? &x0p[0] : nullptr, x0p.size(), kind, eps, maxit, nssor,
                            0, 0, &xp[0], &n, &err, &rho) ;
        if(ret >= 0) AmieB200Shim::from_device_order(*perm, stride, xp, x) ;
    }
    else
        ret = amie_b200_pcg(ctx, &b[0], x0.size() ? &x0[0] : nullptr, x0.size(), kind, eps, maxit, nssor,
                            rowstart, colstart, &x[0], &n, &err, &rho) ;
    nit = n ;
    if(ret == AMIE_B200_ERR_NAN)
    {
        // conjugategradient.cpp:160-165
        assembly->print() ;
        std::cout << "NaN error" << std::endl ;
        exit(0) ;
    }
    if(ret < 0)
    {
        std::cerr << "amie_b200: pcg: " << amie_b200_last_error(ctx) << std::endl ;
        return false ;
    }
    amie_b200_stats st ;
    const bool have_stats = amie_b200_get_stats(ctx, &st) == 0 ;
    if(ret == 1 && have_stats && st.early_return)
    {
        // conjugategradient.cpp:74-78
        std::cerr << "\n CG "<< x.size() << " homogeneous. " << std::abs(b).max() << std::endl ;
        return true ;
    }
    // conjugategradient.cpp:259-264: flops of the iterations over the solve time in microseconds (here: device time)
    if(have_stats && nit)
        std::cerr << "mflops: " << nit*(2.*assembly->getMatrix().array.size()+4.*x.size())/std::max(st.solve_ms*1e3, 1e-32) << std::endl ;
    if(ret)
        std::cerr << "\n CG " << x.size() << " converged after " << nit << " iterations. Error : " << err << ", last rho = " << rho << ", max : "  << x.max() << ", min : "  << x.min() << std::endl ;
    else
        std::cerr << "\n CG " << x.size() << " did not converge after " << nit << " iterations. Error : " << err << ", last rho = " << rho << ", max : "  << x.max() << ", min : "  << x.min() << std::endl ;
    (void)verbose ;
    return ret == 1 ;
}

}
