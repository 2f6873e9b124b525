// biconjugategradientstabilized_b200.cpp -- drop-in replacement of
// solvers/biconjugategradientstabilized.cpp (class of solvers/biconjugategradientstabilized.h:19-24).
#include "solvers/biconjugategradientstabilized.h"
#include "amie_b200_shim.h"
#include <iostream>
#include <algorithm>

using namespace Amie ;

BiConjugateGradientStabilized::BiConjugateGradientStabilized(Assembly * a) :LinearSolver(a) { }

bool BiConjugateGradientStabilized::solve(const Vector &x0, Preconditionner * precond, const double epsilon , const int maxit , bool verbose )
{
    const Vector * diagonal = nullptr ;
    int kind = AmieB200Shim::precond_kind(precond, &diagonal) ;
    if(kind == -2) kind = AMIE_B200_PRECOND_JACOBI ;      // AMIE_B200_BLOCK_JACOBI concerns PCG only
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
    x.resize(b.size(), 0.) ;
    uint64_t n = 0 ;
    double err = 0 ;
    int ret ;
    if(const std::vector<uint32_t> * perm = AmieB200Shim::permutation_for(assembly))
    {
        const size_t stride = b.size()/perm->size() ;
        Vector bp, x0p, xp(0., b.size()) ;
        AmieB200Shim::to_device_order(*perm, stride, b, bp) ;
        if(x0.size() == b.size()) AmieB200Shim::to_device_order(*perm, stride, x0, x0p) ;   // any other size is ignored (:21-24)
        ret = amie_b200_bicgstab(ctx, &bp[0], x0p.size() ? &x0p[0] : nullptr, x0p.size(), kind, epsilon, maxit, &xp[0], &n, &err) ;
        if(ret >= 0) AmieB200Shim::from_device_order(*perm, stride, xp, x) ;
    }
    else
        ret = amie_b200_bicgstab(ctx, &b[0], x0.size() ? &x0[0] : nullptr, x0.size(), kind, epsilon, maxit, &x[0], &n, &err) ;
    if(ret < 0)
    {
        std::cerr << "amie_b200: bicgstab: " << amie_b200_last_error(ctx) << std::endl ;
        return false ;
    }
    amie_b200_stats st ;
    const bool have_stats = amie_b200_get_stats(ctx, &st) == 0 ;
    if(ret == 1 && have_stats && st.early_return)
        return true ;               // the reference returns from :43-44 / :57-63 without a cerr line
    // biconjugategradientstabilized.cpp:131 (device time instead of the host's wall clock)
    if(have_stats)
        std::cerr << "mflops: " << n*2*(2.*assembly->getMatrix().array.size()+6.*x.size())/std::max(st.solve_ms*1e3, 1e-32) << std::endl ;
    if(verbose)
    {
        if(ret)
            std::cerr << "\n BiCGStab " << x.size() << " converged after " << n << " iterations. Error : " << err << ", max : "  << x.max() << ", min : "  << x.min() <<std::endl ;
        else
            std::cerr << "\n BiCGStab " << x.size() << " did not converge after " << n << " iterations. Error : " << err << ", max : "  << x.max() << ", min : "  << x.min() <<std::endl ;
    }
    return ret == 1 ;
}
