// example_solve.cpp -- the C++ mirror in use: build a synthetic S3-hex system with the library's
// generator, solve it the way Assembly::cgsolve does (CG, then a warm-started BiCGStab), print
// what the reference prints.   usage: example_solve <preset> <n>
#include "amie_b200.hpp"
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char ** argv)
{
    const char * preset = argc > 1 ? argv[1] : "S3-hex" ;
    int n = argc > 2 ? atoi(argv[2]) : 16 ;
    amie_b200_synth * s = amie_b200_synth_create(preset, n, 1) ;
    if(!s) { fprintf(stderr, "unknown preset\n") ; return 2 ; }
    int stride ; uint64_t nb, nnzb ;
    amie_b200_synth_sizes(s, &stride, &nb, &nnzb) ;
    std::valarray<unsigned int> rs(nb), ci(nnzb) ;
    amie_b200_synth_count(s, 0, nb, &rs[0], nullptr) ;
    AmieB200::Assembly K(0) ;
    K.externalForces.resize(nb*stride) ;
    {
        std::vector<double> arr(nnzb*stride*(stride+stride%2)) ;
        amie_b200_synth_fill(s, 0, nb, &ci[0], arr.data(), &K.externalForces[0]) ;
        K.coordinateIndexedMatrix = new AmieB200::CoordinateIndexedSparseMatrix(rs, ci, stride) ;
        for(size_t i = 0 ; i < arr.size() ; i++) K.getMatrix().array[i] = arr[i] ;
    }
    bool ok = K.cgsolve() ;
    AmieB200::ConjugateGradient cg(&K) ;
    cg.nssor = 32 ;
    ok = cg.solve(AmieB200::Vector(0), nullptr, 1e-10, -1) && ok ;
    AmieB200::BiConjugateGradientStabilized bi(&K) ;
    bool okb = bi.solve(cg.x, nullptr, 1e-10, -1) ;
    double sum = 0 ;
    for(size_t i = 0 ; i < cg.x.size() ; i++) sum += std::abs(cg.x[i]) ;
    printf("CG %zu converged=%d nit=%zu err=%.6e rho=%.6e checksum=%.15e | BiCGStab converged=%d nit=%zu\n",
           cg.x.size(), (int)ok, cg.nit, cg.last_error, cg.last_rho, sum, (int)okb, bi.nit) ;
    delete K.coordinateIndexedMatrix ;
    amie_b200_synth_destroy(s) ;
    return ok && okb ? 0 : 1 ;
}
