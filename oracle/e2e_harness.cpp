// oracle/e2e_harness.cpp -- TEST INFRASTRUCTURE: an unmodified-FeatureTree driver for the
// end-to-end ("secondary oracle") check of the drop-in translation units.
//
// The same source is linked twice by oracle/build_ref.py:
//   oracle/_ref/amie_e2e_ref   = harness + libAmie.a                      (reference CPU solvers)
//   oracle/_ref/amie_e2e_b200  = harness + host/shim/*.o + libAmie.a + libamie_b200.so
// and both are run on the same problem; the displacement fields they write are compared.
//
//   amie_e2e_* 2d <sampling> <out.bin> [dump.bin]   plain-elastic twin of examples/main_tension_benchmark.cpp:119-134
//   amie_e2e_* 3d <sampling> <out.bin> [dump.bin]   S1 sphere-in-cube of examples/main_3d_benchmark.cpp:184-257 (gridsize 20)
// out.bin  : uint64 n, n doubles (F.getDisplacements())
// dump.bin : the assembled system of the last solve in the reference layout
//            (uint64 stride, nb, nnzb; row_size u32[nb]; column_index u32[nnzb]; array f64; forces f64[N])
// elements.bin (optional 6th argument): the elements FeatureTree::assemble handed to the Assembly, in its order
//            (features/features.cpp:3360-3403): uint64 n_elem, npe, stride; ids u32[n_elem*npe];
//            Ke f64[n_elem*npe*npe*s*s] (block (j,k) column-major: [m*s+n] = getCachedElementaryMatrix()[j][k][n][m]);
//            scales f64[n_elem] (all 1: single layer)
#include "features/features.h"
#include "features/sample.h"
#include "features/sample3d.h"
#include "features/inclusion3d.h"
#include "physics/stiffness.h"
#include "physics/stiffness_with_imposed_deformation.h"
#include "utilities/tensor.h"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace Amie ;

static void write_vec(const char * path, const Vector & v)
{
    FILE * f = fopen(path, "wb") ;
    uint64_t n = v.size() ;
    fwrite(&n, 8, 1, f) ;
    fwrite(&v[0], 8, n, f) ;
    fclose(f) ;
}

static void dump_system(const char * path, Assembly * K)
{
    CoordinateIndexedSparseMatrix & A = K->getMatrix() ;
    FILE * f = fopen(path, "wb") ;
    uint64_t h[3] = { A.stride, A.row_size.size(), A.column_index.size() } ;
    fwrite(h, 8, 3, f) ;
    fwrite(&A.row_size[0], 4, h[1], f) ;
    fwrite(&A.column_index[0], 4, h[2], f) ;
    fwrite(&A.array[0], 8, A.array.size(), f) ;
    fwrite(&K->getForces()[0], 8, K->getForces().size(), f) ;
    fclose(f) ;
}

template<class MESH>
static void dump_elements(const char * path, MESH * mesh, size_t s)
{
    std::vector<uint32_t> ids ;
    std::vector<double> ke ;
    uint64_t n_elem = 0, npe = 0 ;
    for(auto j = mesh->begin() ; j != mesh->end() ; j++)
    {
        if(!(j->getBehaviour() && j->getBehaviour()->type != VOID_BEHAVIOUR)) continue ;
        std::vector<size_t> id = j->getDofIds() ;
        auto & cached = j->getCachedElementaryMatrix() ;
        if(!npe) npe = id.size() ;
        if(id.size() != npe || cached.size() != npe) { fprintf(stderr, "dump_elements: ragged element\n") ; exit(3) ; }
        for(size_t a = 0 ; a < npe ; a++) ids.push_back((uint32_t)id[a]) ;
        for(size_t a = 0 ; a < npe ; a++)
            for(size_t b = 0 ; b < npe ; b++)
                for(size_t m = 0 ; m < s ; m++)
                    for(size_t n = 0 ; n < s ; n++)
                        ke.push_back(cached[a][b][n][m]) ;
        n_elem++ ;
    }
    FILE * f = fopen(path, "wb") ;
    uint64_t h[3] = { n_elem, npe, s } ;
    fwrite(h, 8, 3, f) ;
    fwrite(ids.data(), 4, ids.size(), f) ;
    fwrite(ke.data(), 8, ke.size(), f) ;
    std::vector<double> scales(n_elem, 1.) ;
    fwrite(scales.data(), 8, n_elem, f) ;
    fclose(f) ;
}

int main(int argc, char ** argv)
{
    if(argc < 4) { fprintf(stderr, "usage: %s 2d|3d <sampling> <out.bin> [dump.bin]\n", argv[0]) ; return 2 ; }
    const std::string mode = argv[1] ;
    const int sampling = atoi(argv[2]) ;
    if(mode == "2d")
    {
        RectangularFeature sample(0.2, 0.1, 0., 0.) ;
        sample.setBehaviour(new Stiffness(10e9, 0.2)) ;
        FeatureTree F(&sample) ;
        F.setSamplingNumber(sampling) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_XI, LEFT)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_ETA, BOTTOM)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(SET_ALONG_XI, RIGHT, 1e-5)) ;
        F.step() ;
        write_vec(argv[3], F.getDisplacements()) ;
        if(argc > 4) dump_system(argv[4], F.getAssembly(false)) ;
        if(argc > 5) dump_elements(argv[5], F.get2DMesh(), 2) ;
    }
    else
    {
        const double scale = 100., length = 0.15 ;
        const double size = scale*length, half = size/2 ;
        Sample3D sample(nullptr, size, size, size, half, half, half) ;
        FeatureTree F(&sample, 1, -1, 20) ;           // gridsize >= 5: the default 4 divides to 0 (features/features.cpp:158-160)
        F.setProjectionOnBoundaries(false) ;
        sample.setBehaviour(new Stiffness(Tensor::cauchyGreen(1., 0.2, SPACE_THREE_DIMENSIONAL, PLANE_STRESS, YOUNG_POISSON))) ;
        Vector alpha(0., 6) ;
        Inclusion3D * inc = new Inclusion3D(0.0623*scale, sample.getCenter().getX(), sample.getCenter().getY(), sample.getCenter().getZ()) ;
        inc->setBehaviour(new StiffnessWithImposedStrain(Tensor::cauchyGreen(10., .2, SPACE_THREE_DIMENSIONAL, PLANE_STRESS, YOUNG_POISSON), alpha)) ;
        F.addFeature(&sample, inc) ;
        F.setSamplingNumber(sampling) ;
        F.setMaxIterationsPerStep(2) ;
        F.setDeltaTime(0.001) ;
        F.setElementGenerationMethod(0, true) ;
        F.setOrder(LINEAR) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(SET_STRESS_XI, RIGHT, 1.)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_XI, LEFT)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_ETA, BOTTOM)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_ZETA, BACK)) ;
        F.step() ;
        write_vec(argv[3], F.getDisplacements()) ;
        if(argc > 4) dump_system(argv[4], F.getAssembly(false)) ;
        if(argc > 5) dump_elements(argv[5], F.get3DMesh(), 3) ;
    }
    return 0 ;
}
