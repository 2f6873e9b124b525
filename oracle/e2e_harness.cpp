// oracle/e2e_harness.cpp -- TEST INFRASTRUCTURE: an unmodified-FeatureTree driver for the
// end-to-end ("secondary oracle") check of the drop-in translation units.
//
// The same source is linked twice by oracle/build_ref.py:
//   oracle/_ref/amie_e2e_ref   = harness + libAmie.a                      (reference CPU solvers)
//   oracle/_ref/amie_e2e_b200  = harness + host/shim/*.o + libAmie.a + libamie_b200.so
// and both are run on the same problem; the displacement fields they write are compared.
//
//   amie_e2e_* 2d <sampling> <out.bin> [dump.bin]   plain-elastic twin of examples/main_tension_benchmark.cpp:119-134
//   amie_e2e_* 3d|3di <sampling> <out.bin> [dump.bin]   S1 sphere-in-cube of examples/main_3d_benchmark.cpp:184-257 (gridsize 20)
//   amie_e2e_* 2dst <sampling> <out.bin>            examples/main_tension_benchmark.cpp --space-time (:96-157) with its default
//                                                   parameters: notched space-time damage sample, first step + 3 load steps;
//                                                   the solver sees rowstart = colstart > 0 (space-time planes)
//   amie_e2e_* tripoint <sampling> <out.bin> [nsteps [max_iterations_per_step]]
//                                                   examples/main_tripoint.cpp:274-559 with its usual arguments `<sampling> 0 3.9 1.2`
//                                                   (no stirrups): the reinforced-concrete half beam in three-point bending,
//                                                   ConcreteBehaviour + VonMises rebars in layers; nsteps load steps of the
//                                                   driver's loop (:115-130), every step re-solving on ONE topology while damage
//                                                   changes the values (BASELINE.json config 4).  out.bin then holds one record
//                                                   per load step.
//   amie_e2e_* asr <sampling> <out.bin> [dump.bin [zones [zone_radius]]]
//                                                   examples/main_3d_asr.cpp:402-426 with n = 1 (one aggregate in a paste cube,
//                                                   BASELINE.json config 5) and its gel pockets as ExpansiveZone3D features
//                                                   (features/expansiveZone3d.cpp:38-125): XFEM enrichment of the tetrahedra a
//                                                   pocket's surface cuts -- extra unknowns on their nodes, hence block rows of
//                                                   irregular length behind the regular ones.  Elastic phases (the example's
//                                                   moduli), pockets at fixed positions inside the aggregate, grown to a radius
//                                                   the mesh resolves (the example grows them step by step from 1e-3).
//   amie_e2e_* check <0|1|2> <out.bin>              examples/main_check_behaviour.cpp:66-125 for the three behaviours whose golden files
//                                                   the reference ships (examples/test/check_behaviour_test_stiffness*_base, 1 % bar):
//                                                   0 test_stiffness.ini (`--steps 1 --maximum-value 1e6 --stress --constant`),
//                                                   1 test_stiffness_with_imposed_deformation.ini, 2 ..._with_imposed_stress.ini (both
//                                                   `--steps 1 --free`): the 2-element, 8-unknown sample.  Prints on stderr the line
//                                                   the example writes to its result file: `check: <time> <strain*1e3> <stress/1e6> <damage%>`.
//   amie_e2e_* xfem 0 <out.bin>                     examples/test/main_test_xfem.cpp:35-76, the reference's own XFEM test (golden file
//                                                   examples/test/test_xfem_base): one ExpansiveZone in a square, its radius grown
//                                                   twice -- the enrichment, hence the topology the solver sees, changes between the
//                                                   steps.  Prints `xfem: <time> <3 strains*1e3> <3 stresses/1e6>` per stage on stderr;
//                                                   out.bin holds one record per stage.
// out.bin  : uint64 n, n doubles (F.getDisplacements())
// dump.bin : the assembled system of the last solve in the reference layout
//            (uint64 stride, nb, nnzb; row_size u32[nb]; column_index u32[nnzb]; array f64; forces f64[N])
// elements.bin (optional 6th argument): the elements FeatureTree::assemble handed to the Assembly, in its order
//            (features/features.cpp:3360-3403): uint64 n_elem, npe, stride; ids u32[n_elem*npe];
//            Ke f64[n_elem*npe*npe*s*s] (block (j,k) column-major: [m*s+n] = getCachedElementaryMatrix()[j][k][n][m]);
//            scales f64[n_elem] (all 1: single layer)
// fields.bin (optional 7th argument): what ElementState::getField computes from the
//            displacement field at each element's centre (local coordinates), together with the operands it used:
//            uint64 n_elem, npe, dim, nc(=3|6); ids u32[n_elem*npe];
//            dshape f64[n_elem*npe*dim] (vm.deval(shape function j, XI|ETA|ZETA, p));
//            jinv f64[n_elem*dim*dim] (row-major, the element's ElementState::JinvCache);
//            tensor f64[n_elem*nc*nc] (row-major, getBehaviour()->getTensor(p)); imposed strain / stress f64[n_elem*nc] each;
//            then the reference's own answers: TOTAL_STRAIN_FIELD, MECHANICAL_STRAIN_FIELD, REAL_STRESS_FIELD f64[n_elem*nc] each,
//            ElementState::getDisplacements() f64[n_elem*npe*dim] (what ElementState::step gathered from the solution),
//            and PRINCIPAL_TOTAL_STRAIN_FIELD, PRINCIPAL_MECHANICAL_STRAIN_FIELD, PRINCIPAL_REAL_STRESS_FIELD f64[n_elem*dim] each
// mode 3di = 3d with a non-zero imposed strain in the inclusion (exercises the imposed-strain/-stress terms)
#include "features/features.h"
#include "features/sample.h"
#include "features/sample3d.h"
#include "features/inclusion3d.h"
#include "features/expansiveZone3d.h"
#include "features/expansiveZone.h"
#include "features/inclusion.h"
#include "physics/stiffness.h"
#include "physics/stiffness_with_imposed_deformation.h"
#include "physics/stiffness_with_imposed_stress.h"
#include "physics/stiffness_and_fracture.h"
#include "physics/void_form.h"
#include "physics/fracturecriteria/vonmises.h"
#include "physics/materials/concrete_behaviour.h"
#include "physics/viscoelasticity_and_fracture.h"
#include "physics/damagemodels/spacetimeisotropiclineardamage.h"
#include "physics/fracturecriteria/maxstrain.h"
#include "utilities/tensor.h"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace Amie ;

static void write_vec(const char * path, const Vector & v, const char * how = "wb")
{
    FILE * f = fopen(path, how) ;
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

struct FieldDump
{
    uint64_t n_elem = 0, npe = 0, dim = 0, nc = 0 ;
    std::vector<uint32_t> ids ;
    std::vector<double> dshape, jinv, tensor, istrain, istress, total, mech, stress, disp ;
    std::vector<double> ptotal, pmech, pstress ;     // PRINCIPAL_TOTAL_STRAIN / _MECHANICAL_STRAIN / _REAL_STRESS, dim values each
} ;

// every element's operands and the reference's own answers at the element centre (local coordinates)
template<class MESH>
static FieldDump collect_fields(MESH * mesh, size_t dim)
{
    FieldDump D ;
    D.dim = dim ;
    const size_t nc = D.nc = dim == 2 ? 3 : 6 ;
    const Point centre = dim == 2 ? Point(1./3., 1./3.) : Point(.25, .25, .25) ;
    const Variable var[3] = { XI, ETA, ZETA } ;
    VirtualMachine vm ;
    for(auto j = mesh->begin() ; j != mesh->end() ; j++)
    {
        if(!(j->getBehaviour() && j->getBehaviour()->type != VOID_BEHAVIOUR)) continue ;
        std::vector<size_t> id = j->getDofIds() ;
        if(!D.npe) D.npe = id.size() ;
        const size_t npe = D.npe ;
        if(id.size() != npe || j->getShapeFunctions().size() != npe || j->getEnrichmentFunctions().size())
        { fprintf(stderr, "collect_fields: ragged or enriched element\n") ; exit(3) ; }
        for(size_t a = 0 ; a < npe ; a++) D.ids.push_back((uint32_t)id[a]) ;
        for(size_t a = 0 ; a < npe ; a++)
            for(size_t d = 0 ; d < dim ; d++)
                D.dshape.push_back(vm.deval(j->getShapeFunction(a), var[d], centre)) ;
        Matrix C = j->getBehaviour()->getTensor(centre, &(*j)) ;
        if(C.numRows() != nc || C.numCols() != nc) { fprintf(stderr, "collect_fields: tensor is not %zux%zu\n", nc, nc) ; exit(3) ; }
        for(size_t a = 0 ; a < nc ; a++)
            for(size_t b = 0 ; b < nc ; b++)
                D.tensor.push_back(C[a][b]) ;
        Vector is(0., nc), it(0., nc) ;
        if(j->getBehaviour()->hasInducedForces())
        {
            is = j->getBehaviour()->getImposedStrain(centre, &(*j)) ;
            it = j->getBehaviour()->getImposedStress(centre, &(*j)) ;
        }
        const Vector & ud = j->getState().getDisplacements() ;
        for(size_t a = 0 ; a < npe*dim ; a++) D.disp.push_back(a < ud.size() ? ud[a] : 0.) ;
        Vector e(0., nc), em(0., nc), sg(0., nc) ;
        j->getState().getField(TOTAL_STRAIN_FIELD, centre, e, true, &vm) ;
        j->getState().getField(MECHANICAL_STRAIN_FIELD, centre, em, true, &vm) ;
        j->getState().getField(REAL_STRESS_FIELD, centre, sg, true, &vm) ;
        Vector pe(0., dim), pm(0., dim), ps(0., dim) ;
        j->getState().getField(PRINCIPAL_TOTAL_STRAIN_FIELD, centre, pe, true, &vm) ;
        j->getState().getField(PRINCIPAL_MECHANICAL_STRAIN_FIELD, centre, pm, true, &vm) ;
        j->getState().getField(PRINCIPAL_REAL_STRESS_FIELD, centre, ps, true, &vm) ;
        for(size_t a = 0 ; a < dim ; a++) { D.ptotal.push_back(pe[a]) ; D.pmech.push_back(pm[a]) ; D.pstress.push_back(ps[a]) ; }
        // the inverse Jacobian getField used: the element's cache (ElementState::JinvCache, filled at the latest by the
        // calls above; it is NOT recomputed per call -- getInverseJacobianMatrix adds the current displacements to the
        // node coordinates, elements/integrable_entity.cpp:657-667, so a fresh evaluation would differ)
        const Matrix & J = *j->getState().JinvCache ;
        for(size_t a = 0 ; a < dim ; a++)
            for(size_t b = 0 ; b < dim ; b++)
                D.jinv.push_back(J[a][b]) ;
        for(size_t a = 0 ; a < nc ; a++)
        {
            D.istrain.push_back(is[a]) ; D.istress.push_back(it[a]) ;
            D.total.push_back(e[a]) ; D.mech.push_back(em[a]) ; D.stress.push_back(sg[a]) ;
        }
        D.n_elem++ ;
    }
    return D ;
}

static void write_fields(const char * path, const FieldDump & D)
{
    FILE * f = fopen(path, "wb") ;
    uint64_t h[4] = { D.n_elem, D.npe, D.dim, D.nc } ;
    fwrite(h, 8, 4, f) ;
    fwrite(D.ids.data(), 4, D.ids.size(), f) ;
    for(const std::vector<double> * v : { &D.dshape, &D.jinv, &D.tensor, &D.istrain, &D.istress, &D.total, &D.mech, &D.stress, &D.disp, &D.ptotal, &D.pmech, &D.pstress })
        fwrite(v->data(), 8, v->size(), f) ;
    fclose(f) ;
}

#ifdef AMIE_B200_E2E
// The drop-in build only: the same fields from the device (amie_b200_element_fields on the context the shim keeps
// for this Assembly), compared bit for bit with what ElementState::getField just answered in this very process.
#include "amie_b200_shim.h"
#include <cstring>
static void check_fields_on_device(const FieldDump & D, Assembly * K, const Vector & u)
{
    amie_b200_ctx * ctx = AmieB200Shim::context_for(K) ;
    if(!ctx) { fprintf(stderr, "fields-on-device: no context\n") ; exit(4) ; }
    int rc = amie_b200_set_element_kinematics(ctx, D.n_elem, (int)D.npe, (int)D.dim, D.ids.data(), D.dshape.data(), D.jinv.data()) ;
    if(!rc) rc = amie_b200_set_element_behaviour(ctx, D.n_elem, D.tensor.data(), D.istrain.data(), D.istress.data(), nullptr) ;
    std::vector<double> e(D.total.size()), em(D.mech.size()), sg(D.stress.size()) ;
    if(!rc) rc = amie_b200_element_fields(ctx, &u[0], u.size(), e.data(), em.data(), sg.data()) ;
    if(rc) { fprintf(stderr, "fields-on-device: %s\n", amie_b200_last_error(ctx)) ; exit(4) ; }
    size_t bad = 0 ;
    for(size_t i = 0 ; i < e.size() ; i++)
        bad += (memcmp(&e[i], &D.total[i], 8) != 0)+(memcmp(&em[i], &D.mech[i], 8) != 0)+(memcmp(&sg[i], &D.stress[i], 8) != 0) ;
    fprintf(stderr, "fields-on-device: %zu elements, %zu values compared with ElementState::getField, %zu differ\n",
            (size_t)D.n_elem, 3*e.size(), bad) ;
}
#endif

int main(int argc, char ** argv)
{
    if(argc < 4) { fprintf(stderr, "usage: %s 2d|3d <sampling> <out.bin> [dump.bin]\n", argv[0]) ; return 2 ; }
    const std::string mode = argv[1] ;
    const int sampling = atoi(argv[2]) ;
    if(mode == "xfem")
    {
        Matrix C = Tensor::cauchyGreen(10e9, 0.2, SPACE_TWO_DIMENSIONAL, PLANE_STRAIN, YOUNG_POISSON) ;
        Vector v(3) ; v[0] = 0.01 ; v[1] = 0.01 ;
        RectangularFeature box(0.1, 0.1, 0, 0) ;
        box.setBehaviour(new Stiffness(C)) ;
        FeatureTree F(&box) ;
        ExpansiveZone * exp = new ExpansiveZone(&box, 0.02, 0, 0, new StiffnessWithImposedStrain(C*2, v)) ;
        F.addFeature(&box, exp) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_XI, BOTTOM_LEFT)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_ETA, BOTTOM)) ;
        F.setSamplingNumber(4) ;
        F.setDeltaTime(1) ;
        remove(argv[3]) ;
        const double radius[3] = { 0.02, 0.03, 0.5 } ;
        for(int stage = 0 ; stage < 3 ; stage++)
        {
            if(stage) exp->setRadius(radius[stage]) ;
            F.step() ;
            F.step() ;
            Vector str = F.getAverageField(TOTAL_STRAIN_FIELD)*1e3 ;
            Vector sig = F.getAverageField(REAL_STRESS_FIELD)/1e6 ;
            fprintf(stderr, "\nxfem: %.10g %.10g %.10g %.10g %.10g %.10g %.10g\n", F.getCurrentTime(), str[0], str[1], str[2], sig[0], sig[1], sig[2]) ;
            write_vec(argv[3], F.getDisplacements(), "ab") ;
        }
    }
    else if(mode == "check")
    {
        const int which = sampling ;              // argv[2]
        Form * behaviour = nullptr ;
        if(which == 0)      behaviour = new Stiffness(10e9, 0.2) ;
        else if(which == 1) behaviour = new StiffnessWithImposedStrain(10e9, 0.2, 0.001) ;
        else                behaviour = new StiffnessWithImposedStress(10e9, 0.2, 1e6) ;
        const bool is_free = which != 0 ;
        const double val = which == 0 ? 1e6 : 0.001 ;             // --maximum-value (default 0.001)
        RectangularFeature sample(0.01, 0.01, 0, 0) ;
        sample.setBehaviour(behaviour) ;
        FeatureTree F(&sample) ;
        F.setSamplingNumber(0) ;
        F.step() ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_ETA, BOTTOM)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_XI, BOTTOM_LEFT)) ;
        // test_stiffness.ini: --stress --constant -> SET_STRESS_ETA on TOP at the full value from the start
        BoundingBoxDefinedBoundaryCondition * up = new BoundingBoxDefinedBoundaryCondition(which == 0 ? SET_STRESS_ETA : SET_ALONG_ETA, TOP, which == 0 ? val : 0.) ;
        if(!is_free) F.addBoundaryCondition(up) ;
        if(is_free) up->setData(val) ;                            // (--steps 1, not --constant: the example sets it, attached or not)
        F.step() ;
        fprintf(stderr, "\ncheck: %.10g %.10g %.10g %.10g\n", F.getCurrentTime(), F.getAverageField(TOTAL_STRAIN_FIELD)[1]*1e3,
                F.getAverageField(REAL_STRESS_FIELD)[1]/1e6, F.getAverageField(SCALAR_DAMAGE_FIELD)[0]*100) ;
        write_vec(argv[3], F.getDisplacements()) ;
    }
    else if(mode == "2dst")
    {
        const double yieldstrain = 0.0005, maxstrain = 0.0001, young = 10e9, radius = 0.01 ;     // the parser defaults (:45-50)
        RectangularFeature sample(nullptr, 0.2, 0.1, 0, 0) ;
        Matrix stiffness = Stiffness(young, 0.2).param ;
        SpaceTimeNonLocalLinearSofteningMaximumStrain * fracST = new SpaceTimeNonLocalLinearSofteningMaximumStrain(maxstrain, maxstrain*young, yieldstrain) ;
        fracST->setMaterialCharacteristicRadius(radius) ;
        SpaceTimeIsotropicLinearDamage * damST = new SpaceTimeIsotropicLinearDamage(1.) ;
        sample.setBehaviour(new ViscoelasticityAndFracture(PURE_ELASTICITY, stiffness, fracST, damST)) ;
        FeatureTree F(&sample) ;
        F.setSamplingNumber(sampling) ;
        F.setMaxIterationsPerStep(20000) ;
        F.setMinDeltaTime(1e-9) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_XI, LEFT_AFTER)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_ETA, BOTTOM_AFTER)) ;
        RectangularFeature * notch = new RectangularFeature(nullptr, 0.002, 0.04, 0., 0.05) ;
        notch->setBehaviour(new VoidForm()) ;
        F.addFeature(&sample, notch) ;
        F.setSamplingFactor(notch, 1.5) ;
        F.step() ;
        F.getAssembly()->setRemoveZeroOnlyLines(false) ;
        BoundingBoxDefinedBoundaryCondition * disp = new BoundingBoxDefinedBoundaryCondition(SET_ALONG_XI, RIGHT_AFTER, 0.) ;
        F.addBoundaryCondition(disp) ;
        for(size_t i = 0 ; i < 3 ; i++)
        {
            disp->setData((i+1)*0.0000001) ;
            F.step() ;
        }
        write_vec(argv[3], F.getDisplacements()) ;
        if(argc > 4) dump_system(argv[4], F.getAssembly(false)) ;
        fprintf(stderr, "2dst: rowstart %zu colstart %zu\n", (size_t)F.getAssembly()->rowstart, (size_t)F.getAssembly()->colstart) ;
    }
    else if(mode == "tripoint")
    {
        // the set-up of examples/main_tripoint.cpp:274-556 for `tripoint <sampling> 0 3.9 1.2`, call by call
        const size_t nsteps = argc > 4 ? (size_t)atoi(argv[4]) : 3 ;
        const int maxiter = argc > 5 ? atoi(argv[5]) : 900 ;
        const double softeningFactor = M_PI*.24 ;
        const double sampleLength = 3.9, sampleHeight = 1.2, supportLever = sampleLength*.5-.250 ;
        const double supportMidPointToEndClearance = 0.25, platewidth = 0.15, plateHeight = 0.051 ;
        const double rebarDiametre = 0.025, rebarEndCover = 0.047, phi = 3.*rebarDiametre ;
        const double compressionCrit = -34.2e6 ;
        const double E_steel = 200e9, nu_steel = 0.01, nu = 0.3, E_paste = 37e9 ;
        const double E_steel_effective = M_PI*0.5*rebarDiametre*rebarDiametre*E_steel/(rebarDiametre*rebarDiametre)*.75 ;
        const double halfSampleOffset = sampleLength*.25 ;
        Matrix m0_steel = Tensor::cauchyGreen(E_steel, nu_steel, SPACE_TWO_DIMENSIONAL, PLANE_STRAIN, YOUNG_POISSON) ;
        Matrix m0_steel_effective = Tensor::cauchyGreen(E_steel_effective, nu_steel, SPACE_TWO_DIMENSIONAL, PLANE_STRAIN, YOUNG_POISSON) ;

        RectangularFeature sample(nullptr, sampleLength*.5, sampleHeight+2.*plateHeight, halfSampleOffset, 0) ;
        RectangularFeature samplebulk(nullptr, sampleLength*.5, sampleHeight+2.*plateHeight, halfSampleOffset, 0) ;
        RectangularFeature topsupport(nullptr, platewidth, plateHeight, platewidth*.5, sampleHeight*.5+plateHeight*.5) ;
        topsupport.setBehaviour(new Stiffness(m0_steel)) ;
        RectangularFeature topsupportbulk(nullptr, platewidth, plateHeight, platewidth*.5, sampleHeight*.5+plateHeight*.5) ;
        topsupportbulk.setBehaviour(new Stiffness(m0_steel)) ;
        RectangularFeature toprightvoid(nullptr, sampleLength*.5-platewidth, plateHeight, (sampleLength*.5-platewidth)*.5+platewidth, sampleHeight*.5+plateHeight*.5) ;
        toprightvoid.setBehaviour(new VoidForm()) ;
        RectangularFeature toprightvoidbulk(nullptr, sampleLength*.5-platewidth, plateHeight, (sampleLength*.5-platewidth)*.5+platewidth, sampleHeight*.5+plateHeight*.5) ;
        toprightvoidbulk.setBehaviour(new VoidForm()) ;
        RectangularFeature baseright(platewidth, plateHeight, supportLever, -sampleHeight*.5-plateHeight*.5) ;
        baseright.setBehaviour(new Stiffness(m0_steel)) ;
        RectangularFeature baserightbulk(platewidth, plateHeight, supportLever, -sampleHeight*.5-plateHeight*.5) ;
        baserightbulk.setBehaviour(new Stiffness(m0_steel)) ;
        RectangularFeature bottomcentervoid(supportLever-platewidth*.5, plateHeight, (supportLever-platewidth*.5)*.5, -sampleHeight*.5-plateHeight*.5) ;
        bottomcentervoid.setBehaviour(new VoidForm()) ;
        bottomcentervoid.isVirtualFeature = true ;
        RectangularFeature rightbottomvoid(supportMidPointToEndClearance-platewidth*.5, plateHeight, sampleLength*.5-(supportMidPointToEndClearance-platewidth*.5)*.5, -sampleHeight*.5-plateHeight*.5) ;
        rightbottomvoid.setBehaviour(new VoidForm()) ;
        rightbottomvoid.isVirtualFeature = true ;
        RectangularFeature bottomcentervoidbulk(supportLever-platewidth*.5, plateHeight, (supportLever-platewidth*.5)*.5, -sampleHeight*.5-plateHeight*.5) ;
        bottomcentervoidbulk.setBehaviour(new VoidForm()) ;
        RectangularFeature rightbottomvoidbulk(supportMidPointToEndClearance-platewidth*.5, plateHeight, sampleLength*.5-(supportMidPointToEndClearance-platewidth*.5)*.5, -sampleHeight*.5-plateHeight*.5) ;
        rightbottomvoidbulk.setBehaviour(new VoidForm()) ;

        const double rebarcenter = (sampleLength*.5-rebarEndCover)*.5, rebarlength = (sampleLength-rebarEndCover*2.)*.5 ;
        const double rebary[4] = { -sampleHeight*.5+0.064, -sampleHeight*.5+0.064+0.085, sampleHeight*.5-0.064, sampleHeight*.5-0.064-0.085 } ;
        std::vector<RectangularFeature *> rebar ;
        for(int i = 0 ; i < 4 ; i++)
        {
            rebar.push_back(new RectangularFeature(&sample, rebarlength, rebarDiametre, rebarcenter, rebary[i])) ;
            rebar.back()->setBehaviour(new StiffnessAndFracture(m0_steel_effective*softeningFactor, new VonMises(490e6))) ;
        }

        FeatureTree F(&samplebulk, .4-3.*rebarDiametre) ;
        for(RectangularFeature * conc : { &samplebulk, &sample })
        {
            conc->setBehaviour(new ConcreteBehaviour(E_paste, nu, compressionCrit, PLANE_STRAIN, UPPER_BOUND, SPACE_TWO_DIMENSIONAL)) ;
            ConcreteBehaviour * cb = dynamic_cast<ConcreteBehaviour *>(conc->getBehaviour()) ;
            cb->variability = 0.00 ;
            cb->rebarLocationsAndDiameters.push_back(std::make_pair(rebar[0]->getCenter().getY(), rebarDiametre)) ;
            cb->rebarLocationsAndDiameters.push_back(std::make_pair(rebar[1]->getCenter().getY(), rebarDiametre)) ;
        }
        samplebulk.getBehaviour()->setSource(sample.getPrimitive()) ;
        sample.isVirtualFeature = true ;

        const int rebarlayer = -1 ;
        F.addFeature(nullptr, &sample, rebarlayer, phi) ;
        F.addFeature(&samplebulk, &baserightbulk) ;
        F.addFeature(&sample, &baseright, rebarlayer, phi) ;
        F.addFeature(&baseright, &bottomcentervoid, rebarlayer, phi) ;
        F.addFeature(&baseright, &rightbottomvoid, rebarlayer, phi) ;
        F.addFeature(&samplebulk, &topsupportbulk) ;
        F.addFeature(&sample, &topsupport, rebarlayer, phi) ;
        F.addFeature(&topsupportbulk, &toprightvoidbulk) ;
        F.addFeature(&topsupport, &toprightvoid, rebarlayer, phi) ;
        F.addFeature(&baserightbulk, &bottomcentervoidbulk) ;
        F.addFeature(&baserightbulk, &rightbottomvoidbulk) ;
        for(int i = 0 ; i < 4 ; i++) F.addFeature(&samplebulk, rebar[i], rebarlayer, phi) ;
        F.setSamplingFactor(rebar[2], 1./20) ;
        F.setSamplingFactor(rebar[3], 1./20) ;
        F.setSamplingFactor(&baseright, 2) ;
        F.setSamplingFactor(&topsupport, 2) ;
        F.setSamplingNumber(sampling) ;
        F.setSamplingRestriction(0) ;
        F.setMaxIterationsPerStep(maxiter) ;
        F.thresholdScoreMet = 0.001 ;
        F.addPoint(new Point(supportLever, -sampleHeight*.5-plateHeight)) ;
        F.addPoint(new Point(platewidth, sampleHeight*.5)) ;
        BoundingBoxAndRestrictionDefinedBoundaryCondition * load = new BoundingBoxAndRestrictionDefinedBoundaryCondition(SET_ALONG_ETA, TOP, -platewidth, platewidth, -10, 10, 0.) ;
        F.addBoundaryCondition(load) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_XI, LEFT)) ;
        F.addBoundaryCondition(new BoundingBoxNearestNodeDefinedBoundaryCondition(FIX_ALONG_ETA, BOTTOM, Point(supportLever, -sampleHeight*.5-plateHeight))) ;
        F.setOrder(LINEAR) ;

        // the driver's loop (:115-130): a load step, and the next displacement increment once it converged
        const double delta_d = argc > 6 ? atof(argv[6]) : 5.*0.0175e-3 ;
        for(size_t v = 0 ; v < nsteps ; v++)
        {
            const bool go_on = F.step() ;
            if(go_on) load->setData(load->getData()-delta_d) ;
            write_vec(argv[3], F.getDisplacements(-1, false), v ? "ab" : "wb") ;
            fprintf(stderr, "tripoint: load step %zu converged %d unknowns %zu average damage %g\n", v, (int)go_on,
                    (size_t)F.getDisplacements(-1, false).size(), F.averageDamage) ;
        }
    }
    else if(mode == "asr")
    {
        const double scale = 100. ;
        const double size = 0.15*scale, half = size/2 ;
        const int nzones = argc > 5 ? atoi(argv[5]) : 6 ;
        const double rz = argc > 6 ? atof(argv[6]) : 1.4 ;
        Sample3D sample(nullptr, size, size, size, half, half, half) ;
        FeatureTree F(&sample, 1, -1, 20) ;
        sample.setBehaviour(new Stiffness(Tensor::cauchyGreen(12e9, 0.3, SPACE_THREE_DIMENSIONAL, PLANE_STRESS, YOUNG_POISSON))) ;
        Inclusion3D * agg = new Inclusion3D(0.0623*scale, half, half, half) ;
        agg->setBehaviour(new Stiffness(Tensor::cauchyGreen(59e9, 0.3, SPACE_THREE_DIMENSIONAL, PLANE_STRESS, YOUNG_POISSON))) ;
        F.addFeature(&sample, agg) ;
        const Matrix gel = Tensor::cauchyGreen(22e9, 0.3, SPACE_THREE_DIMENSIONAL, PLANE_STRESS, YOUNG_POISSON) ;
        Vector swelling(0., 6) ;
        swelling[0] = swelling[1] = swelling[2] = 0.05 ;
        F.setSamplingNumber(sampling) ;
        F.setMaxIterationsPerStep(2) ;
        F.setOrder(LINEAR) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_XI, LEFT)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_ETA, BOTTOM)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_ZETA, BACK)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(SET_STRESS_XI, RIGHT, 1e6)) ;
        // a first step meshes the sample: features without mesh points of their own are dropped from the tree when the
        // features are sampled (features/features.cpp:1811, :2561-2580), so the pockets join the meshed tree afterwards
        // -- which is also when the example's pockets reach a size that cuts elements
        F.step() ;
        // pockets on a fixed pattern around the aggregate's centre, well inside it and apart from one another
        static const double dir[8][3] = { {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}, {.6, .6, .5}, {-.6, -.5, -.6} } ;
        for(int z = 0 ; z < nzones && z < 8 ; z++)
        {
            const double d = 3.4 ;
            ExpansiveZone3D * pocket = new ExpansiveZone3D(nullptr, rz, half+d*dir[z][0], half+d*dir[z][1], half+d*dir[z][2], gel, swelling) ;
            F.addFeature(agg, pocket) ;
        }
        F.step() ;
        write_vec(argv[3], F.getDisplacements()) ;
        if(argc > 4) dump_system(argv[4], F.getAssembly(false)) ;
    }
    else if(mode == "asr2d")
    {
        // the 2D member of the same family (examples/main_asr_simple.cpp:641-704, examples/main_asr.cpp): an aggregate in a
        // paste square with gel pockets as ExpansiveZone features (features/expansiveZone.cpp) -- XFEM enrichment of the
        // triangles a pocket's rim cuts; elastic phases, pockets on a fixed pattern
        const int nzones = argc > 5 ? atoi(argv[5]) : 6 ;
        const double rz = argc > 6 ? atof(argv[6]) : 0.008 ;
        RectangularFeature sample(0.2, 0.2, 0., 0.) ;
        sample.setBehaviour(new Stiffness(12e9, 0.3)) ;
        FeatureTree F(&sample) ;
        Inclusion * agg = new Inclusion(0.06, 0., 0.) ;
        agg->setBehaviour(new Stiffness(59e9, 0.3)) ;
        F.addFeature(&sample, agg) ;
        const Matrix gel = Stiffness(22e9, 0.3).param ;
        Vector swelling(0., 3) ;
        swelling[0] = swelling[1] = 0.05 ;
        F.setSamplingNumber(sampling) ;
        F.setOrder(LINEAR) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_XI, LEFT)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(FIX_ALONG_ETA, BOTTOM)) ;
        F.addBoundaryCondition(new BoundingBoxDefinedBoundaryCondition(SET_STRESS_XI, RIGHT, -5e6)) ;
        const bool after = argc > 7 && atoi(argv[7]) ;
        if(after) F.step() ;
        static const double dir[8][2] = { {1, 0}, {-1, 0}, {0, 1}, {0, -1}, {.7, .7}, {-.7, .7}, {.7, -.7}, {-.7, -.7} } ;
        for(int z = 0 ; z < nzones && z < 8 ; z++)
        {
            const double d = 0.031 ;
            ExpansiveZone * pocket = new ExpansiveZone(nullptr, rz, d*dir[z][0], d*dir[z][1], gel, swelling) ;
            F.addFeature(agg, pocket) ;
        }
        F.step() ;
        write_vec(argv[3], F.getDisplacements()) ;
        if(argc > 4) dump_system(argv[4], F.getAssembly(false)) ;
    }
    else if(mode == "2d")
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
        if(argc > 6)
        {
            FieldDump D = collect_fields(F.get2DMesh(), 2) ;
            write_fields(argv[6], D) ;
#ifdef AMIE_B200_E2E
            check_fields_on_device(D, F.getAssembly(false), F.getDisplacements()) ;
#endif
        }
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
        if(mode == "3di") { alpha[0] = 1e-3 ; alpha[1] = 2e-3 ; alpha[2] = -5e-4 ; }
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
        if(argc > 6)
        {
            FieldDump D = collect_fields(F.get3DMesh(), 3) ;
            write_fields(argv[6], D) ;
#ifdef AMIE_B200_E2E
            check_fields_on_device(D, F.getAssembly(false), F.getDisplacements()) ;
#endif
        }
    }
    return 0 ;
}
