/* amie_oracle_fields.c -- CPU restatement of the reference's per-element field recovery
 * (the step right after the Krylov solve).
 *
 * TEST INFRASTRUCTURE ONLY (see amie_oracle.c).  Never on the product path.
 *
 * Parity status: PINNED.  tests/test_oracle_fields.py checks it bit for bit against
 * tests/golden/AMIE-*-fields.npz: the answers ElementState::getField itself gave inside an unmodified
 * FeatureTree run (TOTAL_STRAIN_FIELD, MECHANICAL_STRAIN_FIELD, REAL_STRESS_FIELD at every element's centre),
 * dumped together with the operands it used by oracle/e2e_harness.cpp (dump_fields) and turned into fixtures
 * by tests/golden/make_golden_fields.py.
 *
 * What is restated:
 *   ElementState::step               elements/integrable_entity.cpp:3607-3667  (gather of the solution per element;
 *                                                                              dof ids beyond the vector read as 0)
 *   getField(TOTAL_STRAIN_FIELD)     elements/integrable_entity.cpp:977-1104   (2D :979-1017, 3D :1019-1095)
 *   getField(MECHANICAL_STRAIN_FIELD) :964-975   (total strain minus the behaviour's imposed strain)
 *   getField(REAL_STRESS_FIELD)      :1379-1392  (tensor * mechanical strain - imposed stress; the product is
 *                                                 matrix_vector_multiply, utilities/matrixops.h:545-558:
 *                                                 std::inner_product from 0, left to right)
 * The shape-function derivatives at the evaluation point (vm->deval(shape function j, XI|ETA|ZETA, p)) and the
 * element's cached inverse Jacobian (ElementState::JinvCache) are INPUTS: they come out of the reference's
 * polynomial virtual machine and geometry, which are outside this path.  Enrichment functions follow the shape
 * functions in the same accumulators (:1002-1011, :1049-1068), so they are simply further slots of an element.
 *
 * Layouts: ids[e*npe + j] (0xFFFFFFFF = unused slot), dshape[(e*npe + j)*dim + d], jinv[(e*dim + a)*dim + b],
 * tensors[(t*nc + i)*nc + k], imposed_strain/imposed_stress[t*nc + i] (NULL = zero),
 * tensor_of_elem[e] (NULL = e), outputs [e*nc + i]; nc = 3 (dim 2) or 6 (dim 3).
 */
#include <stdint.h>
#include <stddef.h>
#include <math.h>

#define NO_NODE 0xFFFFFFFFu

int amie_oracle_element_fields(int dim, uint64_t n_elem, int npe, const uint32_t * ids,
                               const double * dshape, const double * jinv,
                               const double * tensors, const double * imposed_strain, const double * imposed_stress,
                               const uint32_t * tensor_of_elem,
                               const double * u, uint64_t n_u,
                               double * total_strain, double * mechanical_strain, double * real_stress)
{
    if(dim != 2 && dim != 3) return -1 ;
    const int nc = dim == 2 ? 3 : 6 ;
    for(uint64_t e = 0 ; e < n_elem ; e++)
    {
        /* g[c][d] = sum_j d(shape_j)/d(local_d) * u_j[c]  (x_xi, x_eta, ..., z_zeta) */
        double g[3][3] = {{0., 0., 0.}, {0., 0., 0.}, {0., 0., 0.}} ;
        for(int j = 0 ; j < npe ; j++)
        {
            const uint32_t id = ids[e*npe+j] ;
            if(id == NO_NODE) continue ;
            const double * f = dshape+(e*npe+j)*dim ;
            for(int c = 0 ; c < dim ; c++)
            {
                const uint64_t k = (uint64_t)id*dim+c ;
                const double d = k < n_u ? u[k] : 0. ;            /* ElementState::step, :3641-3648 */
                for(int l = 0 ; l < dim ; l++)
                    g[c][l] += f[l]*d ;
            }
        }
        const double * J = jinv+e*dim*dim ;
        double t[6] ;
        if(dim == 2)
        {
            const double x_xi = g[0][0], x_eta = g[0][1], y_xi = g[1][0], y_eta = g[1][1] ;
            t[0] = x_xi*J[0] + x_eta*J[1] ;                                       /* :1015 */
            t[1] = y_xi*J[2] + y_eta*J[3] ;                                       /* :1016 */
            t[2] = x_xi*J[2] + x_eta*J[3] + y_xi*J[0] + y_eta*J[1] ;              /* :1017 */
        }
        else
        {
            const double * x = g[0], * y = g[1], * z = g[2] ;
            t[0] = x[0]*J[0] + x[1]*J[1] + x[2]*J[2] ;                            /* :1072 */
            t[1] = y[0]*J[3] + y[1]*J[4] + y[2]*J[5] ;
            t[2] = z[0]*J[6] + z[1]*J[7] + z[2]*J[8] ;
            t[3] = y[0]*J[6] + y[1]*J[7] + y[2]*J[8] + z[0]*J[3] + z[1]*J[4] + z[2]*J[5] ;   /* :1076-1081 */
            t[4] = x[0]*J[6] + x[1]*J[7] + x[2]*J[8] + z[0]*J[0] + z[1]*J[1] + z[2]*J[2] ;   /* :1083-1088 */
            t[5] = y[0]*J[0] + y[1]*J[1] + y[2]*J[2] + x[0]*J[3] + x[1]*J[4] + x[2]*J[5] ;   /* :1090-1095 */
        }
        const uint64_t ti = tensor_of_elem ? tensor_of_elem[e] : e ;
        double m[6] ;
        for(int i = 0 ; i < nc ; i++)
        {
            m[i] = imposed_strain ? t[i]-imposed_strain[ti*nc+i] : t[i] ;         /* :967-968 */
            if(total_strain) total_strain[e*nc+i] = t[i] ;
            if(mechanical_strain) mechanical_strain[e*nc+i] = m[i] ;
        }
        if(real_stress && tensors)
        {
            const double * C = tensors+ti*nc*nc ;
            for(int i = 0 ; i < nc ; i++)
            {
                double s = 0. ;
                for(int k = 0 ; k < nc ; k++)
                    s = s + C[i*nc+k]*m[k] ;                                     /* matrixops.h:555 */
                real_stress[e*nc+i] = imposed_stress ? s-imposed_stress[ti*nc+i] : s ;   /* :1392 */
            }
        }
    }
    return 0 ;
}

/* toPrincipal(stressOrStrain, composition)  (elements/integrable_entity.cpp:475-596): principal values of a stress
 * (SINGLE_OFF_DIAGONAL_VALUES, double_offdiag = 0: getField(PRINCIPAL_REAL_STRESS_FIELD), :1505-1509) or of a strain
 * whose shear components are engineering shears (DOUBLE_OFF_DIAGONAL_VALUES, double_offdiag = 1:
 * PRINCIPAL_TOTAL_STRAIN_FIELD / PRINCIPAL_MECHANICAL_STRAIN_FIELD, :1236-1262).  2D: closed form (:482-498, :539-555);
 * 3D: the trigonometric solution of the characteristic cubic (:508-529, :563-584) on
 * makeStressOrStrainMatrix (:430-452: [0][2] = v[3], [1][2] = v[4], [0][1] = v[5]).
 * in[e*nc + i], out[e*dim + i].  The 3D branch calls pow / atan2 / cos / sin of the C library, like the reference. */
int amie_oracle_principal(int dim, uint64_t n_elem, const double * in, int double_offdiag, double * out)
{
    if(dim != 2 && dim != 3) return -1 ;
    const double POINT_TOLERANCE = 1e-12 ;                                       /* geometry/geometry_base.h:292 */
    for(uint64_t e = 0 ; e < n_elem ; e++)
    {
        if(dim == 2)
        {
            const double * s = in+e*3 ;
            double * ret = out+e*2 ;
            double trace = s[0] + s[1] ;
            double det = double_offdiag ? s[0]*s[1] - 0.25*s[2]*s[2] : s[0]*s[1] - s[2]*s[2] ;
            double delta = sqrt(trace*trace - 4.*det) ;
            double angle = double_offdiag ? 0.5*atan2(0.5*s[2], s[0] - s[1]) : 0.5*atan2(s[2], s[0] - s[1]) ;
            if(cos(angle) < 0)
            {
                ret[0] = (trace + delta)*.5 ;
                ret[1] = (trace - delta)*.5 ;
            }
            else
            {
                ret[0] = (trace - delta)*.5 ;
                ret[1] = (trace + delta)*.5 ;
            }
        }
        else
        {
            const double * v = in+e*6 ;
            double * ret = out+e*3 ;
            const double m00 = v[0], m11 = v[1], m22 = v[2], m02 = v[3], m12 = v[4], m01 = v[5] ;
            double tr = 0 ;
            tr += m00 ; tr += m11 ; tr += m22 ;                                  /* trace(), utilities/matrixops.cpp:213-220 */
            double trmat, detmat, m2mat ;
            if(double_offdiag)
            {
                trmat = -1.*tr ;
                detmat = -2.0*m01*m02*m12*0.125 + m00*m12*m12*0.25 + m11*m02*m02*0.25 + 0.25*m22*m01*m01 - m00*m11*m22 ;
                m2mat = (m00*m11 + m11*m22 + m22*m00) - 0.25*m02*m02 - 0.25*m01*m01 - 0.25*m12*m12 ;
            }
            else
            {
                trmat = -tr ;
                detmat = -2.0*m01*m02*m12 + m00*m12*m12 + m11*m02*m02 + m22*m01*m01 - m00*m11*m22 ;
                m2mat = (m00*m11 + m11*m22 + m22*m00) - m02*m02 - m01*m01 - m12*m12 ;
            }
            double q = m2mat/3. - trmat*trmat/9. ;
            double r = (trmat*m2mat - 3.*detmat)/6. - trmat*trmat*trmat/27. ;
            double d = q*q*q + r*r ;
            double r0 = pow(r*r - d, 1./6.) ;
            double phi = atan2(sqrt(-1.*d), r)/3. ;
            if(fabs(phi) < POINT_TOLERANCE) phi = 0. ;
            if(phi < 0.) phi += 3.14159265358979323846 ;                        /* M_PI */
            double som = r0*cos(phi) ;
            double dif = r0*sin(phi) ;
            ret[0] = 2.*som - trmat/3. ;
            ret[1] = -som - trmat/3. - dif*sqrt(3.) ;
            ret[2] = -som - trmat/3. + dif*sqrt(3.) ;
        }
    }
    return 0 ;
}
