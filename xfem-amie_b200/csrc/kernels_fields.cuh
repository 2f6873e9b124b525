// kernels_fields.cuh -- the kernels of the per-element field recovery (fields.cu holds the description, the host side
// and the C-ABI).  Depends on device_utils.cuh only (see there).
#pragma once
#include "device_utils.cuh"

#ifndef NO_NODE
#define NO_NODE 0xFFFFFFFFu
#endif
#define FIELD_THREADS 128

// in[e*K + k] -> out[k*n + e]  (once per topology)
template<typename T>
static __global__ void k_to_component_major(const T * __restrict__ in, T * __restrict__ out, uint64_t n, int K)
{
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    const uint64_t total = n*(uint64_t)K ;
    for(uint64_t i = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; i < total ; i += stride)
    {
        const uint64_t e = i/K ;
        const int k = (int)(i-e*K) ;
        out[(uint64_t)k*n+e] = in[i] ;
    }
}

// One thread per element.  g[c][l] accumulates d(shape_j)/d(local_l) * u_j[c] over the element's slots in slot
// order, exactly as the reference's x_xi ... z_zeta accumulators do.
// NPE > 0 (option "fields_variant" = 1, elements of exactly NPE slots): the slot loop is unrolled and split into phases --
// every id and derivative first, then every gathered dof, then the sums in slot order -- so that a thread has all its
// loads in flight at once instead of one dependent id -> dof pair per slot (the plain form stalls on exactly that:
// long_scoreboard only, profiles/r01b_ncu_element_fields.txt).  Same products, same sums, same order -> same bits.
template<int DIM, int NPE = 0>
static __global__ void __launch_bounds__(FIELD_THREADS)
k_element_fields(const uint32_t * __restrict__ ids, const double * __restrict__ dshape, const double * __restrict__ jinv,
                 const double * __restrict__ tensors, const double * __restrict__ istrain, const double * __restrict__ istress,
                 const uint32_t * __restrict__ tensor_of_elem, const double * __restrict__ u, uint64_t n_u,
                 uint64_t n_elem, int npe, double * __restrict__ total_out, double * __restrict__ mech_out,
                 double * __restrict__ stress_out)
{
    constexpr int NC = DIM == 2 ? 3 : 6 ;
    __shared__ double stage[FIELD_THREADS*NC] ;
    const uint64_t ntiles = (n_elem+FIELD_THREADS-1)/FIELD_THREADS ;
    for(uint64_t tile = blockIdx.x ; tile < ntiles ; tile += gridDim.x)
    {
        const uint64_t e0 = tile*FIELD_THREADS ;
        const uint64_t e = e0+threadIdx.x ;
        const bool live = e < n_elem ;
        const int cnt = (int)min((uint64_t)FIELD_THREADS, n_elem-e0) ;
        double t[NC], m[NC], s[NC] ;
        #pragma unroll
        for(int i = 0 ; i < NC ; i++) { t[i] = 0. ; m[i] = 0. ; s[i] = 0. ; }
        if(live)
        {
            double g[DIM][DIM] ;
            #pragma unroll
            for(int c = 0 ; c < DIM ; c++)
                #pragma unroll
                for(int l = 0 ; l < DIM ; l++) g[c][l] = 0. ;
            if constexpr(NPE > 0)
            {
                uint32_t idj[NPE] ;
                double f[NPE][DIM], uj[NPE][DIM] ;
                #pragma unroll
                for(int j = 0 ; j < NPE ; j++) idj[j] = __ldg(ids+(uint64_t)j*n_elem+e) ;
                #pragma unroll
                for(int j = 0 ; j < NPE ; j++)
                    #pragma unroll
                    for(int l = 0 ; l < DIM ; l++) f[j][l] = __ldg(dshape+(uint64_t)(j*DIM+l)*n_elem+e) ;   // (ld_stream is a volatile asm: it would pin the load next to its use)
                #pragma unroll
                for(int j = 0 ; j < NPE ; j++)
                    #pragma unroll
                    for(int c = 0 ; c < DIM ; c++)
                    {
                        const uint64_t k = (uint64_t)idj[j]*DIM+c ;
                        uj[j][c] = (idj[j] != NO_NODE && k < n_u) ? __ldg(u+k) : 0. ;
                    }
                #pragma unroll
                for(int j = 0 ; j < NPE ; j++)
                {
                    const bool used = idj[j] != NO_NODE ;         // selects, not a branch: a branch would sink the loads back
                    #pragma unroll
                    for(int c = 0 ; c < DIM ; c++)
                        #pragma unroll
                        for(int l = 0 ; l < DIM ; l++)
                        {
                            const double sum = __dadd_rn(g[c][l], __dmul_rn(f[j][l], uj[j][c])) ;
                            g[c][l] = used ? sum : g[c][l] ;
                        }
                }
            }
            else
            for(int j = 0 ; j < npe ; j++)
            {
                const uint32_t id = __ldg(ids+(uint64_t)j*n_elem+e) ;
                if(id == NO_NODE) continue ;
                double f[DIM] ;
                #pragma unroll
                for(int l = 0 ; l < DIM ; l++) f[l] = ld_stream(dshape+(uint64_t)(j*DIM+l)*n_elem+e) ;
                #pragma unroll
                for(int c = 0 ; c < DIM ; c++)
                {
                    const uint64_t k = (uint64_t)id*DIM+c ;
                    const double d = k < n_u ? __ldg(u+k) : 0. ;          // ElementState::step, :3641-3648
                    #pragma unroll
                    for(int l = 0 ; l < DIM ; l++) g[c][l] = __dadd_rn(g[c][l], __dmul_rn(f[l], d)) ;
                }
            }
            double J[DIM*DIM] ;
            #pragma unroll
            for(int k = 0 ; k < DIM*DIM ; k++) J[k] = ld_stream(jinv+(uint64_t)k*n_elem+e) ;
            // a[0]*J[r][0] + a[1]*J[r][1] (+ a[2]*J[r][2]), left to right
            auto row = [&](const double * a, int r)
            {
                double v = __dadd_rn(__dmul_rn(a[0], J[r*DIM]), __dmul_rn(a[1], J[r*DIM+1])) ;
                if constexpr(DIM == 3) v = __dadd_rn(v, __dmul_rn(a[2], J[r*DIM+2])) ;
                return v ;
            } ;
            // ... continued with b[0]*J[q][0] + b[1]*J[q][1] (+ b[2]*J[q][2])
            auto row2 = [&](const double * a, int r, const double * b, int q)
            {
                double v = row(a, r) ;
                v = __dadd_rn(v, __dmul_rn(b[0], J[q*DIM])) ;
                v = __dadd_rn(v, __dmul_rn(b[1], J[q*DIM+1])) ;
                if constexpr(DIM == 3) v = __dadd_rn(v, __dmul_rn(b[2], J[q*DIM+2])) ;
                return v ;
            } ;
            if constexpr(DIM == 2)
            {
                t[0] = row(g[0], 0) ;                         // :1015
                t[1] = row(g[1], 1) ;                         // :1016
                t[2] = row2(g[0], 1, g[1], 0) ;               // :1017
            }
            else
            {
                t[0] = row(g[0], 0) ;                         // :1072-1074
                t[1] = row(g[1], 1) ;
                t[2] = row(g[2], 2) ;
                t[3] = row2(g[1], 2, g[2], 1) ;               // :1076-1081
                t[4] = row2(g[0], 2, g[2], 0) ;               // :1083-1088
                t[5] = row2(g[1], 0, g[0], 1) ;               // :1090-1095
            }
            const uint64_t ti = tensor_of_elem ? __ldg(tensor_of_elem+e) : e ;
            #pragma unroll
            for(int i = 0 ; i < NC ; i++) m[i] = __dsub_rn(t[i], __ldg(istrain+ti*NC+i)) ;      // :967-968
            const double * C = tensors+ti*NC*NC ;
            #pragma unroll
            for(int i = 0 ; i < NC ; i++)
            {
                double acc = 0. ;
                #pragma unroll
                for(int k = 0 ; k < NC ; k++) acc = __dadd_rn(acc, __dmul_rn(__ldg(C+i*NC+k), m[k])) ;   // matrixops.h:555
                s[i] = __dsub_rn(acc, __ldg(istress+ti*NC+i)) ;                              // :1392
            }
        }
        // element-major results through shared memory: one contiguous, coalesced store per field and tile
        double * outs[3] = { total_out, mech_out, stress_out } ;
        #pragma unroll
        for(int w = 0 ; w < 3 ; w++)
        {
            __syncthreads() ;
            #pragma unroll
            for(int i = 0 ; i < NC ; i++) stage[threadIdx.x*NC+i] = w == 0 ? t[i] : (w == 1 ? m[i] : s[i]) ;
            __syncthreads() ;
            double * dst = outs[w]+e0*NC ;
            for(int i = threadIdx.x ; i < cnt*NC ; i += FIELD_THREADS) dst[i] = stage[i] ;
        }
    }
}

// toPrincipal(v, composition) per element (elements/integrable_entity.cpp:475-596): PRINCIPAL_REAL_STRESS_FIELD
// (DOUBLE_OFFDIAG = false, :1505-1509) and PRINCIPAL_TOTAL / _MECHANICAL_STRAIN_FIELD (true: engineering shears,
// :1236-1262) from the element-major fields k_element_fields left in HBM.  One thread per element, the reference's
// expressions as written -- this translation unit is compiled with -fmad=false, so nothing is contracted.  2D needs
// sqrt only in its results (same bits as the CPU); 3D goes through pow / atan2 / cos / sin, whose device versions may
// differ from glibc's in the last place.
template<int DIM, bool DOUBLE_OFFDIAG>
static __global__ void k_element_principal(const double * __restrict__ in, double * __restrict__ out, uint64_t n_elem)
{
    constexpr int NC = DIM == 2 ? 3 : 6 ;
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x ;
    for(uint64_t e = (uint64_t)blockIdx.x*blockDim.x+threadIdx.x ; e < n_elem ; e += stride)
    {
        const double * v = in+e*NC ;
        double * ret = out+e*DIM ;
        if constexpr(DIM == 2)
        {
            const double s0 = v[0], s1 = v[1], s2 = v[2] ;
            const double trace = s0 + s1 ;
            const double det = DOUBLE_OFFDIAG ? s0*s1 - 0.25*s2*s2 : s0*s1 - s2*s2 ;
            const double delta = sqrt(trace*trace - 4.*det) ;
            const double angle = DOUBLE_OFFDIAG ? 0.5*atan2(0.5*s2, s0 - s1) : 0.5*atan2(s2, s0 - s1) ;
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
            // makeStressOrStrainMatrix (:430-452)
            const double m00 = v[0], m11 = v[1], m22 = v[2], m02 = v[3], m12 = v[4], m01 = v[5] ;
            double tr = 0 ;
            tr += m00 ; tr += m11 ; tr += m22 ;
            double trmat, detmat, m2mat ;
            if(DOUBLE_OFFDIAG)
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
            const double q = m2mat/3. - trmat*trmat/9. ;
            const double r = (trmat*m2mat - 3.*detmat)/6. - trmat*trmat*trmat/27. ;
            const double d = q*q*q + r*r ;
            const double r0 = pow(r*r - d, 1./6.) ;
            double phi = atan2(sqrt(-1.*d), r)/3. ;
            if(fabs(phi) < 1e-12) phi = 0. ;                          // POINT_TOLERANCE, geometry/geometry_base.h:292
            if(phi < 0.) phi += 3.14159265358979323846 ;
            const double som = r0*cos(phi) ;
            const double dif = r0*sin(phi) ;
            ret[0] = 2.*som - trmat/3. ;
            ret[1] = -som - trmat/3. - dif*sqrt(3.) ;
            ret[2] = -som - trmat/3. + dif*sqrt(3.) ;
        }
    }
}

