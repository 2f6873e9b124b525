// synth.cpp -- host side of the synthetic structured elastic systems.
// Builds the recipe (element tables) and fills reference-layout arrays with OpenMP.
// Host-only: no CUDA here.
#include "synth_recipe.h"
#include "synth.h"
#include "../../include/amie_b200.h"

#include <cstring>
#include <cstdlib>
#include <random>
#include <vector>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// K_ab[i][j] = lam G_ab[i][j] + mu G_ab[j][i] + mu delta_ij tr(G_ab),  G_ab[i][j] = int dN_a/dx_i dN_b/dx_j
void blocks_from_G(SynthTemplate & T, const double G[SYNTH_MAX_NODES][SYNTH_MAX_NODES][3][3], int dim, double lam, double mu)
{
    for(int a = 0 ; a < T.nn ; a++)
        for(int b = 0 ; b < T.nn ; b++)
        {
            double tr = 0 ;
            for(int i = 0 ; i < dim ; i++) tr += G[a][b][i][i] ;
            for(int i = 0 ; i < 3 ; i++)
                for(int j = 0 ; j < 3 ; j++)
                    T.K0[a][b][i][j] = (i < dim && j < dim) ? lam*G[a][b][i][j] + mu*G[a][b][j][i] + (i == j ? mu*tr : 0.) : 0. ;
        }
}

// Q1 element on the unit square / cube, 2-point Gauss per direction (exact here)
void make_q1(SynthTemplate & T, int dim, double lam, double mu)
{
    std::memset(&T, 0, sizeof(T)) ;
    T.nn = 1 << dim ;
    for(int a = 0 ; a < T.nn ; a++)
    {
        T.corner[a][0] = a & 1 ;
        T.corner[a][1] = (a >> 1) & 1 ;
        T.corner[a][2] = dim == 3 ? (a >> 2) & 1 : 0 ;
    }
    static double G[SYNTH_MAX_NODES][SYNTH_MAX_NODES][3][3] ;
    std::memset(G, 0, sizeof(G)) ;
    const double gp[2] = { 0.5-0.5/std::sqrt(3.), 0.5+0.5/std::sqrt(3.) } ;
    const int nz = dim == 3 ? 2 : 1 ;
    const double w = dim == 3 ? 1./8. : 1./4. ;
    for(int qz = 0 ; qz < nz ; qz++)
        for(int qy = 0 ; qy < 2 ; qy++)
            for(int qx = 0 ; qx < 2 ; qx++)
            {
                double p[3] = { gp[qx], gp[qy], dim == 3 ? gp[qz] : 0. } ;
                double dN[SYNTH_MAX_NODES][3] ;
                for(int a = 0 ; a < T.nn ; a++)
                {
                    double f[3], d[3] ;
                    for(int k = 0 ; k < 3 ; k++)
                    {
                        f[k] = T.corner[a][k] ? p[k] : 1.-p[k] ;
                        d[k] = T.corner[a][k] ? 1. : -1. ;
                    }
                    if(dim == 2) { f[2] = 1. ; d[2] = 0. ; }
                    dN[a][0] = d[0]*f[1]*f[2] ;
                    dN[a][1] = f[0]*d[1]*f[2] ;
                    dN[a][2] = dim == 3 ? f[0]*f[1]*d[2] : 0. ;
                    for(int i = 0 ; i < dim ; i++) T.g0[a][i] += w*dN[a][i] ;
                }
                for(int a = 0 ; a < T.nn ; a++)
                    for(int b = 0 ; b < T.nn ; b++)
                        for(int i = 0 ; i < dim ; i++)
                            for(int j = 0 ; j < dim ; j++)
                                G[a][b][i][j] += w*dN[a][i]*dN[b][j] ;
            }
    blocks_from_G(T, G, dim, lam, mu) ;
}

double det3(const double m[3][3])
{
    return m[0][0]*(m[1][1]*m[2][2]-m[1][2]*m[2][1]) - m[0][1]*(m[1][0]*m[2][2]-m[1][2]*m[2][0]) + m[0][2]*(m[1][0]*m[2][1]-m[1][1]*m[2][0]) ;
}

// linear simplex with the given corner offsets (unit cell): constant gradients
void make_simplex(SynthTemplate & T, int dim, const int corners[][3], double lam, double mu)
{
    std::memset(&T, 0, sizeof(T)) ;
    T.nn = dim+1 ;
    for(int a = 0 ; a < T.nn ; a++)
        for(int k = 0 ; k < 3 ; k++) T.corner[a][k] = corners[a][k] ;
    // edge matrix J[k][i] = x_{k+1,i} - x_{0,i} ; grad N_{k+1} = row k of J^{-1}^T ...
    double J[3][3] = { {1,0,0},{0,1,0},{0,0,1} } ;
    for(int k = 0 ; k < dim ; k++)
        for(int i = 0 ; i < dim ; i++)
            J[k][i] = corners[k+1][i]-corners[0][i] ;
    double d = det3(J) ;
    // inverse of J
    double inv[3][3] ;
    inv[0][0] =  (J[1][1]*J[2][2]-J[1][2]*J[2][1])/d ; inv[0][1] = -(J[0][1]*J[2][2]-J[0][2]*J[2][1])/d ; inv[0][2] =  (J[0][1]*J[1][2]-J[0][2]*J[1][1])/d ;
    inv[1][0] = -(J[1][0]*J[2][2]-J[1][2]*J[2][0])/d ; inv[1][1] =  (J[0][0]*J[2][2]-J[0][2]*J[2][0])/d ; inv[1][2] = -(J[0][0]*J[1][2]-J[0][2]*J[1][0])/d ;
    inv[2][0] =  (J[1][0]*J[2][1]-J[1][1]*J[2][0])/d ; inv[2][1] = -(J[0][0]*J[2][1]-J[0][1]*J[2][0])/d ; inv[2][2] =  (J[0][0]*J[1][1]-J[0][1]*J[1][0])/d ;
    // x - x0 = J^T lambda  ->  lambda_k = sum_i inv^T ... : grad lambda_{k+1}[i] = inv[i][k]
    double dN[SYNTH_MAX_NODES][3] = {{0}} ;
    for(int k = 0 ; k < dim ; k++)
        for(int i = 0 ; i < dim ; i++)
        {
            dN[k+1][i] = inv[i][k] ;
            dN[0][i] -= inv[i][k] ;
        }
    double vol = std::fabs(d)/(dim == 3 ? 6. : 2.) ;
    static double G[SYNTH_MAX_NODES][SYNTH_MAX_NODES][3][3] ;
    std::memset(G, 0, sizeof(G)) ;
    for(int a = 0 ; a < T.nn ; a++)
    {
        for(int i = 0 ; i < dim ; i++) T.g0[a][i] = vol*dN[a][i] ;
        for(int b = 0 ; b < T.nn ; b++)
            for(int i = 0 ; i < dim ; i++)
                for(int j = 0 ; j < dim ; j++)
                    G[a][b][i][j] = vol*dN[a][i]*dN[b][j] ;
    }
    blocks_from_G(T, G, dim, lam, mu) ;
}

void stencil_from_templates(SynthRecipe & R)
{
    R.stencil_mask = 0 ;
    for(int t = 0 ; t < R.ntemplates ; t++)
        for(int a = 0 ; a < R.tpl[t].nn ; a++)
            for(int b = 0 ; b < R.tpl[t].nn ; b++)
            {
                int dx = R.tpl[t].corner[b][0]-R.tpl[t].corner[a][0] ;
                int dy = R.tpl[t].corner[b][1]-R.tpl[t].corner[a][1] ;
                int dz = R.tpl[t].corner[b][2]-R.tpl[t].corner[a][2] ;
                R.stencil_mask |= 1u << ((dz+1)*9+(dy+1)*3+(dx+1)) ;
            }
}

}

int synth_build_recipe(SynthRecipe & R, const char * preset, int n, uint64_t seed)
{
    if(!preset || n < 2) return -1 ;
    std::memset(&R, 0, sizeof(R)) ;
    std::string p(preset) ;
    R.n = n ;
    R.h = 1./(n-1) ;
    double E_unit = 1. ;
    if(p == "S3-hex" || p == "S3-tet")
    {
        R.kind = p == "S3-hex" ? SYNTH_S3_HEX : SYNTH_S3_TET ;
        R.dim = 3 ;
        R.nu = 0.2 ;
        R.E_matrix = 1. ;
        R.nspheres = 1 ;
        R.sphere[0][0] = R.sphere[0][1] = R.sphere[0][2] = 0.5 ;
        R.sphere[0][3] = 0.415 ;
        R.sphere_E[0] = 10. ;
        R.traction = 1. ;
    }
    else if(p == "S2-tri")
    {
        R.kind = SYNTH_S2_TRI ;
        R.dim = 2 ;
        R.nu = 0.2 ;
        R.E_matrix = 10e9 ;
        R.fix_right = 1 ;
        R.imposed_ux = 1e-5 ;
    }
    else if(p == "ASR-hex")
    {
        R.kind = SYNTH_ASR_HEX ;
        R.dim = 3 ;
        R.nu = 0.3 ;
        R.E_matrix = 12e9 ;
        std::mt19937 rng((uint32_t)(seed ? seed : 1)) ;
        std::uniform_real_distribution<double> U(0., 1.) ;
        R.nspheres = 20 ;
        for(int s = 0 ; s < R.nspheres ; s++)
        {
            R.sphere[s][3] = 0.06+0.08*U(rng) ;
            for(int k = 0 ; k < 3 ; k++) R.sphere[s][k] = 0.1+0.8*U(rng) ;
            R.sphere_E[s] = 59e9 ;
        }
        R.nzones = 20 ;
        for(int s = 0 ; s < R.nzones ; s++)
        {
            // gel pockets sit inside the aggregates
            int host = s % R.nspheres ;
            for(int k = 0 ; k < 3 ; k++) R.zone[s][k] = R.sphere[host][k]+0.3*R.sphere[host][3]*(U(rng)-0.5) ;
            R.zone[s][3] = std::max(0.35*R.sphere[host][3], 1.01*R.h) ;
        }
    }
    else
        return -1 ;
    (void)E_unit ;
    R.stride = R.dim ;
    R.kscale = R.dim == 3 ? R.h : 1. ;
    R.face_area = R.dim == 3 ? R.h*R.h : R.h ;
    const double eps0 = 1e-3 ;   // imposed volumetric eigenstrain in the zones
    R.zone_force_scale = eps0/(1.-2.*R.nu)*R.face_area ;

    // unit-E Lame constants (plane stress in 2D)
    const double mu = 1./(2.*(1.+R.nu)) ;
    const double lam = R.dim == 3 ? R.nu/((1.+R.nu)*(1.-2.*R.nu)) : R.nu/(1.-R.nu*R.nu) ;

    if(R.kind == SYNTH_S3_HEX || R.kind == SYNTH_ASR_HEX)
    {
        R.ntemplates = 1 ;
        make_q1(R.tpl[0], 3, lam, mu) ;
    }
    else if(R.kind == SYNTH_S3_TET)
    {
        // Kuhn split: one tet per permutation of the axes, path 000 -> e_p0 -> e_p0+e_p1 -> 111
        const int perm[6][3] = { {0,1,2},{0,2,1},{1,0,2},{1,2,0},{2,0,1},{2,1,0} } ;
        R.ntemplates = 6 ;
        for(int t = 0 ; t < 6 ; t++)
        {
            int c[4][3] = { {0,0,0},{0,0,0},{0,0,0},{1,1,1} } ;
            c[1][perm[t][0]] = 1 ;
            c[2][perm[t][0]] = 1 ; c[2][perm[t][1]] = 1 ;
            make_simplex(R.tpl[t], 3, c, lam, mu) ;
        }
    }
    else
    {
        const int c0[3][3] = { {0,0,0},{1,0,0},{1,1,0} } ;
        const int c1[3][3] = { {0,0,0},{1,1,0},{0,1,0} } ;
        R.ntemplates = 2 ;
        make_simplex(R.tpl[0], 2, c0, lam, mu) ;
        make_simplex(R.tpl[1], 2, c1, lam, mu) ;
    }
    stencil_from_templates(R) ;
    return 0 ;
}

struct amie_b200_synth
{
    SynthRecipe R ;
} ;

const SynthRecipe * synth_recipe_of(const amie_b200_synth * s) { return s ? &s->R : nullptr ; }

extern "C" {

amie_b200_synth * amie_b200_synth_create(const char * preset, int n, uint64_t seed)
{
    amie_b200_synth * s = new amie_b200_synth ;
    if(synth_build_recipe(s->R, preset, n, seed) != 0)
    {
        delete s ;
        return nullptr ;
    }
    return s ;
}

void amie_b200_synth_destroy(amie_b200_synth * s) { delete s ; }

int amie_b200_synth_count(const amie_b200_synth * s, uint64_t row0, uint64_t row1, uint32_t * row_size, uint64_t * nnzb_out)
{
    if(!s || row1 < row0 || row1 > synth_num_nodes(s->R)) return AMIE_B200_ERR_ARG ;
    uint64_t total = 0 ;
    #pragma omp parallel for schedule(static) reduction(+:total)
    for(int64_t r = (int64_t)row0 ; r < (int64_t)row1 ; r++)
    {
        int c = synth_row_count(s->R, (uint64_t)r) ;
        if(row_size) row_size[r-row0] = (uint32_t)c ;
        total += (uint64_t)c ;
    }
    if(nnzb_out) *nnzb_out = total ;
    return AMIE_B200_OK ;
}

int amie_b200_synth_sizes(const amie_b200_synth * s, int * stride, uint64_t * nb, uint64_t * nnzb)
{
    if(!s) return AMIE_B200_ERR_ARG ;
    if(stride) *stride = s->R.stride ;
    if(nb) *nb = synth_num_nodes(s->R) ;
    if(nnzb) return amie_b200_synth_count(s, 0, synth_num_nodes(s->R), nullptr, nnzb) ;
    return AMIE_B200_OK ;
}

int amie_b200_synth_fill(const amie_b200_synth * s, uint64_t row0, uint64_t row1,
                         uint32_t * column_index, double * array_padded, double * b)
{
    if(!s || row1 < row0 || row1 > synth_num_nodes(s->R)) return AMIE_B200_ERR_ARG ;
    const SynthRecipe & R = s->R ;
    const int st = R.stride ;
    const int cl = st + st%2 ;
    const uint64_t nrows = row1-row0 ;
    // exclusive prefix of row sizes over the range (two-pass, parallel by chunks)
    std::vector<uint64_t> start(nrows+1) ;
    start[0] = 0 ;
    {
        std::vector<uint32_t> rs(nrows) ;
        #pragma omp parallel for schedule(static)
        for(int64_t r = 0 ; r < (int64_t)nrows ; r++) rs[r] = (uint32_t)synth_row_count(R, row0+(uint64_t)r) ;
        for(uint64_t r = 0 ; r < nrows ; r++) start[r+1] = start[r]+rs[r] ;
    }
    #pragma omp parallel for schedule(dynamic, 1024)
    for(int64_t r = 0 ; r < (int64_t)nrows ; r++)
    {
        uint32_t cols[27] ;
        double blocks[27*9] ;
        double rhs[3] ;
        int cnt = synth_row(R, row0+(uint64_t)r, cols, blocks, rhs) ;
        uint64_t k0 = start[r] ;
        for(int k = 0 ; k < cnt ; k++)
        {
            if(column_index) column_index[k0+k] = cols[k] ;
            if(array_padded)
            {
                double * dst = array_padded + (k0+k)*(uint64_t)(st*cl) ;
                for(int c = 0 ; c < st ; c++)
                    for(int rr = 0 ; rr < cl ; rr++)
                        dst[c*cl+rr] = rr < st ? blocks[k*9+c*3+rr] : 0. ;
            }
        }
        if(b)
            for(int m = 0 ; m < st ; m++) b[(uint64_t)r*st+m] = rhs[m] ;
    }
    return AMIE_B200_OK ;
}

}
