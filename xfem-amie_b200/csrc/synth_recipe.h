// synth_recipe.h -- synthetic structured 2D/3D linear-elastic systems.
//
// One description ("recipe") of a structured mesh problem, and ONE row generator
// (synth_row) that is compiled both for the host (synth.cpp, OpenMP) and for the device
// (synth_device.cu), so the same bits feed the CPU oracle and the GPU solver (every
// multiply-add that matters is an explicit fma(); both TUs are built with contraction off):
//
//   S3-hex-n : unit cube, n^3 nodes lexicographic (i fastest), Q1 hexahedra, 27 blocks/row,
//              E=1 matrix + centred sphere r=0.415 E=10, nu=0.2: the S1 geometry/contrast of the
//              reference's examples/main_3d_benchmark.cpp:213-236, with its BCs (:252-255):
//              u_x=0 on x=0, u_y=0 on y=0, u_z=0 on z=0, unit traction on x=1.
//   S3-tet-n : same grid, every cube split into 6 linear tets (Kuhn) -> 15 blocks/row, like
//              AMIE's linear tetrahedra.
//   S2-tri-n : n^2 nodes, T3 triangles, stride 2, plane stress E=10e9 nu=0.2, u_x=0 left,
//              u_y=0 bottom, imposed u_x on the right (examples/main_tension_benchmark.cpp:119-134).
//   ASR-hex-n: S3-hex with 20 random aggregate spheres E=59e9 in paste E=12e9, nu=0.3
//              (examples/main_3d_asr.cpp:406-410) and a RHS from an imposed eigenstrain in
//              20 small gel zones.
//
// Dirichlet DOFs are eliminated the way Assembly::setBoundaryConditions does it
// (solvers/assembly.cpp:165-253): row and column zeroed, diagonal 1, RHS moved; the stored
// pattern never shrinks, so nnzb is BC-independent.
//
// Blocks are produced in the reference's storage convention (sparse/sparse_matrix.h:129-136):
// sorted block columns, element (r,c) of a block at  c*cl + r  with cl = stride + stride%2.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define SYNTH_HD __host__ __device__ inline
#else
#define SYNTH_HD inline
#endif

#define SYNTH_MAX_TEMPLATES 6
#define SYNTH_MAX_NODES 8
#define SYNTH_MAX_SPHERES 24

enum { SYNTH_S3_HEX = 0, SYNTH_S3_TET = 1, SYNTH_S2_TRI = 2, SYNTH_ASR_HEX = 3 };

struct SynthTemplate
{
    int nn ;                               // nodes of this element type
    int corner[SYNTH_MAX_NODES][3] ;       // cell-corner offset (0/1) of each local node
    // K0[a][b][i][j] : unit-E, unit-h stiffness block between local nodes a and b
    double K0[SYNTH_MAX_NODES][SYNTH_MAX_NODES][3][3] ;
    // g0[a][i] = integral of dN_a/dx_i over the element (unit h): eigenstrain load
    double g0[SYNTH_MAX_NODES][3] ;
} ;

struct SynthRecipe
{
    int kind ;
    int dim ;               // 2 or 3
    int stride ;            // = dim
    int n ;                 // nodes per side
    int ntemplates ;
    double h ;              // 1/(n-1)
    double kscale ;         // h^(dim-2): element stiffness scale
    double nu ;
    double E_matrix ;
    int nspheres ;
    double sphere[SYNTH_MAX_SPHERES][4] ;   // cx cy cz r
    double sphere_E[SYNTH_MAX_SPHERES] ;
    int nzones ;                            // eigenstrain zones (ASR-like)
    double zone[SYNTH_MAX_SPHERES][4] ;
    double zone_force_scale ;               // nodal force = E_cell * zone_force_scale * g0  (= eps0/(1-2nu) * h^(dim-1))
    double face_area ;                      // h^(dim-1)
    double traction ;                       // on x = 1 face, along x
    double imposed_ux ;                     // Dirichlet value on x = 1 (S2-tri), 0 = none
    int fix_right ;                         // 1: u_x imposed on x = 1
    uint32_t stencil_mask ;                 // bit (dz+1)*9+(dy+1)*3+(dx+1) set if offset is a neighbour
    SynthTemplate tpl[SYNTH_MAX_TEMPLATES] ;
} ;

SYNTH_HD uint64_t synth_num_nodes(const SynthRecipe & R)
{
    uint64_t n = (uint64_t)R.n ;
    return R.dim == 3 ? n*n*n : n*n ;
}

SYNTH_HD void synth_node_ijk(const SynthRecipe & R, uint64_t node, int ijk[3])
{
    uint64_t n = (uint64_t)R.n ;
    ijk[0] = (int)(node % n) ;
    ijk[1] = (int)((node / n) % n) ;
    ijk[2] = R.dim == 3 ? (int)(node / (n*n)) : 0 ;
}

// number of stored blocks of the block row of `node` (pattern only)
SYNTH_HD int synth_row_count(const SynthRecipe & R, uint64_t node)
{
    int ijk[3] ;
    synth_node_ijk(R, node, ijk) ;
    int cnt = 0 ;
    int zlo = R.dim == 3 ? -1 : 0, zhi = R.dim == 3 ? 1 : 0 ;
    for(int dz = zlo ; dz <= zhi ; dz++)
        for(int dy = -1 ; dy <= 1 ; dy++)
            for(int dx = -1 ; dx <= 1 ; dx++)
            {
                int code = (dz+1)*9+(dy+1)*3+(dx+1) ;
                if(!((R.stencil_mask >> code) & 1u)) continue ;
                int x = ijk[0]+dx, y = ijk[1]+dy, z = ijk[2]+dz ;
                if(x < 0 || y < 0 || z < 0 || x >= R.n || y >= R.n || (R.dim == 3 && z >= R.n)) continue ;
                cnt++ ;
            }
    return cnt ;
}

// Young modulus of the cell whose origin node is (cx,cy,cz)
SYNTH_HD double synth_cell_E(const SynthRecipe & R, int cx, int cy, int cz)
{
    double px = (cx+0.5)*R.h, py = (cy+0.5)*R.h, pz = R.dim == 3 ? (cz+0.5)*R.h : 0. ;
    double E = R.E_matrix ;
    for(int s = 0 ; s < R.nspheres ; s++)
    {
        double ddx = px-R.sphere[s][0], ddy = py-R.sphere[s][1], ddz = R.dim == 3 ? pz-R.sphere[s][2] : 0. ;
        if(fma(ddx, ddx, fma(ddy, ddy, ddz*ddz)) < R.sphere[s][3]*R.sphere[s][3])
            E = R.sphere_E[s] ;
    }
    return E ;
}

SYNTH_HD int synth_cell_in_zone(const SynthRecipe & R, int cx, int cy, int cz)
{
    double px = (cx+0.5)*R.h, py = (cy+0.5)*R.h, pz = R.dim == 3 ? (cz+0.5)*R.h : 0. ;
    for(int s = 0 ; s < R.nzones ; s++)
    {
        double ddx = px-R.zone[s][0], ddy = py-R.zone[s][1], ddz = R.dim == 3 ? pz-R.zone[s][2] : 0. ;
        if(fma(ddx, ddx, fma(ddy, ddy, ddz*ddz)) < R.zone[s][3]*R.zone[s][3])
            return 1 ;
    }
    return 0 ;
}

// is DOF (node at ijk, component m) a Dirichlet DOF?  value returned in *g
SYNTH_HD int synth_fixed(const SynthRecipe & R, const int ijk[3], int m, double * g)
{
    *g = 0. ;
    if(ijk[m] == 0) return 1 ;                       // symmetry planes: u_m = 0 on x_m = 0
    if(R.fix_right && m == 0 && ijk[0] == R.n-1)      // imposed u_x on x = 1
    {
        *g = R.imposed_ux ;
        return 1 ;
    }
    return 0 ;
}

// Generate the whole block row of `node`.
//   cols[k]      : block column indices, ascending
//   blocks[k*9+..]: block k, element (r,c) at c*3 + r (compact column-major, s x s used)
//   rhs[m]       : right-hand side of the row's DOFs
// returns the number of blocks.
SYNTH_HD int synth_row(const SynthRecipe & R, uint64_t node, uint32_t cols[27], double blocks[27*9], double rhs[3])
{
    const int dim = R.dim ;
    const int n = R.n ;
    int ijk[3] ;
    synth_node_ijk(R, node, ijk) ;

    int slot_of_code[27] ;
    int cnt = 0 ;
    int zlo = dim == 3 ? -1 : 0, zhi = dim == 3 ? 1 : 0 ;
    for(int c = 0 ; c < 27 ; c++) slot_of_code[c] = -1 ;
    for(int dz = zlo ; dz <= zhi ; dz++)
        for(int dy = -1 ; dy <= 1 ; dy++)
            for(int dx = -1 ; dx <= 1 ; dx++)
            {
                int code = (dz+1)*9+(dy+1)*3+(dx+1) ;
                if(!((R.stencil_mask >> code) & 1u)) continue ;
                int x = ijk[0]+dx, y = ijk[1]+dy, z = ijk[2]+dz ;
                if(x < 0 || y < 0 || z < 0 || x >= n || y >= n || (dim == 3 && z >= n)) continue ;
                slot_of_code[code] = cnt ;
                cols[cnt] = (uint32_t)((int64_t)node + dx + (int64_t)n*dy + (int64_t)n*n*dz) ;
                cnt++ ;
            }
    for(int k = 0 ; k < cnt*9 ; k++) blocks[k] = 0. ;
    for(int m = 0 ; m < 3 ; m++) rhs[m] = 0. ;

    // element contributions: all cells touching the node
    for(int oz = (dim == 3 ? 1 : 0) ; oz >= 0 ; oz--)
        for(int oy = 1 ; oy >= 0 ; oy--)
            for(int ox = 1 ; ox >= 0 ; ox--)
            {
                int cx = ijk[0]-ox, cy = ijk[1]-oy, cz = ijk[2]-oz ;    // cell origin
                if(cx < 0 || cy < 0 || cz < 0 || cx >= n-1 || cy >= n-1 || (dim == 3 && cz >= n-1)) continue ;
                const double E = synth_cell_E(R, cx, cy, cz) ;
                const double ke = E*R.kscale ;
                const int zone = R.nzones ? synth_cell_in_zone(R, cx, cy, cz) : 0 ;
                const double fz = zone ? E*R.zone_force_scale : 0. ;
                for(int t = 0 ; t < R.ntemplates ; t++)
                {
                    const SynthTemplate & T = R.tpl[t] ;
                    for(int a = 0 ; a < T.nn ; a++)
                    {
                        if(T.corner[a][0] != ox || T.corner[a][1] != oy || T.corner[a][2] != oz) continue ;
                        if(zone)
                            for(int i = 0 ; i < dim ; i++)
                                rhs[i] = fma(fz, T.g0[a][i], rhs[i]) ;
                        for(int b = 0 ; b < T.nn ; b++)
                        {
                            int dx = T.corner[b][0]-ox, dy = T.corner[b][1]-oy, dz = T.corner[b][2]-oz ;
                            int slot = slot_of_code[(dz+1)*9+(dy+1)*3+(dx+1)] ;
                            double * B = blocks + slot*9 ;
                            for(int i = 0 ; i < dim ; i++)
                                for(int j = 0 ; j < dim ; j++)
                                    B[j*3+i] = fma(ke, T.K0[a][b][i][j], B[j*3+i]) ;
                        }
                    }
                }
            }

    // traction on x = 1: consistent Q1 face load  t * h^(dim-1) * w_j * w_k
    if(R.traction != 0. && ijk[0] == n-1)
    {
        double w = (ijk[1] == 0 || ijk[1] == n-1) ? 0.5 : 1. ;
        if(dim == 3) w *= (ijk[2] == 0 || ijk[2] == n-1) ? 0.5 : 1. ;
        rhs[0] = fma(R.traction*w, R.face_area, rhs[0]) ;
    }

    // Dirichlet elimination (solvers/assembly.cpp:165-253)
    double grow[3] ; int frow[3] ;
    for(int m = 0 ; m < dim ; m++) frow[m] = synth_fixed(R, ijk, m, &grow[m]) ;
    for(int k = 0 ; k < cnt ; k++)
    {
        int qijk[3] ;
        synth_node_ijk(R, cols[k], qijk) ;
        double * B = blocks + k*9 ;
        for(int c = 0 ; c < dim ; c++)
        {
            double g ;
            int fc = synth_fixed(R, qijk, c, &g) ;
            for(int m = 0 ; m < dim ; m++)
            {
                if(frow[m])
                    B[c*3+m] = (cols[k] == node && c == m) ? 1. : 0. ;
                else if(fc)
                {
                    rhs[m] = fma(-g, B[c*3+m], rhs[m]) ;
                    B[c*3+m] = 0. ;
                }
            }
        }
    }
    for(int m = 0 ; m < dim ; m++)
        if(frow[m]) rhs[m] = grow[m] ;
    return cnt ;
}
