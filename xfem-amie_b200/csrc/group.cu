// group.cu -- amie_b200_create(devices, ndev > 1): several GPUs behind one context and one caller thread.
//
// What the caller sees is the single-device C-ABI: global host arrays in, global host arrays out
// (Assembly::cgsolve, solvers/assembly.cpp:1841-1850, stays one thread in one process).  Inside, the block rows are
// split into contiguous ranges balanced by stored blocks; child context r (device devices[r], worker thread r) holds
// range r and runs the per-rank code of dist.cu: interior SpMV overlapped with NVLink halo pushes, mailbox
// reductions, identical loop decisions on every device.  Host<->device copies of the slices run concurrently, one
// PCIe link per GPU.  See group.h.
#include "group.h"
#include "dist.h"
#include "synth.h"
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <dlfcn.h>

// ------------------------------------------------------------------ CUDA module loading
// The parts of a group wait for one another ON THE DEVICES (halo flags, reduction mailboxes: dist.cu).  With lazy
// module loading (the CUDA 12 default) the first launch of a kernel loads it, and that load can wait for the context
// to go idle -- which never happens while a kernel of another part spins for the very kernel being launched
// (CUDA programming guide, "Lazy Loading": concurrent execution).  One process per GPU never meets this; one process
// driving several parts does, measured: S3-hex-14 on devices {0,0} hangs in the first PCG iterations under LAZY and
// solves under EAGER (profiles/r02_notes.md).  So: a host that announces a multi-device context through
// AMIE_B200_DEVICES gets eager loading before CUDA initialises (this constructor runs when the library is loaded);
// any other host sets CUDA_MODULE_LOADING=EAGER itself, and group_create refuses to build a group under LAZY.
__attribute__((constructor)) static void amie_b200_pick_module_loading()
{
    const char * e = getenv("AMIE_B200_DEVICES") ;
    if(e && strchr(e, ',')) setenv("CUDA_MODULE_LOADING", "EAGER", 0) ;
}

// 1 eager, 2 lazy, 0 unknown (driver older than cuModuleGetLoadingMode)
static int module_loading_mode()
{
    void * h = dlopen("libcuda.so.1", RTLD_NOW) ;
    if(!h) return 0 ;
    typedef int (*fn_t)(int *) ;
    fn_t f = reinterpret_cast<fn_t>(dlsym(h, "cuModuleGetLoadingMode")) ;
    int mode = 0 ;
    if(!f || f(&mode) != 0) return 0 ;
    return mode ;
}

// ------------------------------------------------------------------ barrier + worker pool

bool LocalGroup::barrier()
{
    std::unique_lock<std::mutex> lk(bm) ;
    if(aborted) return false ;
    const uint64_t gen = generation ;
    if(++waiting == world)
    {
        waiting = 0 ;
        generation++ ;
        bcv.notify_all() ;
        return true ;
    }
    bcv.wait(lk, [&] { return generation != gen || aborted ; }) ;
    return !aborted ;
}

void LocalGroup::abort()
{
    std::lock_guard<std::mutex> lk(bm) ;
    aborted = true ;
    broken = true ;
    bcv.notify_all() ;
}

void LocalGroup::worker(int rank)
{
    uint64_t seen = 0 ;
    for( ;; )
    {
        const std::function<int(int)> * f = nullptr ;
        {
            std::unique_lock<std::mutex> lk(jm) ;
            jcv.wait(lk, [&] { return quit || job_gen != seen ; }) ;
            if(quit) return ;
            seen = job_gen ;
            f = job ;
        }
        if(getenv("AMIE_B200_TRACE")) { fprintf(stderr, "[amie_b200 worker %d] job %llu starts\n", rank, (unsigned long long)seen) ; fflush(stderr) ; }
        int rc = (*f)(rank) ;
        if(getenv("AMIE_B200_TRACE")) { fprintf(stderr, "[amie_b200 worker %d] job %llu returned %d\n", rank, (unsigned long long)seen, rc) ; fflush(stderr) ; }
        // a rank that failed may have left the others inside a collective: release them, the group is unusable afterwards
        if(rc < 0 && world > 1) abort() ;
        {
            std::lock_guard<std::mutex> lk(jm) ;
            results[rank] = rc ;
            if(--job_left == 0) dcv.notify_all() ;
        }
    }
}

void LocalGroup::start(int w)
{
    world = w ;
    for(int r = 0 ; r < w ; r++) workers.emplace_back([this, r] { worker(r) ; }) ;
}

void LocalGroup::stop()
{
    {
        std::lock_guard<std::mutex> lk(jm) ;
        quit = true ;
    }
    jcv.notify_all() ;
    for(auto & t : workers) t.join() ;
    workers.clear() ;
}

int LocalGroup::run(const std::function<int(int)> & f)
{
    {
        std::unique_lock<std::mutex> lk(jm) ;
        job = &f ;
        job_left = world ;
        job_gen++ ;
        jcv.notify_all() ;
        dcv.wait(lk, [&] { return job_left == 0 ; }) ;
        job = nullptr ;
    }
    for(int r = 0 ; r < world ; r++) if(results[r] < 0) return results[r] ;
    return results[0] ;
}

// ------------------------------------------------------------------ helpers

namespace {

// first failing child's message becomes the group's
int harvest(amie_b200_ctx * ctx, int rc)
{
    if(rc >= 0) return rc ;
    LocalGroup * g = ctx->group ;
    for(int r = 0 ; r < g->world ; r++)
        if(g->results[r] < 0 && !g->child[r]->err.empty())
        {
            ctx->set_error("device "+std::to_string(g->child[r]->device)+" (part "+std::to_string(r)+"): "+g->child[r]->err) ;
            return rc ;
        }
    ctx->set_error("group: a device failed") ;
    return rc ;
}

int check_usable(amie_b200_ctx * ctx)
{
    if(ctx->group->broken)
    {
        ctx->set_error("group context: an earlier call failed on one device and left the devices out of step; destroy the context") ;
        return AMIE_B200_ERR_STATE ;
    }
    return AMIE_B200_OK ;
}

}

int group_unsupported(amie_b200_ctx * ctx, const char * what)
{
    ctx->set_error(std::string(what)+": not available on a multi-device context") ;
    return AMIE_B200_ERR_UNSUPPORTED ;
}

// ------------------------------------------------------------------ create / destroy

amie_b200_ctx * group_create(const int * devices, int ndev, std::string & err)
{
    if(ndev < 2 || ndev > GROUP_MAX || !devices) { err = "amie_b200_create: 2 to 8 devices per context" ; return nullptr ; }
    int count = 0 ;
    if(cudaGetDeviceCount(&count) != cudaSuccess || count < 1) { err = "amie_b200_create: no CUDA device" ; return nullptr ; }
    for(int i = 0 ; i < ndev ; i++)
        if(devices[i] < 0 || devices[i] >= count) { err = "amie_b200_create: no such CUDA device" ; return nullptr ; }
    if(module_loading_mode() == 2)
    {
        err = "amie_b200_create: a multi-device context needs CUDA_MODULE_LOADING=EAGER in the environment before CUDA "
              "initialises (its parts wait for one another on the devices; lazy kernel loading deadlocks there). "
              "Set it, or list the devices in AMIE_B200_DEVICES before the process starts." ;
        return nullptr ;
    }
    // every pair of distinct devices needs a peer path (NVLink / NVSwitch on the target box): the halo pushes and the
    // reduction mailboxes are plain stores into the neighbour's memory
    for(int i = 0 ; i < ndev ; i++)
        for(int j = 0 ; j < ndev ; j++)
        {
            if(devices[i] == devices[j]) continue ;
            int can = 0 ;
            cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) ;
            if(!can)
            {
                err = "amie_b200_create: devices "+std::to_string(devices[i])+" and "+std::to_string(devices[j])+" have no peer-to-peer path" ;
                return nullptr ;
            }
            cudaSetDevice(devices[i]) ;
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0) ;
            if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            {
                err = std::string("cudaDeviceEnablePeerAccess: ")+cudaGetErrorString(e) ;
                return nullptr ;
            }
            cudaGetLastError() ;
        }
    amie_b200_ctx * ctx = new amie_b200_ctx ;
    LocalGroup * g = new LocalGroup ;
    ctx->group = g ;
    ctx->device = devices[0] ;
    for(int i = 0 ; i < ndev ; i++)
    {
        amie_b200_ctx * c = amie_b200_create(devices+i, 1) ;
        if(!c)
        {
            err = amie_b200_global_error() ;
            for(auto * k : g->child) amie_b200_destroy(k) ;
            delete g ; delete ctx ;
            return nullptr ;
        }
        g->child.push_back(c) ;
    }
    ctx->num_sms = g->child[0]->num_sms ;
    g->start(ndev) ;
    return ctx ;
}

void group_destroy(amie_b200_ctx * ctx)
{
    LocalGroup * g = ctx->group ;
    g->stop() ;
    for(auto * c : g->child) amie_b200_destroy(c) ;
    delete g ;
    ctx->group = nullptr ;
    delete ctx ;
}

// ------------------------------------------------------------------ matrix

int group_set_structure(amie_b200_ctx * ctx, int stride, uint64_t nb, const uint32_t * row_size, const uint32_t * column_index, uint64_t nnzb)
{
    if(!row_size || (!column_index && nnzb)) return AMIE_B200_ERR_ARG ;
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    if(stride != 2 && stride != 3) { ctx->set_error("set_structure: a multi-device context takes stride 2 or 3") ; return AMIE_B200_ERR_UNSUPPORTED ; }
    LocalGroup * g = ctx->group ;
    const double t0 = wall_now() ;
    g->bounds.assign(g->world+1, 0) ;
    if((rc = amie_b200_partition_rows(nb, row_size, g->world, g->bounds.data()))) return rc ;
    g->blk_off.assign(g->world+1, 0) ;
    {
        uint64_t acc = 0 ;
        int p = 0 ;
        for(uint64_t i = 0 ; i <= nb ; i++)
        {
            while(p <= g->world && g->bounds[p] == i) g->blk_off[p++] = acc ;
            if(i < nb) acc += row_size[i] ;
        }
        if(acc != nnzb) { ctx->set_error("set_structure: sum(row_size) != nnzb") ; return AMIE_B200_ERR_ARG ; }
    }
    ctx->have_structure = ctx->have_values = ctx->have_rhs = false ;
    g->block_from.clear() ;                   // a block map belongs to the structure it was given for
    rc = g->run([&](int r) -> int
    {
        amie_b200_ctx * c = g->child[r] ;
        int e = dist_init_local(c, r, g) ;
        if(e) return e ;
        return dist_set_structure_local(c, stride, nb, row_size+g->bounds[r], column_index+g->blk_off[r], g->blk_off[r+1]-g->blk_off[r]) ;
    }) ;
    if(rc) return harvest(ctx, rc) ;
    ctx->S = stride ; ctx->nb = ctx->nb_global = nb ; ctx->nnzb = nnzb ; ctx->N = nb*(uint64_t)stride ;
    ctx->have_structure = true ;
    ctx->stats.structure_ms = (wall_now()-t0)*1e3 ;
    return AMIE_B200_OK ;
}

int group_set_values(amie_b200_ctx * ctx, const double * array)
{
    if(!array && ctx->nnzb) return AMIE_B200_ERR_ARG ;
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    if(!ctx->have_structure) { ctx->set_error("set_values before set_structure") ; return AMIE_B200_ERR_STATE ; }
    LocalGroup * g = ctx->group ;
    const double t0 = wall_now() ;
    const uint64_t per_block = (uint64_t)ctx->S*(ctx->S+ctx->S%2) ;
    if(!g->block_from.empty())
        // under a block map the blocks of a device's rows are scattered over the caller's array: gathered on the host
        rc = g->run([&](int r) -> int { return ctx_set_values_from(g->child[r], array, g->block_from.data()+g->blk_off[r]) ; }) ;
    else
        rc = g->run([&](int r) -> int { return amie_b200_set_values(g->child[r], array+g->blk_off[r]*per_block) ; }) ;
    if(rc) return harvest(ctx, rc) ;
    ctx->have_values = true ;
    ctx->stats.values_ms = (wall_now()-t0)*1e3 ;
    return AMIE_B200_OK ;
}

int group_synth_to_device(amie_b200_ctx * ctx, const amie_b200_synth * s)
{
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    LocalGroup * g = ctx->group ;
    int stride = 0 ;
    uint64_t nb = 0, nnzb = 0 ;
    if((rc = amie_b200_synth_sizes(s, &stride, &nb, &nnzb))) return rc ;
    const double t0 = wall_now() ;
    std::vector<uint32_t> rs(nb) ;
    uint64_t counted = 0 ;
    if((rc = amie_b200_synth_count(s, 0, nb, rs.data(), &counted))) return rc ;
    g->bounds.assign(g->world+1, 0) ;
    if((rc = amie_b200_partition_rows(nb, rs.data(), g->world, g->bounds.data()))) return rc ;
    g->blk_off.assign(g->world+1, 0) ;
    for(int p = 0 ; p < g->world ; p++)
    {
        uint64_t acc = 0 ;
        for(uint64_t i = g->bounds[p] ; i < g->bounds[p+1] ; i++) acc += rs[i] ;
        g->blk_off[p+1] = g->blk_off[p]+acc ;
    }
    rs.clear() ; rs.shrink_to_fit() ;
    rc = g->run([&](int r) -> int
    {
        amie_b200_ctx * c = g->child[r] ;
        int e = dist_init_local(c, r, g) ;
        if(e) return e ;
        return amie_b200_dist_synth_to_device(c, s) ;
    }) ;
    if(rc) return harvest(ctx, rc) ;
    ctx->S = stride ; ctx->nb = ctx->nb_global = nb ; ctx->nnzb = nnzb ; ctx->N = nb*(uint64_t)stride ;
    ctx->have_structure = ctx->have_values = ctx->have_rhs = true ;
    ctx->stats.structure_ms = (wall_now()-t0)*1e3 ;
    return AMIE_B200_OK ;
}

// ------------------------------------------------------------------ per-slice calls

int group_sliced(amie_b200_ctx * ctx, const std::function<int(amie_b200_ctx *, uint64_t, uint64_t)> & f)
{
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    if(!ctx->have_structure) { ctx->set_error("no matrix structure on the devices yet") ; return AMIE_B200_ERR_STATE ; }
    LocalGroup * g = ctx->group ;
    const uint64_t S = (uint64_t)ctx->S ;
    rc = g->run([&](int r) -> int { return f(g->child[r], g->bounds[r]*S, (g->bounds[r+1]-g->bounds[r])*S) ; }) ;
    return harvest(ctx, rc) ;
}

// the four solver entry points.  resident: b / x0 already on the devices, x stays there.
int group_solve(amie_b200_ctx * ctx, bool bicg, bool resident, const double * b, const double * x0, uint64_t nx0,
                int precond_kind, double eps, int maxit, uint64_t nssor, uint64_t rowstart, uint64_t colstart,
                double * x_out, uint64_t * nit_out, double * err_out, double * rho_out)
{
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    if(!ctx->have_values) { ctx->set_error("solve before set_values") ; return AMIE_B200_ERR_STATE ; }
    if(resident && !ctx->have_rhs) { ctx->set_error("solve: rhs not on the devices") ; return AMIE_B200_ERR_STATE ; }
    LocalGroup * g = ctx->group ;
    const uint64_t S = (uint64_t)ctx->S ;
    // BiCGStab only takes x0 when the sizes agree (biconjugategradientstabilized.cpp:21-24): a GLOBAL decision
    if(bicg && nx0 != ctx->N) nx0 = 0 ;
    uint64_t nit[GROUP_MAX] = {} ;
    double err[GROUP_MAX] = {}, rho[GROUP_MAX] = {} ;
    rc = g->run([&](int r) -> int
    {
        amie_b200_ctx * c = g->child[r] ;
        const uint64_t d0 = g->bounds[r]*S, nd = (g->bounds[r+1]-g->bounds[r])*S ;
        int e ;
        if(!resident)
        {
            if((e = amie_b200_upload_rhs(c, b+d0))) return e ;
            const uint64_t n0 = nx0 > d0 ? std::min<uint64_t>(nx0-d0, nd) : 0 ;
            if((e = amie_b200_upload_x0(c, n0 ? x0+d0 : nullptr, n0))) return e ;
        }
        int ret = bicg ? amie_b200_bicgstab_resident(c, precond_kind, eps, maxit, nit+r, err+r)
                       : amie_b200_pcg_resident(c, precond_kind, eps, maxit, nssor, rowstart, colstart, nit+r, err+r, rho+r) ;
        if(ret < 0) return ret ;
        if(!resident && (e = amie_b200_download_x(c, x_out+d0))) return e ;
        return ret ;
    }) ;
    if(rc < 0) return harvest(ctx, rc) ;
    // every device took the same decisions: the answers of part 0 are everybody's
    for(int r = 1 ; r < g->world ; r++)
        if(g->results[r] != g->results[0] || nit[r] != nit[0])
        {
            ctx->set_error("group solve: the devices disagree on the outcome (internal error)") ;
            g->broken = true ;
            return AMIE_B200_ERR_STATE ;
        }
    if(!resident) ctx->have_rhs = true ;
    if(nit_out) *nit_out = nit[0] ;
    if(err_out) *err_out = err[0] ;
    if(rho_out) *rho_out = rho[0] ;
    return rc ;
}

// ------------------------------------------------------------------ stats / options

int group_get_stats(const amie_b200_ctx * ctx, amie_b200_stats * out)
{
    const LocalGroup * g = ctx->group ;
    amie_b200_stats a {} ;
    for(int r = 0 ; r < g->world ; r++)
    {
        amie_b200_stats s ;
        amie_b200_get_stats(g->child[r], &s) ;
        if(r == 0) a = s ;
        else
        {
            // sums: work and bytes; maxima: times (the devices run side by side)
            a.kernel_launches += s.kernel_launches ;
            a.h2d_bytes += s.h2d_bytes ; a.d2h_bytes += s.d2h_bytes ;
            a.spmv_algorithmic_bytes += s.spmv_algorithmic_bytes ;
            a.device_bytes += s.device_bytes ;
            a.solve_ms = std::max(a.solve_ms, s.solve_ms) ;
            a.assemble_ms = std::max(a.assemble_ms, s.assemble_ms) ; a.bc_ms = std::max(a.bc_ms, s.bc_ms) ;
            a.fields_ms = std::max(a.fields_ms, s.fields_ms) ;
            a.field_elements += s.field_elements ;
            a.h2d_ms = std::max(a.h2d_ms, s.h2d_ms) ; a.d2h_ms = std::max(a.d2h_ms, s.d2h_ms) ;
            // per-launch SpMV time: keep the slowest device's average (spmv_ms_total / spmv_timed)
            if(s.spmv_timed && a.spmv_timed && s.spmv_ms_total/s.spmv_timed > a.spmv_ms_total/a.spmv_timed)
            { a.spmv_ms_total = s.spmv_ms_total ; a.spmv_timed = s.spmv_timed ; }
        }
    }
    a.stride = ctx->S ; a.nb = ctx->nb ; a.nnzb = ctx->nnzb ; a.ndof = ctx->N ;
    a.structure_ms = ctx->stats.structure_ms ; a.values_ms = ctx->stats.values_ms ;
    if(ctx->stats.elements_ms > 0.) a.elements_ms = ctx->stats.elements_ms ;
    *out = a ;
    return AMIE_B200_OK ;
}

int group_set_option(amie_b200_ctx * ctx, const char * key, int64_t value)
{
    LocalGroup * g = ctx->group ;
    for(auto * c : g->child)
    {
        int rc = amie_b200_set_option(c, key, value) ;
        if(rc) { ctx->set_error(c->err) ; return rc ; }
    }
    return AMIE_B200_OK ;
}

// ------------------------------------------------------------------ the matrix back on the host (tests)

int group_download_matrix(amie_b200_ctx * ctx, uint32_t * row_size_out, uint32_t * column_index_out, double * array_padded_out)
{
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    if(!ctx->have_structure) { ctx->set_error("download_matrix: no matrix") ; return AMIE_B200_ERR_STATE ; }
    LocalGroup * g = ctx->group ;
    const uint64_t per_block = (uint64_t)ctx->S*(ctx->S+ctx->S%2) ;
    rc = g->run([&](int r) -> int
    {
        amie_b200_ctx * c = g->child[r] ;
        uint32_t * ci = column_index_out ? column_index_out+g->blk_off[r] : nullptr ;
        int e = amie_b200_download_matrix(c, row_size_out ? row_size_out+g->bounds[r] : nullptr, ci,
                                          array_padded_out ? array_padded_out+g->blk_off[r]*per_block : nullptr) ;
        if(e || !ci) return e ;
        // the part's column indices are local (owned, then the halo): back to the caller's numbering
        const uint64_t nhalo = c->ncols_local > c->nb ? c->ncols_local-c->nb : 0 ;
        std::vector<uint32_t> halo(nhalo) ;
        if(nhalo && cudaMemcpy(halo.data(), c->halo_glob, nhalo*sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess)
        { c->set_error("download_matrix: halo list") ; return AMIE_B200_ERR_CUDA ; }
        for(uint64_t k = 0 ; k < c->nnzb ; k++)
            ci[k] = ci[k] < c->nb ? ci[k]+(uint32_t)c->row_base : halo[ci[k]-c->nb] ;
        return AMIE_B200_OK ;
    }) ;
    return harvest(ctx, rc) ;
}

// ------------------------------------------------------------------ value assembly + elimination (assemble.cu per device)

int group_set_elements(amie_b200_ctx * ctx, uint64_t n_elem, int npe, const uint32_t * elem_ids)
{
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    if(!ctx->have_structure) { ctx->set_error("set_elements before set_structure") ; return AMIE_B200_ERR_STATE ; }
    LocalGroup * g = ctx->group ;
    const double t0 = wall_now() ;
    // the whole list on every device: the map build keeps what lands on the device's own block rows
    rc = g->run([&](int r) -> int { return amie_b200_set_elements(g->child[r], n_elem, npe, elem_ids) ; }) ;
    if(rc) return harvest(ctx, rc) ;
    ctx->stats.elements_ms = (wall_now()-t0)*1e3 ;
    return AMIE_B200_OK ;
}

int group_update_elements(amie_b200_ctx * ctx, uint64_t first, uint64_t count, const double * ke, const double * scales)
{
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    LocalGroup * g = ctx->group ;
    // every device reads the same host range over its own PCIe link and places the blocks of its rows
    rc = g->run([&](int r) -> int { return amie_b200_update_elements(g->child[r], first, count, ke, scales) ; }) ;
    return harvest(ctx, rc) ;
}

int group_assemble(amie_b200_ctx * ctx)
{
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    LocalGroup * g = ctx->group ;
    rc = g->run([&](int r) -> int { return amie_b200_assemble(g->child[r]) ; }) ;
    if(rc) return harvest(ctx, rc) ;
    ctx->have_values = true ;
    return AMIE_B200_OK ;
}

int group_set_boundary_conditions(amie_b200_ctx * ctx, uint64_t nfix, const uint32_t * fix_ids, const double * fix_values,
                                  uint64_t nforce, const uint32_t * force_ids, const double * force_values,
                                  const double * add_to_forces, double * natural_inout)
{
    if(!ctx->have_values || !ctx->have_rhs)
    { ctx->set_error("set_boundary_conditions needs the matrix values (set_values / assemble) and the force vector (upload_rhs)") ; return AMIE_B200_ERR_STATE ; }
    // the id lists are global and go to every device (a row also needs the imposed values of the foreign nodes it
    // couples to); the vectors are cut into the devices' rows
    return group_sliced(ctx, [&](amie_b200_ctx * c, uint64_t d0, uint64_t) -> int
    {
        return amie_b200_set_boundary_conditions(c, nfix, fix_ids, fix_values, nforce, force_ids, force_values,
                                                 add_to_forces ? add_to_forces+d0 : nullptr, natural_inout ? natural_inout+d0 : nullptr) ;
    }) ;
}

// ------------------------------------------------------------------ field recovery (fields.cu per device)

int group_set_element_kinematics(amie_b200_ctx * ctx, uint64_t n_elem, int npe, int dim, const uint32_t * elem_ids,
                                 const double * dshape, const double * jinv)
{
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    if(!ctx->have_structure) { ctx->set_error("set_element_kinematics before set_structure") ; return AMIE_B200_ERR_STATE ; }
    LocalGroup * g = ctx->group ;
    g->elem_bounds.assign(g->world+1, 0) ;
    for(int r = 0 ; r <= g->world ; r++) g->elem_bounds[r] = n_elem*(uint64_t)r/(uint64_t)g->world ;
    g->field_dim = dim ; g->field_nc = dim == 2 ? 3 : 6 ;
    rc = g->run([&](int r) -> int
    {
        const uint64_t e0 = g->elem_bounds[r], ne = g->elem_bounds[r+1]-e0 ;
        return amie_b200_set_element_kinematics(g->child[r], ne, npe, dim, elem_ids ? elem_ids+e0*npe : nullptr,
                                                dshape ? dshape+e0*(uint64_t)npe*dim : nullptr, jinv ? jinv+e0*(uint64_t)dim*dim : nullptr) ;
    }) ;
    if(rc) { g->elem_bounds.clear() ; return harvest(ctx, rc) ; }
    return AMIE_B200_OK ;
}

int group_set_element_behaviour(amie_b200_ctx * ctx, uint64_t n_tensors, const double * tensors, const double * imposed_strain,
                                const double * imposed_stress, const uint32_t * tensor_of_elem)
{
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    LocalGroup * g = ctx->group ;
    if(g->elem_bounds.empty()) { ctx->set_error("set_element_behaviour before set_element_kinematics") ; return AMIE_B200_ERR_STATE ; }
    const uint64_t nc = (uint64_t)g->field_nc ;
    if(!tensor_of_elem && n_tensors != g->elem_bounds[g->world])
    { ctx->set_error("set_element_behaviour: without tensor_of_elem there must be one behaviour per element") ; return AMIE_B200_ERR_ARG ; }
    rc = g->run([&](int r) -> int
    {
        const uint64_t e0 = g->elem_bounds[r], ne = g->elem_bounds[r+1]-e0 ;
        if(!ne) return AMIE_B200_OK ;
        // a behaviour table is small and goes to every device whole; one behaviour per element is cut like the elements
        if(tensor_of_elem) return amie_b200_set_element_behaviour(g->child[r], n_tensors, tensors, imposed_strain, imposed_stress, tensor_of_elem+e0) ;
        return amie_b200_set_element_behaviour(g->child[r], ne, tensors+e0*nc*nc, imposed_strain ? imposed_strain+e0*nc : nullptr,
                                               imposed_stress ? imposed_stress+e0*nc : nullptr, nullptr) ;
    }) ;
    return harvest(ctx, rc) ;
}

int group_element_fields(amie_b200_ctx * ctx, const double * u, uint64_t n_u, double * total_strain_out,
                         double * mechanical_strain_out, double * real_stress_out)
{
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    LocalGroup * g = ctx->group ;
    if(g->elem_bounds.empty()) { ctx->set_error("element_fields before set_element_kinematics") ; return AMIE_B200_ERR_STATE ; }
    const uint64_t nc = (uint64_t)g->field_nc, S = (uint64_t)ctx->S ;
    rc = g->run([&](int r) -> int
    {
        amie_b200_ctx * c = g->child[r] ;
        const uint64_t e0 = g->elem_bounds[r] ;
        if(g->elem_bounds[r+1] == e0) return AMIE_B200_OK ;
        double * t = total_strain_out ? total_strain_out+e0*nc : nullptr ;
        double * m = mechanical_strain_out ? mechanical_strain_out+e0*nc : nullptr ;
        double * s = real_stress_out ? real_stress_out+e0*nc : nullptr ;
        if(u) return amie_b200_element_fields(c, u, n_u, t, m, s) ;
        // the resident solution: every part's rows, straight from the devices that hold them (peer copies), into this
        // device's full-length buffer.  The solve that produced them has returned, so the parts' x are complete.
        double * buf = nullptr ;
        int e = fields_u_buffer(c, ctx->N, &buf) ;
        if(e) return e ;
        for(int q = 0 ; q < g->world ; q++)
        {
            const uint64_t d0 = g->bounds[q]*S, nd = (g->bounds[q+1]-g->bounds[q])*S ;
            if(nd && cudaMemcpyAsync(buf+d0, g->child[q]->x, nd*sizeof(double), cudaMemcpyDefault, c->stream) != cudaSuccess)
            { c->set_error(std::string("element_fields: gathering the solution: ")+cudaGetErrorString(cudaGetLastError())) ; return AMIE_B200_ERR_CUDA ; }
        }
        return fields_run(c, buf, ctx->N, 0, t, m, s) ;
    }) ;
    return harvest(ctx, rc) ;
}

int group_element_principal(amie_b200_ctx * ctx, int field, double * principal_out)
{
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    LocalGroup * g = ctx->group ;
    if(g->elem_bounds.empty()) { ctx->set_error("element_principal before element_fields") ; return AMIE_B200_ERR_STATE ; }
    const uint64_t dim = (uint64_t)g->field_dim ;
    rc = g->run([&](int r) -> int
    {
        const uint64_t e0 = g->elem_bounds[r] ;
        if(g->elem_bounds[r+1] == e0) return AMIE_B200_OK ;
        return amie_b200_element_principal(g->child[r], field, principal_out+e0*dim) ;
    }) ;
    return harvest(ctx, rc) ;
}


// ------------------------------------------------------------------ renumbered device matrix (amie_b200_set_block_map)

// block k of the caller's array -> stored block block_to[k] of the structure set_structure was given (the renumbered
// one).  The stored blocks are cut by rows over the devices, so each device needs the INVERSE restricted to its range.
int group_set_block_map(amie_b200_ctx * ctx, const uint32_t * block_to)
{
    int rc = check_usable(ctx) ;
    if(rc) return rc ;
    if(!ctx->have_structure) { ctx->set_error("set_block_map before set_structure") ; return AMIE_B200_ERR_STATE ; }
    LocalGroup * g = ctx->group ;
    g->block_from.clear() ;
    ctx->have_values = false ;                // whatever was uploaded before was in the other order
    for(auto * c : g->child) { c->have_values = false ; c->dinv_valid = false ; }
    if(!block_to) return AMIE_B200_OK ;
    const uint32_t none = 0xFFFFFFFFu ;
    std::vector<uint32_t> from(ctx->nnzb, none) ;
    for(uint64_t k = 0 ; k < ctx->nnzb ; k++)
    {
        if(block_to[k] >= ctx->nnzb || from[block_to[k]] != none) { ctx->set_error("set_block_map: not a permutation of the stored blocks") ; return AMIE_B200_ERR_ARG ; }
        from[block_to[k]] = (uint32_t)k ;
    }
    g->block_from.swap(from) ;
    return AMIE_B200_OK ;
}
