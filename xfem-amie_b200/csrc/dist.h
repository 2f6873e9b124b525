// dist.h -- internal interface of the row-partitioned (multi-GPU) layer, dist.cu
#pragma once
#include "context.h"
#include "synth_recipe.h"

struct SpmvCall ;

int  dist_world(const amie_b200_ctx * ctx) ;
int  dist_rank(const amie_b200_ctx * ctx) ;
void dist_destroy(amie_b200_ctx * ctx) ;
int  dist_spmv(amie_b200_ctx * ctx, const SpmvCall & c) ;
int  dist_finalize(amie_b200_ctx * ctx, int kind) ;              // allreduce of st->red_local + scalar step on every rank
int  dist_allreduce_max(amie_b200_ctx * ctx, double * value) ;
int  dist_inverse_diagonal(amie_b200_ctx * ctx) ;
int  dist_host_barrier(amie_b200_ctx * ctx) ;                    // in-process groups: all parts past their host-side allocations
// a GLOBAL rowstart (DOF units) in the local row numbering of this rank: 0 .. N_local
uint64_t dist_local_rowstart(const amie_b200_ctx * ctx, uint64_t rowstart_global) ;

// synth_device.cu
int synth_rows_to_device(amie_b200_ctx * ctx, const SynthRecipe & R, uint64_t row0, uint64_t row1,
                         uint32_t ** rowptr_out, uint32_t ** col_out, double ** vals_out, double * b_dev, uint64_t * nnzb_out) ;
// api.cu: set_structure with an explicit column bound (global columns on a distributed context)
int ctx_set_structure(amie_b200_ctx * ctx, int stride, uint64_t nb, const uint32_t * row_size,
                      const uint32_t * column_index, uint64_t nnzb, uint64_t ncols) ;
