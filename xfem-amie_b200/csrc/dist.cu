// dist.cu -- row-partitioned execution over the GPUs of one NVSwitch box.
// Placeholder entry points until the halo exchange lands (this round: single device).
#include "launch.cuh"

extern "C" {

int amie_b200_nccl_unique_id(void *) { return AMIE_B200_ERR_UNSUPPORTED ; }
int amie_b200_dist_init(amie_b200_ctx * ctx, int, int, const void *, const uint64_t *)
{
    if(ctx) ctx->set_error("distributed context: not built yet") ;
    return AMIE_B200_ERR_UNSUPPORTED ;
}
int amie_b200_dist_set_structure(amie_b200_ctx * ctx, int, uint64_t, const uint32_t *, const uint32_t *, uint64_t)
{
    if(ctx) ctx->set_error("distributed context: not built yet") ;
    return AMIE_B200_ERR_UNSUPPORTED ;
}
int amie_b200_dist_synth_to_device(amie_b200_ctx * ctx, const amie_b200_synth *)
{
    if(ctx) ctx->set_error("distributed context: not built yet") ;
    return AMIE_B200_ERR_UNSUPPORTED ;
}

}
