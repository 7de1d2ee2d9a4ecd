#define AECF_POOL_T __nv_bfloat16
#define AECF_POOL_DROP false
#include "pool_fwd_inst.inc"
