#define AECF_POOL_T __nv_bfloat16
#define AECF_POOL_DROP true
#include "pool_bwd_inst.inc"
