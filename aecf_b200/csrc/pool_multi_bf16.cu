#define AECF_POOL_T __nv_bfloat16
#include "pool_multi_inst.inc"
