#define AECF_POOL_T float
#define AECF_POOL_DROP false
#include "pool_bwd_inst.inc"
