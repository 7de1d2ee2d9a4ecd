#define AECF_POOL_T float
#define AECF_POOL_DROP false
#include "pool_fwd_inst.inc"
