#define AECF_POOL_T float
#define AECF_POOL_DROP true
#include "pool_fwd_inst.inc"
