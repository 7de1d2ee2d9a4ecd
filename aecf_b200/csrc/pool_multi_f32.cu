#define AECF_POOL_T float
#include "pool_multi_inst.inc"
