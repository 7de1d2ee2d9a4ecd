// Functional stand-ins for the inline-PTX wrappers of aecf_b200/csrc/gemm_tcgen05.cu -- TEST INFRASTRUCTURE ONLY.
//
// Included by that file INSIDE namespace aecf::tc when it is compiled for the host emulation (cuda_emu.h).  Models what
// the kernels rely on, functionally and with no notion of time: mbarriers (arrival counts, transaction bytes, phase
// parity), 2-D tiled TMA loads / stores with the 128-byte shared-memory swizzle, out-of-bounds zero fill and clipping,
// cluster multicast, tensor memory (128 lanes x 512 columns of fp32 per CTA), tcgen05.mma kind::f16 through the
// shared-memory and instruction descriptors (K-major and MN-major operands, cta_group 1 and 2), tcgen05.commit,
// tcgen05.ld 32x32b.  Loads and MMAs complete at once; TMA STORES are deferred to the latest moment the kernel allows --
// a store reads its shared-memory box only when the issuing thread's wait_group[.read] retires its bulk group -- so a
// staging box that is refilled before the wait that protects it produces wrong output, and a kernel that exits with
// stores it never waited for is an error.  What it can catch is therefore logic -- barrier counts and phases, tile / box /
// column indexing, descriptor arithmetic, who arrives where -- and two classes of missing wait (that one, and tcgen05.ld registers used
// before tcgen05.wait::ld: they are poisoned until the wait); not fences, not the other asynchronous hazards.
// Its layouts are validated by the kernels that were measured on hardware producing correct products under it.
#pragma once

#include <cuda.h>

namespace cuda_emu { namespace tc {
uint32_t shared_address(const void* p);
void mbar_init(const void* bar, uint32_t count);
void mbar_arrive(const void* bar, int target_cta, uint32_t expect_tx_bytes);   // target_cta < 0: the barrier's own CTA
bool mbar_test(const void* bar, uint32_t parity);
void tma_load(const CUtensorMap* map, const void* bar, int bar_cta, void* dst, int c0, int c1, unsigned cta_mask);
void tma_store(const CUtensorMap* map, const void* src, int c0, int c1);   // deferred: see store_wait_read
void store_commit();                                    // cp.async.bulk.commit_group of the calling thread
void store_wait_read(int pending_allowed);              // cp.async.bulk.wait_group[.read] N: older groups read shared memory NOW
void tmem_alloc(uint32_t* slot, uint32_t columns);
void mma_f16(int cta_group, uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate);
void tmem_load_32x32(uint32_t taddr, uint32_t* out32);   // registers are poisoned until tmem_load_wait()
void tmem_load_wait();                                  // tcgen05.wait::ld
CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, cuuint32_t rank, void* ptr, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* element_strides, CUtensorMapInterleave,
                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
} }
