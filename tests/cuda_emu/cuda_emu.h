// Host emulation of the CUDA execution model -- TEST INFRASTRUCTURE ONLY.
//
// Lets `-m "not gpu"` tests run the SAME kernel sources as the product (aecf_b200/csrc/*.cu, *.cuh), compiled by g++
// with -DAECF_CUDA_EMU, on the CPU: every CUDA thread of a block is a fiber (ucontext), blocks run one after the other,
// __syncthreads / named barriers / __syncwarp / warp shuffles are cooperative barriers between the fibers.  It checks
// the kernels' indexing, reductions, masking and host-side sequencing against the oracle without a GPU; it says
// nothing about memory-model races, performance, or anything in gemm_tcgen05.cu (TMA / TMEM / tcgen05 are not
// emulated: the tensor-core GEMM reports AECF_ERR_UNSUPPORTED here and the SIMT kernel runs instead).
// Nothing under aecf_b200/ loads the library built from this; only tests/ do (tests/test_emu_*.py).
#pragma once

#include <cuda_runtime_api.h>
#include <vector_functions.h>
#include <vector_types.h>

#include <cuda_bf16.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>

// host_defines.h turns __shared__ into an (ignored) attribute for a host compiler; here a block's shared variables are
// function-local statics: blocks run one at a time, so one instance per kernel instantiation is exactly one block's copy.
// LIMIT: the CTAs of a CLUSTER run together and would share that one instance, so a kernel launched in clusters takes
// such arrays from its dynamic shared memory when built for the emulation (bias_tile of gemm_tcgen05_2sm.inc); dynamic
// shared memory is per CTA here as on the device.
#undef __shared__
#define __shared__ static
#undef __global__
#define __global__
#undef __device__
#define __device__
#undef __host__
#define __host__
#undef __forceinline__
#define __forceinline__ inline
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__

namespace cuda_emu {

struct ThreadState {
    uint3 tid;
    int linear, lane, warp;
};
struct BlockState {
    dim3 grid_dim, block_dim;
    uint3 block_idx;
};

constexpr int MAX_CLUSTER = 2;                          // thread-block clusters: the CTAs of one cluster run together
constexpr size_t SMEM_BYTES = 232448;                   // the 227 KB a CTA can opt into on sm_100

ThreadState& thread();
BlockState& block();
int cta_rank();                                         // %cluster_ctarank of the running thread
int cluster_size();
unsigned char* smem_window(int cta);                    // shared memory of CTA `cta` of the running cluster (1024-byte aligned)
void* dynamic_smem();
void named_barrier(int id, int count, bool wait);      // count == 0: every thread of the block that has not exited
void cluster_barrier();                                 // barrier.cluster.arrive + wait
void warp_barrier();
uint32_t warp_exchange(uint32_t value, int source_lane);
// cp.async (LDGSTS) with the copies DEFERRED until the issuing thread's wait_group retires their group: staged data read
// before its wait is stale here, as it may be on the device
void cp_async_enqueue(void* smem_dst, const void* gmem_src);
void cp_async_commit_group();
void cp_async_wait_group(int pending_allowed);
void yield();                                           // let the other fibers run (inside a wait loop)
void made_progress();                                   // a wait condition changed: the deadlock detector's heartbeat
void reset_block_resources(int cluster);                // tcgen05_emu.cpp: per-cluster mbarrier / tensor-memory state
void run_grid(const std::function<void()>& body, dim3 grid, dim3 block_dim, size_t smem_bytes, int cluster = 1);

template <typename... KArgs>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block_dim, size_t smem, KArgs... args) {
    std::tuple<KArgs...> copy(args...);                  // kernel parameters are passed by value
    run_grid([&] { std::apply(kernel, copy); }, grid, block_dim, smem);
    return cudaSuccess;
}
template <typename... KArgs>
inline cudaError_t launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block_dim, size_t smem, int cluster, KArgs... args) {
    std::tuple<KArgs...> copy(args...);
    run_grid([&] { std::apply(kernel, copy); }, grid, block_dim, smem, cluster);
    return cudaSuccess;
}

inline uint32_t bf16_bits(float f) {                    // round to nearest even, like cvt.rn.bf16.f32
    const __nv_bfloat16 h = __float2bfloat16_rn(f);
    uint16_t u;
    std::memcpy(&u, &h, 2);
    return u;
}

}  // namespace cuda_emu

#define threadIdx (cuda_emu::thread().tid)
#define blockIdx (cuda_emu::block().block_idx)
// the same for every thread of a launch, and member names of cudaLaunchConfig_t: plain variables, not macros
extern dim3 blockDim, gridDim;

// ---- synchronisation and warp primitives ------------------------------------------------------------------
inline void __syncthreads() { cuda_emu::named_barrier(0, 0, true); }
inline void __syncwarp(unsigned = 0xffffffffu) { cuda_emu::warp_barrier(); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}
[[noreturn]] inline void __trap() { std::fprintf(stderr, "cuda_emu: __trap()\n"); std::abort(); }
inline long long clock64() { return 0; }

template <typename T>
inline T emu_shfl(T v, int src) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    uint32_t bits;
    std::memcpy(&bits, &v, 4);
    bits = cuda_emu::warp_exchange(bits, src & 31);
    std::memcpy(&v, &bits, 4);
    return v;
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu_shfl(v, cuda_emu::thread().lane ^ lane_mask); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl(v, src); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
    const int lane = cuda_emu::thread().lane;
    return emu_shfl(v, lane + static_cast<int>(delta) < 32 ? lane + static_cast<int>(delta) : lane);
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    const int lane = cuda_emu::thread().lane;
    return emu_shfl(v, lane >= static_cast<int>(delta) ? lane - static_cast<int>(delta) : lane);
}

// ---- memory and arithmetic intrinsics ----------------------------------------------------------------------
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
inline float __uint2float_rn(uint32_t x) { return static_cast<float>(x); }          // default rounding mode: nearest even
inline float __expf(float x) { return expf(x); }
inline float __frcp_rn(float x) { return 1.0f / x; }
// the sources are compiled with -ffp-contract=off, so these stay separately rounded operations
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline long long min(long long a, int b) { return a < b ? a : b; }
inline long long min(int a, long long b) { return a < b ? a : b; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned old = *p; *p = old + v; return old; }   // fibers: no preemption

// ---- runtime API pieces that cuda_runtime.h only provides under nvcc -----------------------------------------
template <typename K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename K> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 2; return cudaSuccess; }
template <typename... KArgs, typename... Args>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(KArgs...), Args&&... args) {
    int cluster = 1;
    for (unsigned i = 0; i < cfg->numAttrs; ++i)
        if (cfg->attrs[i].id == cudaLaunchAttributeClusterDimension) cluster = static_cast<int>(cfg->attrs[i].val.clusterDim.x);
    return cuda_emu::launch_cluster(kernel, cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, cluster, static_cast<KArgs>(args)...);
}

#include "tcgen05_emu.h"
