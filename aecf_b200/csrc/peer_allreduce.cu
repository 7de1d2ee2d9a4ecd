// Sum-all-reduce of the fusion-parameter gradient bucket over NVLink peer memory: one kernel per rank, no NCCL.
//
// The reference has no distributed code; this is the collective of the north_star's item (3).
// The bucket is small (1 051 136 bf16 elements = 2.1 MB at D = 512), so what an all-reduce costs is latency, not
// bandwidth: NCCL takes 46 us for it on 8 B200s (profiles/r1_run17_dp8_study.json).  Here every rank maps every
// peer's bucket (CUDA IPC, set up once by aecf_b200.dp.PeerAllReduce) and runs, in place:
//
//   1. barrier  : "my bucket is final" -- a release store of the call's epoch into every peer's flag block, then an
//                 acquire spin on the own block until every peer has stored it
//   2. reduce   : rank r owns slice r of the bucket: it loads that slice from ALL W buckets (16-byte peer loads),
//                 sums in rank order 0..W-1 in fp32 (so every rank ends with bit-identical results), optionally
//                 divides by W, and stores the result into the same slice of ALL W buckets (posted peer stores)
//   3. barrier  : the last CTA of the rank fences, tells every peer "my slice has landed", waits until every peer
//                 has said so, advances the epoch and leaves; the kernel therefore completes only when the whole
//                 bucket is reduced, and whatever follows it in the stream (or in the CUDA graph) may read it
//
// Per rank that is one bucket's worth of NVLink reads and one of writes, and two flag round trips.  The epoch
// lives in device memory and is advanced by the kernel, so the call can be captured into a CUDA graph.  Flags only
// ever grow and are compared with >=, so a rank that is already in the next call cannot confuse a slower one.
// Every spin has a clock-based bound and traps instead of hanging the GPU.
#include <cuda.h>

#include <cstring>

#include "common.cuh"
#include "grad_tail.cuh"

namespace aecf {

constexpr int PEER_MAX_WORLD = 8;
constexpr int PEER_FLAG_WORDS = 64;                 // u32 per rank: [0, 8) data-ready, [8, 16) slice-landed, 16 epoch, 17 done-count
constexpr long long PEER_SPIN_LIMIT = 4000000000LL; // ~2 s of SM clocks

struct PeerParams {
    void* out[PEER_MAX_WORLD];                      // where the reduced slices go: == data for the in-place all-reduce, every
                                                    // rank's `reduced` buffer for the backward's fused sum (grad_tail.cu)
    void* data[PEER_MAX_WORLD];                     // every rank's bucket as mapped in THIS process
    uint32_t* flags[PEER_MAX_WORLD];                // every rank's flag block
    long long count;                                // elements
    int world, rank, average;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void spin_until_at_least(const uint32_t* p, uint32_t epoch) {
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(p) - epoch) < 0) {
        if (clock64() - t0 > PEER_SPIN_LIMIT) __trap();      // a lost peer must fail, not hang the box
    }
}
// peer memory: plain (coherent at system scope) 16-byte accesses, no read-only / non-coherent path
__device__ __forceinline__ uint4 ld_peer(const void* p) {
    uint4 r;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_peer(void* p, const uint4& v) {
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(512)
peer_allreduce_kernel(const PeerParams p) {
    constexpr int V = Vec<T>::N;
    const int W = p.world;
    uint32_t* mine = p.flags[p.rank];
    pdl_wait();                                              // the bucket's producers (the backward) are complete
    const uint32_t epoch = ld_relaxed_sys(mine + 16) + 1u;   // same value in every CTA: written back by the last CTA only

    // ---- 1. every rank's bucket is final ------------------------------------------------------------------
    if (blockIdx.x == 0 && threadIdx.x < W) {
        __threadfence_system();
        st_release_sys(p.flags[threadIdx.x] + p.rank, epoch);
    }
    if (threadIdx.x < W) spin_until_at_least(mine + threadIdx.x, epoch);
    __syncthreads();

    // ---- 2. reduce my slice from all buckets, store it into all buckets -----------------------------------
    const long long chunks = (p.count + V - 1) / V;                       // 16-byte chunks; the bucket is padded to one
    const long long per = (chunks + W - 1) / W;
    const long long c0 = per * p.rank, c1 = min(chunks, c0 + per);
    const float inv = p.average ? 1.0f / static_cast<float>(W) : 1.0f;
    for (long long c = c0 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; c < c1;
         c += static_cast<long long>(gridDim.x) * blockDim.x) {
        uint4 raw[PEER_MAX_WORLD];
#pragma unroll
        for (int r = 0; r < PEER_MAX_WORLD; ++r)
            if (r < W) raw[r] = ld_peer(static_cast<const char*>(p.data[r]) + c * 16);      // all loads in flight first
        float acc[V];
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = 0.f;
#pragma unroll
        for (int r = 0; r < PEER_MAX_WORLD; ++r) {                        // fixed order: identical bits on every rank
            if (r < W) {
                float f[V];
                Vec<T>::unpack(raw[r], f);
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] += f[v];
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] *= inv;
        const uint4 out = Vec<T>::pack(acc);
#pragma unroll
        for (int r = 0; r < PEER_MAX_WORLD; ++r)
            if (r < W) st_peer(static_cast<char*>(p.out[r]) + c * 16, out);
    }

    // ---- 3. my slice has landed everywhere; leave when everybody's has -------------------------------------
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) last = (atomicAdd(mine + 17, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    if (threadIdx.x < W) {
        __threadfence_system();
        st_release_sys(p.flags[threadIdx.x] + 8 + p.rank, epoch);
        spin_until_at_least(mine + 8 + threadIdx.x, epoch);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mine[17] = 0u;                                                    // done-count for the next call
        __threadfence();
        mine[16] = epoch;                                                 // next call's epoch base
    }
}

int launch_peer_sum(int device, const aecf_dp_desc* dp, long long count, cudaStream_t s) {
    if (!dp || !dp->sums || !dp->reduced || !dp->flags) return AECF_ERR_INVALID;
    if (dp->world < 2 || dp->world > PEER_MAX_WORLD || dp->rank < 0 || dp->rank >= dp->world || count <= 0 || count % 4 != 0)
        return AECF_ERR_INVALID;
    PeerParams p{};
    for (int r = 0; r < dp->world; ++r) {
        if (!dp->sums[r] || !dp->reduced[r] || !dp->flags[r]) return AECF_ERR_INVALID;
        if (!aligned16(dp->sums[r]) || !aligned16(dp->reduced[r]) || !aligned16(dp->flags[r])) return AECF_ERR_ALIGNMENT;
        p.data[r] = dp->sums[r];
        p.out[r] = dp->reduced[r];
        p.flags[r] = static_cast<uint32_t*>(dp->flags[r]);
    }
    p.count = count; p.world = dp->world; p.rank = dp->rank; p.average = dp->average;
    DeviceScope device_scope__(device);
    int rc = device_scope__.rc;
    if (rc != AECF_OK) return rc;
    // few CTAs: the kernel shares the SMs with the dX product's persistent CTAs (it needs no shared memory, so it fits
    // next to them) and what it moves is small (world x 2 MB at D = 512) -- latency, not bandwidth
    const long long per = (count / 4 + dp->world - 1) / dp->world;
    long long blocks = (per + 511) / 512;
    if (blocks > 32) blocks = 32;
    if (blocks < 1) blocks = 1;
    TimedLaunch timed(s, AECF_SITE_GRAD_PEER_SUM);
    AECF_CUDA_OK(launch_pdl(peer_allreduce_kernel<float>, dim3(static_cast<unsigned>(blocks)), dim3(512), 0, s, p));
    count_launch();
    return AECF_OK;
}

}  // namespace aecf

using namespace aecf;

extern "C" {

size_t aecf_peer_flag_bytes(void) { return PEER_FLAG_WORDS * sizeof(uint32_t); }

int aecf_peer_enable_access(int32_t device, int32_t peer_device) {
    if (device == peer_device) return AECF_OK;
    DeviceScope device_scope__(device);
    int rc = device_scope__.rc;
    if (rc != AECF_OK) return rc;
    int can = 0;
    AECF_CUDA_OK(cudaDeviceCanAccessPeer(&can, device, peer_device));
    if (!can) return AECF_ERR_UNSUPPORTED;
    const cudaError_t err = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (err == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return AECF_OK; }
    AECF_CUDA_OK(err);
    return AECF_OK;
}

// CUDA IPC, mapped into the context of THIS rank's device.  (torch's tensor sharing re-opens a handle under the EXPORTING
// device's ordinal, which is right for torch ops on that tensor but leaves a kernel running on this rank's GPU without a
// mapping: r2 run 8/9, illegal address in the first peer access.)  The handle names the whole allocation a pointer lies
// in, so the offset inside it travels with it.
int aecf_peer_export(int32_t device, const void* ptr, void* handle64, int64_t* offset) {
    if (!ptr || !handle64 || !offset) return AECF_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI carries IPC handles as 64 bytes");
    DeviceScope device_scope__(device);
    int rc = device_scope__.rc;
    if (rc != AECF_OK) return rc;
    using RangeFn = CUresult (*)(CUdeviceptr*, size_t*, CUdeviceptr);
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    AECF_CUDA_OK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || sym == nullptr) return AECF_ERR_UNSUPPORTED;
    CUdeviceptr base = 0;
    size_t size = 0;
    if (reinterpret_cast<RangeFn>(sym)(&base, &size, reinterpret_cast<CUdeviceptr>(ptr)) != CUDA_SUCCESS) return AECF_ERR_INVALID;
    cudaIpcMemHandle_t h;
    AECF_CUDA_OK(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
    std::memcpy(handle64, &h, sizeof(h));
    *offset = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(ptr) - base);
    return AECF_OK;
}

int aecf_peer_import(int32_t device, const void* handle64, void** base_out) {
    if (!handle64 || !base_out) return AECF_ERR_INVALID;
    DeviceScope device_scope__(device);
    int rc = device_scope__.rc;
    if (rc != AECF_OK) return rc;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    AECF_CUDA_OK(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
    return AECF_OK;
}

int aecf_peer_allreduce(const aecf_peer_desc* d, void* const* peer_data, void* const* peer_flags, void* stream) {
    if (!d || !peer_data || !peer_flags) return AECF_ERR_INVALID;
    if (d->world < 1 || d->world > PEER_MAX_WORLD || d->rank < 0 || d->rank >= d->world || d->count < 0) return AECF_ERR_INVALID;
    if (d->dtype != AECF_F32 && d->dtype != AECF_BF16) return AECF_ERR_INVALID;
    if (d->count == 0) return AECF_OK;
    if (d->count % (d->dtype == AECF_BF16 ? 8 : 4) != 0) return AECF_ERR_ALIGNMENT;      // whole 16-byte chunks only
    PeerParams p{};
    for (int r = 0; r < d->world; ++r) {
        if (!peer_data[r] || !peer_flags[r]) return AECF_ERR_INVALID;
        if (!aligned16(peer_data[r]) || !aligned16(peer_flags[r])) return AECF_ERR_ALIGNMENT;
        p.data[r] = peer_data[r];
        p.out[r] = peer_data[r];
        p.flags[r] = static_cast<uint32_t*>(peer_flags[r]);
    }
    p.count = d->count; p.world = d->world; p.rank = d->rank; p.average = d->average;
    DeviceScope device_scope__(d->device);
    int rc = device_scope__.rc;
    if (rc != AECF_OK) return rc;
    // enough CTAs to keep a bucket's worth of 16-byte peer loads in flight, few enough that W ranks emulated on ONE
    // device (the single-GPU test) are all resident at once
    const int V = d->dtype == AECF_BF16 ? 8 : 4;
    const long long per = ((d->count + V - 1) / V + d->world - 1) / d->world;
    long long blocks = (per + 511) / 512;
    const long long cap = d->grid_limit > 0 ? d->grid_limit : 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TimedLaunch timed(s);
    if (d->dtype == AECF_BF16)
        AECF_CUDA_OK(launch_pdl(peer_allreduce_kernel<__nv_bfloat16>, dim3(static_cast<unsigned>(blocks)), dim3(512), 0, s, p));
    else
        AECF_CUDA_OK(launch_pdl(peer_allreduce_kernel<float>, dim3(static_cast<unsigned>(blocks)), dim3(512), 0, s, p));
    count_launch();
    return AECF_OK;
}

}  // extern "C"
