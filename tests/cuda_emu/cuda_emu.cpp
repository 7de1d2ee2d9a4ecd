// Fiber scheduler and CUDA-runtime stand-ins of the host emulation -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
#include "cuda_emu.h"

#include <sys/mman.h>
#include <ucontext.h>

#include <mutex>
#include <vector>

#include "../../include/aecf_b200.h"

dim3 blockDim, gridDim;

namespace cuda_emu {
namespace {

constexpr size_t STACK_BYTES = 1 << 20;
constexpr int MAX_THREADS = 1024, MAX_WARPS = 32, NAMED_BARRIERS = 16;

struct AsyncCopy { void* dst; const void* src; };
struct Fiber {
    ThreadState ts; ucontext_t ctx; bool done; int cta;
    std::vector<AsyncCopy> open_copies;                   // cp.async issued since the last commit_group
    std::vector<std::vector<AsyncCopy>> copy_groups;      // committed, not yet waited for
};
struct Warp { uint32_t slot[32]; int arrived; unsigned generation; int live; };
struct Barrier { int arrived; unsigned generation; int count; };
struct Cta {                                             // one thread block of the cluster that is running
    BlockState bs{};
    Warp warps[MAX_WARPS]{};
    Barrier barriers[NAMED_BARRIERS]{};
    int live = 0;
};

struct Engine {
    Cta ctas[MAX_CLUSTER];
    int cluster = 1;
    Barrier cluster_barrier{};
    int cluster_live = 0;
    std::vector<Fiber> fibers;
    Fiber* current = nullptr;
    ucontext_t scheduler{};
    const std::function<void()>* body = nullptr;
    unsigned long progress = 0;
    unsigned char* stacks = nullptr;
    unsigned char* smem = nullptr;                       // MAX_CLUSTER windows of SMEM_BYTES, 1024-byte aligned
    std::mutex lock;
};

Engine& engine() {
    static Engine* e = new Engine();
    return *e;
}

void try_release(int live, Barrier& b, Engine& e) {
    const int need = b.count > 0 ? b.count : live;
    if (b.arrived > 0 && b.arrived >= need) { b.arrived = 0; ++b.generation; ++e.progress; }
}

void try_release(Engine& e, Warp& w) {
    if (w.arrived > 0 && w.arrived >= w.live) { w.arrived = 0; ++w.generation; ++e.progress; }
}

void fiber_main() {
    Engine& e = engine();
    (*e.body)();
    Fiber& f = *e.current;
    for (const auto& group : f.copy_groups)               // a thread that exits lets its outstanding copies land
        for (const AsyncCopy& c : group) std::memcpy(c.dst, c.src, 16);
    for (const AsyncCopy& c : f.open_copies) std::memcpy(c.dst, c.src, 16);
    Cta& c = e.ctas[f.cta];
    f.done = true;
    ++e.progress;
    --c.live;
    --e.cluster_live;
    --c.warps[f.ts.warp].live;
    try_release(e, c.warps[f.ts.warp]);                  // an exited lane no longer takes part in __syncwarp / shuffles
    try_release(c.live, c.barriers[0], e);               // ... nor in __syncthreads
    try_release(e.cluster_live, e.cluster_barrier, e);
}

void run_cluster(Engine& e, dim3 block_dim) {
    const int n = static_cast<int>(block_dim.x * block_dim.y * block_dim.z);
    e.fibers.assign(static_cast<size_t>(n) * e.cluster, Fiber{});
    e.cluster_barrier = Barrier{};
    e.cluster_live = n * e.cluster;
    for (int c = 0; c < e.cluster; ++c) {
        Cta& cta = e.ctas[c];
        for (Warp& w : cta.warps) w = Warp{};
        for (Barrier& b : cta.barriers) b = Barrier{};
        cta.live = n;
        for (int i = 0; i < n; ++i) {
            Fiber& f = e.fibers[static_cast<size_t>(c) * n + i];
            f.cta = c;
            f.ts.linear = i; f.ts.lane = i & 31; f.ts.warp = i >> 5;
            f.ts.tid = make_uint3(i % block_dim.x, (i / block_dim.x) % block_dim.y, i / (block_dim.x * block_dim.y));
            f.done = false;
            ++cta.warps[f.ts.warp].live;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = e.stacks + (static_cast<size_t>(c) * MAX_THREADS + i) * STACK_BYTES;
            f.ctx.uc_stack.ss_size = STACK_BYTES;
            f.ctx.uc_link = &e.scheduler;
            makecontext(&f.ctx, fiber_main, 0);
        }
    }
    reset_block_resources(e.cluster);                    // tcgen05_emu.cpp: mbarriers, tensor memory, bulk-copy state
    while (e.cluster_live > 0) {
        const unsigned long before = e.progress;
        for (Fiber& f : e.fibers) {
            if (f.done) continue;
            e.current = &f;
            swapcontext(&e.scheduler, &f.ctx);
        }
        if (e.progress == before) {
            const BlockState& bs = e.ctas[0].bs;
            std::fprintf(stderr, "cuda_emu: deadlock in the cluster of block (%u,%u,%u): %d threads wait on barriers nobody else reaches\n",
                         bs.block_idx.x, bs.block_idx.y, bs.block_idx.z, e.cluster_live);
            std::abort();
        }
    }
}

}  // namespace

void cp_async_enqueue(void* smem_dst, const void* gmem_src) { engine().current->open_copies.push_back({smem_dst, gmem_src}); }
void cp_async_commit_group() {
    Fiber& f = *engine().current;
    f.copy_groups.push_back(std::move(f.open_copies));
    f.open_copies.clear();
}
void cp_async_wait_group(int pending_allowed) {
    Fiber& f = *engine().current;
    while (static_cast<int>(f.copy_groups.size()) > pending_allowed) {
        for (const AsyncCopy& c : f.copy_groups.front()) std::memcpy(c.dst, c.src, 16);
        f.copy_groups.erase(f.copy_groups.begin());
    }
}

void yield() {
    Engine& e = engine();
    swapcontext(&e.current->ctx, &e.scheduler);
}
void made_progress() { ++engine().progress; }

ThreadState& thread() { return engine().current->ts; }
BlockState& block() { return engine().ctas[engine().current->cta].bs; }
int cta_rank() { return engine().current->cta; }
int cluster_size() { return engine().cluster; }
unsigned char* smem_window(int cta) { return engine().smem + static_cast<size_t>(cta) * SMEM_BYTES; }
void* dynamic_smem() { return smem_window(engine().current->cta); }

void named_barrier(int id, int count, bool wait) {
    Engine& e = engine();
    Cta& c = e.ctas[e.current->cta];
    Barrier& b = c.barriers[id];
    const unsigned generation = b.generation;
    ++b.arrived;
    b.count = count;
    try_release(c.live, b, e);
    if (!wait) return;
    while (b.generation == generation) yield();
}

void cluster_barrier() {
    Engine& e = engine();
    Barrier& b = e.cluster_barrier;
    const unsigned generation = b.generation;
    ++b.arrived;
    b.count = 0;
    try_release(e.cluster_live, b, e);
    while (b.generation == generation) yield();
}

void warp_barrier() {
    Engine& e = engine();
    Warp& w = e.ctas[e.current->cta].warps[e.current->ts.warp];
    const unsigned generation = w.generation;
    ++w.arrived;
    try_release(e, w);
    while (w.generation == generation) yield();
}

uint32_t warp_exchange(uint32_t value, int source_lane) {
    Engine& e = engine();
    Warp& w = e.ctas[e.current->cta].warps[e.current->ts.warp];
    w.slot[e.current->ts.lane] = value;
    warp_barrier();
    const uint32_t got = w.slot[source_lane];
    warp_barrier();                                      // nobody overwrites a slot before every lane has read
    return got;
}

void run_grid(const std::function<void()>& body, dim3 grid, dim3 block_dim, size_t smem_bytes, int cluster) {
    Engine& e = engine();
    std::lock_guard<std::mutex> guard(e.lock);
    const size_t threads = static_cast<size_t>(block_dim.x) * block_dim.y * block_dim.z;
    if (threads == 0 || threads > MAX_THREADS || smem_bytes > SMEM_BYTES || cluster < 1 || cluster > MAX_CLUSTER ||
        grid.x % cluster != 0) {
        std::fprintf(stderr, "cuda_emu: launch of %zu threads / %zu bytes of shared memory / cluster %d is outside the model\n",
                     threads, smem_bytes, cluster);
        std::abort();
    }
    if (e.stacks == nullptr) {
        void* p = mmap(nullptr, MAX_CLUSTER * MAX_THREADS * STACK_BYTES, PROT_READ | PROT_WRITE,
                       MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        void* s = mmap(nullptr, MAX_CLUSTER * SMEM_BYTES + 1024, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED || s == MAP_FAILED) { std::perror("cuda_emu: mmap"); std::abort(); }
        e.stacks = static_cast<unsigned char*>(p);
        e.smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(s) + 1023) & ~static_cast<uintptr_t>(1023));
    }
    e.body = &body;
    e.cluster = cluster;
    ::gridDim = grid;
    ::blockDim = block_dim;
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; x += cluster) {
                for (int c = 0; c < cluster; ++c) {
                    e.ctas[c].bs.grid_dim = grid;
                    e.ctas[c].bs.block_dim = block_dim;
                    e.ctas[c].bs.block_idx = make_uint3(x + c, y, z);
                }
                run_cluster(e, block_dim);
            }
    e.body = nullptr;
}

}  // namespace cuda_emu

// ---- the CUDA runtime entry points the host code of the library calls ------------------------------------------
extern "C" {
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorName(cudaError_t) { return "cudaEmu"; }
const char* cudaGetErrorString(cudaError_t) { return "host emulation"; }
cudaError_t cudaDeviceGetAttribute(int* value, enum cudaDeviceAttr, int) { *value = 4; return cudaSuccess; }   // "SM count"
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

// entry points of the translation unit that is not emulated (NVLink peer memory)
size_t aecf_peer_flag_bytes(void) { return 256; }
int aecf_peer_enable_access(int32_t, int32_t) { return AECF_ERR_UNSUPPORTED; }
int aecf_peer_allreduce(const aecf_peer_desc*, void* const*, void* const*, void*) { return AECF_ERR_UNSUPPORTED; }
int aecf_peer_export(int32_t, const void*, void*, int64_t*) { return AECF_ERR_UNSUPPORTED; }
int aecf_peer_import(int32_t, const void*, void**) { return AECF_ERR_UNSUPPORTED; }
}
namespace aecf {
int launch_peer_sum(int, const aecf_dp_desc*, long long, cudaStream_t) { return AECF_ERR_UNSUPPORTED; }   // peer_allreduce.cu
}
