// Fiber scheduler and CUDA-runtime stand-ins of the host emulation -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
#include "cuda_emu.h"

#include <sys/mman.h>
#include <ucontext.h>

#include <mutex>
#include <vector>

#include "../../include/aecf_b200.h"

namespace cuda_emu {
namespace {

constexpr size_t STACK_BYTES = 1 << 20;
constexpr int MAX_THREADS = 1024, MAX_WARPS = 32, NAMED_BARRIERS = 16;
constexpr size_t SMEM_BYTES = 232448;                    // the 227 KB a CTA can opt into on sm_100

struct Fiber { ThreadState ts; ucontext_t ctx; bool done; };
struct Warp { uint32_t slot[32]; int arrived; unsigned generation; int live; };
struct Barrier { int arrived; unsigned generation; int count; };

struct Engine {
    BlockState bs{};
    std::vector<Fiber> fibers;
    Warp warps[MAX_WARPS]{};
    Barrier barriers[NAMED_BARRIERS]{};
    int live = 0;
    Fiber* current = nullptr;
    ucontext_t scheduler{};
    const std::function<void()>* body = nullptr;
    unsigned long progress = 0;
    unsigned char* stacks = nullptr;
    alignas(1024) unsigned char smem[SMEM_BYTES];
    std::mutex lock;
};

Engine& engine() {
    static Engine* e = new Engine();
    return *e;
}

void yield() {
    Engine& e = engine();
    swapcontext(&e.current->ctx, &e.scheduler);
}

void try_release(Engine& e, Barrier& b) {
    const int need = b.count > 0 ? b.count : e.live;
    if (b.arrived > 0 && b.arrived >= need) { b.arrived = 0; ++b.generation; ++e.progress; }
}

void try_release(Engine& e, Warp& w) {
    if (w.arrived > 0 && w.arrived >= w.live) { w.arrived = 0; ++w.generation; ++e.progress; }
}

void fiber_main() {
    Engine& e = engine();
    (*e.body)();
    Fiber& f = *e.current;
    f.done = true;
    ++e.progress;
    --e.live;
    --e.warps[f.ts.warp].live;
    try_release(e, e.warps[f.ts.warp]);                  // an exited lane no longer takes part in __syncwarp / shuffles
    try_release(e, e.barriers[0]);                       // ... nor in __syncthreads
}

void run_block(Engine& e, dim3 block_dim) {
    const int n = static_cast<int>(block_dim.x * block_dim.y * block_dim.z);
    e.fibers.assign(n, Fiber{});
    for (Warp& w : e.warps) w = Warp{};
    for (Barrier& b : e.barriers) b = Barrier{};
    e.live = n;
    for (int i = 0; i < n; ++i) {
        Fiber& f = e.fibers[i];
        f.ts.linear = i; f.ts.lane = i & 31; f.ts.warp = i >> 5;
        f.ts.tid = make_uint3(i % block_dim.x, (i / block_dim.x) % block_dim.y, i / (block_dim.x * block_dim.y));
        f.done = false;
        ++e.warps[f.ts.warp].live;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = e.stacks + static_cast<size_t>(i) * STACK_BYTES;
        f.ctx.uc_stack.ss_size = STACK_BYTES;
        f.ctx.uc_link = &e.scheduler;
        makecontext(&f.ctx, fiber_main, 0);
    }
    while (e.live > 0) {
        const unsigned long before = e.progress;
        for (Fiber& f : e.fibers) {
            if (f.done) continue;
            e.current = &f;
            swapcontext(&e.scheduler, &f.ctx);
        }
        if (e.progress == before) {
            std::fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %d threads wait on barriers nobody else reaches\n",
                         e.bs.block_idx.x, e.bs.block_idx.y, e.bs.block_idx.z, e.live);
            std::abort();
        }
    }
}

}  // namespace

ThreadState& thread() { return engine().current->ts; }
BlockState& block() { return engine().bs; }
void* dynamic_smem() { return engine().smem; }

void named_barrier(int id, int count, bool wait) {
    Engine& e = engine();
    Barrier& b = e.barriers[id];
    const unsigned generation = b.generation;
    ++b.arrived;
    b.count = count;
    try_release(e, b);
    if (!wait) return;
    while (b.generation == generation) yield();
}

void warp_barrier() {
    Engine& e = engine();
    Warp& w = e.warps[e.current->ts.warp];
    const unsigned generation = w.generation;
    ++w.arrived;
    try_release(e, w);
    while (w.generation == generation) yield();
}

uint32_t warp_exchange(uint32_t value, int source_lane) {
    Engine& e = engine();
    Warp& w = e.warps[e.current->ts.warp];
    w.slot[e.current->ts.lane] = value;
    warp_barrier();
    const uint32_t got = w.slot[source_lane];
    warp_barrier();                                      // nobody overwrites a slot before every lane has read
    return got;
}

void run_grid(const std::function<void()>& body, dim3 grid, dim3 block_dim, size_t smem_bytes) {
    Engine& e = engine();
    std::lock_guard<std::mutex> guard(e.lock);
    const size_t threads = static_cast<size_t>(block_dim.x) * block_dim.y * block_dim.z;
    if (threads == 0 || threads > MAX_THREADS || smem_bytes > SMEM_BYTES) {
        std::fprintf(stderr, "cuda_emu: launch of %zu threads / %zu bytes of shared memory is outside the model\n", threads, smem_bytes);
        std::abort();
    }
    if (e.stacks == nullptr) {
        void* p = mmap(nullptr, MAX_THREADS * STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { std::perror("cuda_emu: mmap"); std::abort(); }
        e.stacks = static_cast<unsigned char*>(p);
    }
    e.body = &body;
    e.bs.grid_dim = grid;
    e.bs.block_dim = block_dim;
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                e.bs.block_idx = make_uint3(x, y, z);
                run_block(e, block_dim);
            }
    e.body = nullptr;
}

}  // namespace cuda_emu

// ---- the CUDA runtime entry points the host code of the library calls ------------------------------------------
extern "C" {
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorName(cudaError_t) { return "cudaEmu"; }
const char* cudaGetErrorString(cudaError_t) { return "host emulation"; }
cudaError_t cudaDeviceGetAttribute(int* value, enum cudaDeviceAttr, int) { *value = 4; return cudaSuccess; }   // "SM count"
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

// entry points of the translation units that are not emulated (NVLink peer memory)
size_t aecf_peer_flag_bytes(void) { return 256; }
int aecf_peer_enable_access(int32_t, int32_t) { return AECF_ERR_UNSUPPORTED; }
int aecf_peer_allreduce(const aecf_peer_desc*, void* const*, void* const*, void*) { return AECF_ERR_UNSUPPORTED; }
}

// gemm_tcgen05.cu (TMA / TMEM / tcgen05) is not emulated: the dispatcher falls through to the SIMT kernel
namespace aecf {
int gemm_tcgen05(const aecf_gemm_desc*, const void*, const void*, const void*, void*, void*, size_t, cudaStream_t, float*, int,
                 long long) {
    return AECF_ERR_UNSUPPORTED;
}
size_t gemm_tcgen05_workspace_bytes(const aecf_gemm_desc*) { return 0; }
}  // namespace aecf
