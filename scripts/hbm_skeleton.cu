// What HBM rate can a kernel with the pool kernels' traffic reach on this box?  
//
// Stand-alone measurement, not part of the library: the memory skeleton of pool_bwd_kernel / pool_fwd_stream_kernel at
// config 2 (per sample: 3 value rows of 1 KB + 1 KB of d_ctx + 96 B of scores in; 3 rows of 1040 B out) with the arithmetic
// replaced by a programmable number of dependent FFMAs, next to a plain copy and a plain read on the same box.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_skeleton hbm_skeleton.cu && ./hbm_skeleton
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_plain(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ void stg_cs(void* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ int g_store_hint;   // 0: plain stores, 1: st.global.cs
__device__ __forceinline__ void stg(void* p, uint4 v) { if (g_store_hint) stg_cs(p, v); else stg_plain(p, v); }

struct Args {
    const char* v; const char* dctx; const float* scores; char* dvs; char* ctx;
    long long B; int out_row_bytes; int work;
    int chunk;     // > 0: CTA b handles samples [b * chunk, (b + 1) * chunk), 8 at a time; 0: grid-stride
};

__device__ __forceinline__ float spin(float x, int n) {            // n dependent FFMAs
    for (int i = 0; i < n; ++i) x = fmaf(x, 1.0000001f, 1e-9f);
    return x;
}

// backward skeleton: every load of a sample first, "compute", 6 x 512 B + 3 score-gradient stores.  DEPTH 2: the next
// sample's loads are issued before this sample's compute (register double buffering).
template <int DEPTH, int MINB>
__global__ void __launch_bounds__(256, MINB) bwd_skel(const Args a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long stride = a.chunk ? 8 : static_cast<long long>(gridDim.x) * 8;
    long long s = a.chunk ? static_cast<long long>(blockIdx.x) * a.chunk + warp : static_cast<long long>(blockIdx.x) * 8 + warp;
    const long long end = a.chunk ? min(a.B, static_cast<long long>(blockIdx.x + 1) * a.chunk) : a.B;
    uint4 d[2], v[3][2]; float sc[3];
    auto load = [&](long long smp, uint4 (&dd)[2], uint4 (&vv)[3][2], float (&ss)[3]) {
        const char* dc = a.dctx + smp * 1024 + lane * 16;
        const char* vr = a.v + smp * 3072 + lane * 16;
#pragma unroll
        for (int j = 0; j < 2; ++j) dd[j] = ldg_stream(dc + j * 512);
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 2; ++j) vv[m][j] = ldg_stream(vr + m * 1024 + j * 512);
#pragma unroll
        for (int m = 0; m < 3; ++m) ss[m] = __ldg(a.scores + (smp * 3 + m) * 8 + (lane >> 2));
    };
    if (s < end) load(s, d, v, sc);
    for (; s < end; s += stride) {
        uint4 d2[2], v2[3][2]; float sc2[3];
        const long long nx = s + stride;
        if (DEPTH == 2 && nx < end) load(nx, d2, v2, sc2);
        float f = spin(sc[0] + sc[1] + sc[2] + __uint_as_float(d[0].x ^ d[1].y), a.work);
        const unsigned k = __float_as_uint(f);
        char* o = a.dvs + s * 3 * a.out_row_bytes + lane * 16;
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint4 r = v[m][j];
                r.x ^= d[j].x ^ k; r.y ^= d[j].y; r.z ^= d[j].z; r.w ^= d[j].w;
                stg(o + m * a.out_row_bytes + j * 512, r);
            }
        if (a.out_row_bytes > 1024) {
            if ((lane & 3) == 0)
#pragma unroll
                for (int m = 0; m < 3; ++m)
                    *reinterpret_cast<unsigned short*>(a.dvs + (s * 3 + m) * a.out_row_bytes + 1024 + (lane >> 2) * 2) = static_cast<unsigned short>(k + m);
        }
        if (DEPTH == 2) {
#pragma unroll
            for (int j = 0; j < 2; ++j) d[j] = d2[j];
#pragma unroll
            for (int m = 0; m < 3; ++m) { sc[m] = sc2[m];
#pragma unroll
                for (int j = 0; j < 2; ++j) v[m][j] = v2[m][j]; }
        } else if (nx < end) load(nx, d, v, sc);
    }
}


// backward skeleton, persistent, samples handed out one at a time by an atomic counter (per warp), the next index fetched
// before the current sample's stores
template <int MINB>
__global__ void __launch_bounds__(256, MINB) bwd_skel_dyn(const Args a, unsigned* counter) {
    const int lane = threadIdx.x & 31;
    uint4 d[2], v[3][2]; float sc[3];
    auto load = [&](long long smp, uint4 (&dd)[2], uint4 (&vv)[3][2], float (&ss)[3]) {
        const char* dc = a.dctx + smp * 1024 + lane * 16;
        const char* vr = a.v + smp * 3072 + lane * 16;
#pragma unroll
        for (int j = 0; j < 2; ++j) dd[j] = ldg_stream(dc + j * 512);
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 2; ++j) vv[m][j] = ldg_stream(vr + m * 1024 + j * 512);
#pragma unroll
        for (int m = 0; m < 3; ++m) ss[m] = __ldg(a.scores + (smp * 3 + m) * 8 + (lane >> 2));
    };
    auto fetch = [&]() -> long long {
        unsigned i = 0;
        if (lane == 0) i = atomicAdd(counter, static_cast<unsigned>(a.chunk));
        return __shfl_sync(0xffffffffu, i, 0);
    };
    // a.chunk consecutive samples per fetch
    long long s0 = fetch();
    while (s0 < a.B) {
        const long long nxt = fetch();
        for (long long s = s0; s < min(a.B, s0 + a.chunk); ++s) {
            load(s, d, v, sc);
            float f = spin(sc[0] + sc[1] + sc[2] + __uint_as_float(d[0].x ^ d[1].y), a.work);
            const unsigned k = __float_as_uint(f);
            char* o = a.dvs + s * 3 * a.out_row_bytes + lane * 16;
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint4 r = v[m][j];
                    r.x ^= d[j].x ^ k; r.y ^= d[j].y; r.z ^= d[j].z; r.w ^= d[j].w;
                    stg(o + m * a.out_row_bytes + j * 512, r);
                }
            if ((lane & 3) == 0)
#pragma unroll
                for (int m = 0; m < 3; ++m)
                    *reinterpret_cast<unsigned short*>(a.dvs + (s * 3 + m) * a.out_row_bytes + 1024 + (lane >> 2) * 2) = static_cast<unsigned short>(k + m);
        }
        s0 = nxt;
    }
}

// forward skeleton: 3 KB of values + scores in, 1 KB of context out
template <int DEPTH, int MINB>
__global__ void __launch_bounds__(256, MINB) fwd_skel(const Args a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long stride = a.chunk ? 8 : static_cast<long long>(gridDim.x) * 8;
    long long s = a.chunk ? static_cast<long long>(blockIdx.x) * a.chunk + warp : static_cast<long long>(blockIdx.x) * 8 + warp;
    const long long end = a.chunk ? min(a.B, static_cast<long long>(blockIdx.x + 1) * a.chunk) : a.B;
    uint4 v[3][2]; float sc[3];
    auto load = [&](long long smp, uint4 (&vv)[3][2], float (&ss)[3]) {
        const char* vr = a.v + smp * 3072 + lane * 16;
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int j = 0; j < 2; ++j) vv[m][j] = ldg_stream(vr + m * 1024 + j * 512);
#pragma unroll
        for (int m = 0; m < 3; ++m) ss[m] = __ldg(a.scores + (smp * 3 + m) * 8 + (lane >> 2));
    };
    if (s < end) load(s, v, sc);
    for (; s < end; s += stride) {
        uint4 v2[3][2]; float sc2[3];
        const long long nx = s + stride;
        if (DEPTH == 2 && nx < end) load(nx, v2, sc2);
        float f = spin(sc[0] + sc[1] + sc[2], a.work);
        const unsigned k = __float_as_uint(f);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            uint4 r = v[0][j];
            r.x ^= v[1][j].x ^ v[2][j].x ^ k; r.y ^= v[1][j].y ^ v[2][j].y; r.z ^= v[1][j].z ^ v[2][j].z; r.w ^= v[1][j].w ^ v[2][j].w;
            stg(a.ctx + s * 1024 + j * 512 + lane * 16, r);
        }
        if (DEPTH == 2) {
#pragma unroll
            for (int m = 0; m < 3; ++m) { sc[m] = sc2[m];
#pragma unroll
                for (int j = 0; j < 2; ++j) v[m][j] = v2[m][j]; }
        } else if (nx < end) load(nx, v, sc);
    }
}

__global__ void copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = in[i];
}
__global__ void read_kernel(const uint4* __restrict__ in, unsigned* out, long long n) {
    unsigned acc = 0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        uint4 r = ldg_stream(in + i); acc ^= r.x ^ r.y ^ r.z ^ r.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <typename F> float time_us(F launch, int reps = 20) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    return ms * 1000.f / reps;
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const long long B = 65536;
    // two buffer sets, alternated, so that nothing a launch reads is left in the 126 MB L2 by the launch before it
    Args a[2];
    for (int i = 0; i < 2; ++i) {
        char *v, *dc, *dvs, *ctx; float* sc;
        CK(cudaMalloc(&v, B * 3072)); CK(cudaMalloc(&dc, B * 1024)); CK(cudaMalloc(&sc, B * 96));
        CK(cudaMalloc(&dvs, B * 3 * 1040)); CK(cudaMalloc(&ctx, B * 1024));
        CK(cudaMemset(v, 1, B * 3072)); CK(cudaMemset(dc, 2, B * 1024)); CK(cudaMemset(sc, 0, B * 96));
        a[i] = Args{v, dc, sc, dvs, ctx, B, 1040, 0, 0};
    }
    printf("{\"sms\": %d, \"rows\": [\n", sms);
    const long long n16 = (1ll << 30) / 16;                                   // 1 GiB each way
    uint4 *cin, *cout; CK(cudaMalloc(&cin, n16 * 16)); CK(cudaMalloc(&cout, n16 * 16)); CK(cudaMemset(cin, 3, n16 * 16));
    unsigned* sink; CK(cudaMalloc(&sink, 4));
    for (int blocks : {sms * 8, sms * 32, 1 << 16}) {
        float t = time_us([&] { copy_kernel<<<blocks, 256>>>(cin, cout, n16); });
        printf(" {\"kernel\": \"copy 1 GiB\", \"blocks\": %d, \"us\": %.1f, \"gbs\": %.0f},\n", blocks, t, 2.0 * n16 * 16 / t * 1e-3);
        t = time_us([&] { read_kernel<<<blocks, 256>>>(cin, sink, n16); });
        printf(" {\"kernel\": \"read 1 GiB\", \"blocks\": %d, \"us\": %.1f, \"gbs\": %.0f},\n", blocks, t, 1.0 * n16 * 16 / t * 1e-3);
    }
    {
        float t = time_us([&] { CK(cudaMemcpyAsync(cout, cin, n16 * 16, cudaMemcpyDeviceToDevice)); });
        printf(" {\"kernel\": \"cudaMemcpy D2D 1 GiB\", \"us\": %.1f, \"gbs\": %.0f},\n", t, 2.0 * n16 * 16 / t * 1e-3);
    }
    const double fwd_bytes = B * (3072.0 + 96 + 1024);
    int flip = 0;
    auto run = [&](const char* name, auto kern, bool bwd, int per_sm, int chunk, int work) {
        Args x0 = a[0], x1 = a[1]; x0.work = x1.work = work; x0.chunk = x1.chunk = chunk;
        const int grid = chunk ? static_cast<int>((B + chunk - 1) / chunk) : sms * per_sm;
        float t = time_us([&] { kern<<<grid, 256>>>((flip ^= 1) ? x0 : x1); });
        const double bytes = bwd ? B * (3072.0 + 1024 + 96 + 3 * 1040) : fwd_bytes;
        printf(" {\"kernel\": \"%s\", \"ctas_per_sm\": %d, \"grid\": %d, \"samples_per_cta\": %d, \"ffma_per_sample\": %d, \"us\": %.1f, \"gbs\": %.0f},\n",
               name, per_sm, grid, chunk, work, t, bytes / t * 1e-3);
    };
    unsigned* counter; CK(cudaMalloc(&counter, 4));
    for (int rep = 0; rep < 2; ++rep)
    for (int work : {0, 400}) {
        for (int chunk : {0, 8}) {
            run("bwd skeleton, loads then stores", bwd_skel<1, 3>, true, 3, chunk, work);
        }
        {   // the same static persistent kernel with a memset in front of every launch, as the dynamic one needs
            Args x0 = a[0], x1 = a[1]; x0.work = x1.work = work;
            float t = time_us([&] { CK(cudaMemsetAsync(counter, 0, 4)); bwd_skel<1, 3><<<sms * 3, 256>>>((flip ^= 1) ? x0 : x1); });
            printf(" {\"kernel\": \"bwd skeleton, loads then stores, memset in front\", \"ctas_per_sm\": 3, \"grid\": %d, \"samples_per_cta\": 0, \"ffma_per_sample\": %d, \"us\": %.1f, \"gbs\": %.0f},\n",
                   sms * 3, work, t, B * (3072.0 + 1024 + 96 + 3 * 1040) / t * 1e-3);
        }
        for (int per_sm : {3, 4})
        for (int chunk : {1, 2, 4}) {
            Args x0 = a[0], x1 = a[1]; x0.work = x1.work = work; x0.chunk = x1.chunk = chunk;
            float t = per_sm == 3 ? time_us([&] { CK(cudaMemsetAsync(counter, 0, 4)); bwd_skel_dyn<3><<<sms * 3, 256>>>((flip ^= 1) ? x0 : x1, counter); })
                                  : time_us([&] { CK(cudaMemsetAsync(counter, 0, 4)); bwd_skel_dyn<4><<<sms * 4, 256>>>((flip ^= 1) ? x0 : x1, counter); });
            printf(" {\"kernel\": \"bwd skeleton, persistent, atomic work counter per warp\", \"ctas_per_sm\": %d, \"grid\": %d, \"samples_per_cta\": %d, \"ffma_per_sample\": %d, \"us\": %.1f, \"gbs\": %.0f},\n",
                   per_sm, sms * per_sm, chunk, work, t, B * (3072.0 + 1024 + 96 + 3 * 1040) / t * 1e-3);
        }
    }
    printf(" {}]}\n");
    return 0;
}
