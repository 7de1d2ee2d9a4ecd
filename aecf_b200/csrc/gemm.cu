// aecf_gemm: C[m,n] = sum_k A[m,k] B[n,k] (+ bias[n]) (+ C).  Dispatches to the tcgen05/TMEM/TMA
// kernel (gemm_tcgen05.cu) for bf16 operands and to the SIMT kernel below otherwise (fp32 parity
// path, tiny GEMV-shaped products such as the shared-query projection, ragged shapes).
#include <cstdarg>
#include <cstdio>

#include "gemm.cuh"

namespace aecf {

static thread_local char g_last_gemm_kernel[96] = "";
void note_gemm_kernel(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(g_last_gemm_kernel, sizeof(g_last_gemm_kernel), fmt, ap);
    va_end(ap);
    note_site_gemm_kernel(g_last_gemm_kernel);
}

// ---- SIMT tile kernel: 64x64x16, 256 threads, 4x4 per thread, fp32 accumulate ---------------
constexpr int SBM = 64, SBN = 64, SBK = 16;

template <typename T> __device__ __forceinline__ float load_elem(const T* p) { return to_float<T>(__ldg(p)); }

template <typename TA, typename TB, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, const TB* __restrict__ B, long long M, long long N, long long K,
                 long long lda, long long ldb, long long k_per_split, GemmEpilogue ep) {
    __shared__ __align__(16) float As[SBK][SBM + 4];
    __shared__ __align__(16) float Bs[SBK][SBN + 4];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const long long m0 = static_cast<long long>(blockIdx.y) * SBM, n0 = static_cast<long long>(blockIdx.x) * SBN;
    const long long k_begin = static_cast<long long>(blockIdx.z) * k_per_split;
    const long long k_end = min(K, k_begin + k_per_split);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    pdl_wait();

    for (long long k0 = k_begin; k0 < k_end; k0 += SBK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = t + 256 * i;
            {
                const int mm = A_KMAJOR ? idx / SBK : idx % SBM, kk = A_KMAJOR ? idx % SBK : idx / SBM;
                const long long gm = m0 + mm, gk = k0 + kk;
                float v = 0.f;
                if (gm < M && gk < k_end) v = load_elem(A + (A_KMAJOR ? gm * lda + gk : gk * lda + gm));
                As[kk][mm] = v;
            }
            {
                const int nn = B_KMAJOR ? idx / SBK : idx % SBN, kk = B_KMAJOR ? idx % SBK : idx / SBN;
                const long long gn = n0 + nn, gk = k0 + kk;
                float v = 0.f;
                if (gn < N && gk < k_end) v = load_elem(B + (B_KMAJOR ? gn * ldb + gk : gk * ldb + gn));
                Bs[kk][nn] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SBK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long gn = n0 + tx * 4 + j;
            if (gn < N) ep.store(gm, gn, acc[i][j], blockIdx.z, M, N);
        }
    }
}

// Fold split-K partials [splits][M][N] (fp32) in order, then apply the epilogue.
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, long long M, long long N, int splits,
                                     long long split_stride, GemmEpilogue ep) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    pdl_wait();
    if (i >= M * N) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[static_cast<long long>(z) * split_stride + i];
    GemmEpilogue direct = ep;
    direct.partial = nullptr;
    direct.store(i / N, i % N, s, 0, M, N);
}

int launch_splitk_reduce(const float* partial, long long M, long long N, int splits, long long split_stride,
                         const GemmEpilogue& ep, cudaStream_t s) {
    const long long total = M * N;
    AECF_CUDA_OK(launch_pdl(splitk_reduce_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, partial,
                            M, N, splits, split_stride, ep));
    count_launch();
    return AECF_OK;
}

static int simt_splits(const aecf_gemm_desc* d) {
    const long long tiles = ((d->m + SBM - 1) / SBM) * ((d->n + SBN - 1) / SBN);
    if (d->m == 1 || tiles >= 148 || d->k < 2048) return 1;
    long long s = (2 * 148 + tiles - 1) / tiles;
    const long long by_k = d->k / 512;
    if (s > by_k) s = by_k;
    if (s > 64) s = 64;
    return s < 1 ? 1 : static_cast<int>(s);
}

template <typename TA, typename TB>
static int launch_simt(const aecf_gemm_desc* d, const void* A, const void* B, GemmEpilogue ep, int splits,
                       cudaStream_t s) {
    const dim3 grid(static_cast<unsigned>((d->n + SBN - 1) / SBN), static_cast<unsigned>((d->m + SBM - 1) / SBM), splits);
    long long kps = (d->k + splits - 1) / splits;
    kps = (kps + SBK - 1) / SBK * SBK;
    const TA* a = static_cast<const TA*>(A);
    const TB* b = static_cast<const TB*>(B);
    const bool ak = d->a_layout == AECF_K_MAJOR, bk = d->b_layout == AECF_K_MAJOR;
#define AECF_SIMT(AK, BK) \
    AECF_CUDA_OK(launch_pdl(gemm_simt_kernel<TA, TB, AK, BK>, grid, dim3(256), 0, s, a, b, d->m, d->n, d->k, d->lda, d->ldb, kps, ep))
    if (ak && bk) AECF_SIMT(true, true);
    else if (ak && !bk) AECF_SIMT(true, false);
    else if (!ak && bk) AECF_SIMT(false, true);
    else AECF_SIMT(false, false);
#undef AECF_SIMT
    count_launch();
    return AECF_OK;
}

// ---- single-row products (the shared fusion query and its gradient): out[n] = sum_k a[k] B(n, k) -----
// B K-major: one warp per output, lanes stride k.  B MN-major: 32 outputs x 8 k-lanes per block.
template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
gemv_kmajor_kernel(const TA* __restrict__ a, long long a_stride, const TB* __restrict__ B, long long N, long long K,
                   long long ldb, GemmEpilogue ep) {
    const long long n = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    pdl_wait();
    if (n >= N) return;
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    for (long long k = lane; k < K; k += 32) acc = fmaf(load_elem(a + k * a_stride), load_elem(B + n * ldb + k), acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, off);
    if (lane == 0) ep.store(0, n, acc, 0, 1, N);
}

template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
gemv_mnmajor_kernel(const TA* __restrict__ a, long long a_stride, const TB* __restrict__ B, long long N, long long K,
                    long long ldb, GemmEpilogue ep) {
    __shared__ float red[8][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const long long n = static_cast<long long>(blockIdx.x) * 32 + x;
    float acc = 0.f;
    pdl_wait();
    if (n < N) {
#pragma unroll 4
        for (long long k = y; k < K; k += 8) acc = fmaf(load_elem(a + k * a_stride), load_elem(B + k * ldb + n), acc);
    }
    red[y][x] = acc;
    __syncthreads();
    if (y == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][x];
        ep.store(0, n, s, 0, 1, N);
    }
}

template <typename TA, typename TB>
static int launch_gemv(const aecf_gemm_desc* d, const void* A, const void* B, const GemmEpilogue& ep, cudaStream_t s) {
    const TA* a = static_cast<const TA*>(A);
    const TB* b = static_cast<const TB*>(B);
    const long long a_stride = d->a_layout == AECF_K_MAJOR ? 1 : d->lda;
    if (d->b_layout == AECF_K_MAJOR)
        AECF_CUDA_OK(launch_pdl(gemv_kmajor_kernel<TA, TB>, dim3(static_cast<unsigned>((d->n + 7) / 8)), dim3(256), 0, s, a,
                                a_stride, b, d->n, d->k, d->ldb, ep));
    else
        AECF_CUDA_OK(launch_pdl(gemv_mnmajor_kernel<TA, TB>, dim3(static_cast<unsigned>((d->n + 31) / 32)), dim3(256), 0, s, a,
                                a_stride, b, d->n, d->k, d->ldb, ep));
    count_launch();
    return AECF_OK;
}

static int gemm_simt(const aecf_gemm_desc* d, const void* A, const void* B, const void* bias, void* C,
                     void* workspace, size_t workspace_bytes, cudaStream_t s, bool allow_split = true,
                     GemmPartials* defer = nullptr) {
    GemmEpilogue ep = make_epilogue(d, bias, C);
    const bool a16 = d->dtype_a == AECF_BF16, b16 = d->dtype_b == AECF_BF16;
    note_gemm_kernel(d->m == 1 ? "gemv" : "simt");
    if (d->m == 1) {
        if (a16 && b16) return launch_gemv<__nv_bfloat16, __nv_bfloat16>(d, A, B, ep, s);
        if (a16) return launch_gemv<__nv_bfloat16, float>(d, A, B, ep, s);
        if (b16) return launch_gemv<float, __nv_bfloat16>(d, A, B, ep, s);
        return launch_gemv<float, float>(d, A, B, ep, s);
    }
    const int splits = allow_split ? simt_splits(d) : 1;
    if (splits > 1) {
        if (workspace == nullptr || workspace_bytes < static_cast<size_t>(splits) * d->m * d->n * sizeof(float))
            return AECF_ERR_WORKSPACE;
        ep.partial = static_cast<float*>(workspace);
    }
    int rc;
    if (a16 && b16) rc = launch_simt<__nv_bfloat16, __nv_bfloat16>(d, A, B, ep, splits, s);
    else if (a16) rc = launch_simt<__nv_bfloat16, float>(d, A, B, ep, splits, s);
    else if (b16) rc = launch_simt<float, __nv_bfloat16>(d, A, B, ep, splits, s);
    else rc = launch_simt<float, float>(d, A, B, ep, splits, s);
    if (rc == AECF_OK && defer != nullptr) {
        *defer = splits > 1 ? GemmPartials{ep.partial, splits, d->m * d->n} : GemmPartials{static_cast<const float*>(C), 1, 0};
        return rc;
    }
    if (rc != AECF_OK || splits == 1) return rc;
    return launch_splitk_reduce(ep.partial, d->m, d->n, splits, d->m * d->n, ep, s);
}

static bool valid_dtype(int t) { return t == AECF_F32 || t == AECF_BF16; }

static int check_desc(const aecf_gemm_desc* d) {
    if (d == nullptr) return AECF_ERR_INVALID;
    if (!valid_dtype(d->dtype_a) || !valid_dtype(d->dtype_b) || !valid_dtype(d->dtype_c) || !valid_dtype(d->dtype_bias))
        return AECF_ERR_INVALID;
    if (d->m < 0 || d->n < 0 || d->k < 0) return AECF_ERR_INVALID;
    if (d->a_layout != AECF_K_MAJOR && d->a_layout != AECF_MN_MAJOR) return AECF_ERR_INVALID;
    if (d->b_layout != AECF_K_MAJOR && d->b_layout != AECF_MN_MAJOR) return AECF_ERR_INVALID;
    if (d->lda < (d->a_layout == AECF_K_MAJOR ? d->k : d->m)) return AECF_ERR_INVALID;
    if (d->ldb < (d->b_layout == AECF_K_MAJOR ? d->k : d->n)) return AECF_ERR_INVALID;
    if (d->ldc < d->n) return AECF_ERR_INVALID;
    return AECF_OK;
}

// workspace layout: [ product m x n fp32 (used when the kernel does not split) | split-K partials ]
static size_t partials_head_bytes(const aecf_gemm_desc* d) {
    return (static_cast<size_t>(d->m) * d->n * sizeof(float) + 255) & ~static_cast<size_t>(255);
}

size_t gemm_partials_workspace_bytes(const aecf_gemm_desc* d) {
    aecf_gemm_desc f = *d;
    f.dtype_c = AECF_F32; f.ldc = d->n; f.accumulate = 0;
    return partials_head_bytes(d) + aecf_gemm_workspace_bytes(&f);
}

int gemm_partials(const aecf_gemm_desc* d, const void* A, const void* B, void* workspace, size_t workspace_bytes,
                  cudaStream_t s, GemmPartials* out) {
    aecf_gemm_desc f = *d;
    f.dtype_c = AECF_F32; f.ldc = d->n; f.accumulate = 0;
    int rc = check_desc(&f);
    if (rc != AECF_OK) return rc;
    if (!A || !B || !workspace || !out) return AECF_ERR_INVALID;
    if (workspace_bytes < gemm_partials_workspace_bytes(d)) return AECF_ERR_WORKSPACE;
    DeviceScope device_scope__(d->device);
    if ((rc = device_scope__.rc) != AECF_OK) return rc;
    char* head = static_cast<char*>(workspace);
    char* tail = head + partials_head_bytes(d);
    const size_t tail_bytes = workspace_bytes - partials_head_bytes(d);
    TimedLaunch timed(s);
    if (d->impl != AECF_GEMM_SIMT) {
        rc = gemm_tcgen05(&f, A, B, nullptr, head, tail, tail_bytes, s, nullptr, 0, 0, out);
        if (rc != AECF_ERR_UNSUPPORTED || d->impl == AECF_GEMM_TCGEN05) return rc;
    }
    return gemm_simt(&f, A, B, nullptr, head, tail, tail_bytes, s, true, out);
}

}  // namespace aecf

using namespace aecf;

extern "C" {

const char* aecf_gemm_last_kernel(void) { return g_last_gemm_kernel; }

size_t aecf_gemm_workspace_bytes(const aecf_gemm_desc* d) {
    if (check_desc(d) != AECF_OK) return 0;
    size_t simt = 0;
    const int s = simt_splits(d);
    if (s > 1) simt = static_cast<size_t>(s) * d->m * d->n * sizeof(float);
    const size_t tc = gemm_tcgen05_workspace_bytes(d);
    const size_t need = simt > tc ? simt : tc;
    return need < 16 ? 16 : need;
}

int aecf_gemm(const aecf_gemm_desc* d, const void* A, const void* B, const void* bias, void* C, void* workspace,
              size_t workspace_bytes, void* stream) {
    int rc = check_desc(d);
    if (rc != AECF_OK) return rc;
    if (d->m == 0 || d->n == 0) return AECF_OK;
    if (!A || !B || !C) return AECF_ERR_INVALID;
    DeviceScope device_scope__(d->device);
    if ((rc = device_scope__.rc) != AECF_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TimedLaunch timed(s);
    if (d->impl != AECF_GEMM_SIMT) {
        rc = gemm_tcgen05(d, A, B, bias, C, workspace, workspace_bytes, s);
        if (rc != AECF_ERR_UNSUPPORTED || d->impl == AECF_GEMM_TCGEN05) return rc;
    }
    return gemm_simt(d, A, B, bias, C, workspace, workspace_bytes, s);
}

int aecf_gemm_aux(const aecf_gemm_desc* d, const void* A, const void* B, const void* bias, void* C, float* aux,
                  int32_t aux_cols, int64_t aux_ld, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_desc(d);
    if (rc != AECF_OK) return rc;
    if (aux_cols <= 0 || aux_cols > 32 || aux_ld < ((aux_cols + 3) & ~3) || d->accumulate) return AECF_ERR_INVALID;
    if (d->m == 0 || d->n == 0) return AECF_OK;
    if (!A || !B || !C || !aux) return AECF_ERR_INVALID;
    DeviceScope device_scope__(d->device);
    if ((rc = device_scope__.rc) != AECF_OK) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TimedLaunch timed(s);
    if (d->impl != AECF_GEMM_SIMT) {
        rc = gemm_tcgen05(d, A, B, bias, C, workspace, workspace_bytes, s, aux, aux_cols, aux_ld);
        if (rc != AECF_ERR_UNSUPPORTED || d->impl == AECF_GEMM_TCGEN05) return rc;
    }
    // fp32 / small shapes: the same two products as two SIMT launches
    if ((rc = gemm_simt(d, A, B, bias, C, workspace, workspace_bytes, s)) != AECF_OK) return rc;
    aecf_gemm_desc side = *d;
    side.n = (aux_cols + 3) & ~3;                       // the padding columns come out too (zero rows of B)
    side.dtype_c = AECF_F32; side.ldc = aux_ld;
    const int es_b = d->dtype_b == AECF_BF16 ? 2 : 4;
    const int aux_rows = (aux_cols + 16 / es_b - 1) / (16 / es_b) * (16 / es_b);
    if (side.n > aux_rows) side.n = aux_rows;
    const long long skip = d->b_layout == AECF_K_MAJOR ? d->n * d->ldb : d->n;
    return gemm_simt(&side, A, static_cast<const char*>(B) + skip * es_b, nullptr, aux, workspace, workspace_bytes, s,
                     /*allow_split=*/false);
}

}  // extern "C"
