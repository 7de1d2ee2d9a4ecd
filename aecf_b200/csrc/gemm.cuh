// Shared pieces of the GEMM entry point: the epilogue functor and the tcgen05 hooks.
#pragma once

#include "common.cuh"

namespace aecf {

struct GemmEpilogue {
    void* C;
    long long ldc;
    const void* bias;
    int dtype_c, dtype_bias, accumulate;
    float* partial;                    // split-K: raw fp32 partial sums [split][M][N], epilogue applied by the reduce

    __device__ __forceinline__ float bias_at(long long j) const {
        if (bias == nullptr) return 0.f;
        return dtype_bias == AECF_BF16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(bias)[j])
                                       : static_cast<const float*>(bias)[j];
    }
    __device__ __forceinline__ void store(long long i, long long j, float v, int split, long long M, long long N) const {
        if (partial != nullptr) {
            partial[(static_cast<long long>(split) * M + i) * N + j] = v;
            return;
        }
        v += bias_at(j);
        if (dtype_c == AECF_BF16) {
            __nv_bfloat16* c = static_cast<__nv_bfloat16*>(C) + i * ldc + j;
            if (accumulate) v += __bfloat162float(*c);
            *c = __float2bfloat16_rn(v);
        } else {
            float* c = static_cast<float*>(C) + i * ldc + j;
            if (accumulate) v += *c;
            *c = v;
        }
    }
};

inline GemmEpilogue make_epilogue(const aecf_gemm_desc* d, const void* bias, void* C) {
    GemmEpilogue ep;
    ep.C = C; ep.ldc = d->ldc; ep.bias = bias;
    ep.dtype_c = d->dtype_c; ep.dtype_bias = d->dtype_bias; ep.accumulate = d->accumulate;
    ep.partial = nullptr;
    return ep;
}

int launch_splitk_reduce(const float* partial, long long M, long long N, int splits, long long split_stride,
                         const GemmEpilogue& ep, cudaStream_t s);

// A product whose split-K fold is left to the caller: `splits` fp32 [m, n] slabs, `stride` elements apart (splits == 1:
// the product itself).  grad_tail.cu folds the weight-gradient products together with the other batch reductions of
// the backward instead of paying one reduce launch per product.
struct GemmPartials { const float* partial; int splits; long long stride; };
// fp32 product of `d` (dtype_c / ldc of the descriptor are ignored: fp32, ldc = n) left as partials inside `workspace`
int gemm_partials(const aecf_gemm_desc* d, const void* A, const void* B, void* workspace, size_t workspace_bytes,
                  cudaStream_t s, GemmPartials* out);
size_t gemm_partials_workspace_bytes(const aecf_gemm_desc* d);

// gemm_tcgen05.cu: returns AECF_ERR_UNSUPPORTED when the shape/dtype is outside what it covers.
// aux != nullptr: the side output of aecf_gemm_aux (B then has d->n + roundup8(aux_cols) rows).
// defer != nullptr: no split-K reduce launch; *defer says where the partials (or, unsplit, the product in C) are.
int gemm_tcgen05(const aecf_gemm_desc* d, const void* A, const void* B, const void* bias, void* C,
                 void* workspace, size_t workspace_bytes, cudaStream_t s, float* aux = nullptr, int aux_cols = 0,
                 long long aux_ld = 0, GemmPartials* defer = nullptr);
size_t gemm_tcgen05_workspace_bytes(const aecf_gemm_desc* d);
void note_gemm_kernel(const char* fmt, ...);          // gemm.cu: records what aecf_gemm_last_kernel() reports

}  // namespace aecf
