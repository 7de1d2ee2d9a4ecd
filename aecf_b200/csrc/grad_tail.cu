// The tail of the folded backward: everything between the tensor-core products and the parameter gradients, in two
// kernels (three with data parallelism) that run on a side stream NEXT TO the dX product instead of eight small launches
// in front of and behind it (r1: colsum x2, split-K reduce x2, pool_bwd_finalize, fold_finish, query_tail = 86 us of a
// 604 us step, all on the critical path).
//
//   grad_gather   raw fp32 sums S = [ [dWv ; R] | dWo | colsum(d_out) | pool bias sums ]: folds the split-K partials of
//                 the two weight-gradient products, the per-CTA partials of the pool backward and the column sums of
//                 d_out (the gradient of out_proj.bias, torch/nn/functional.py:6653 backward) -- one pass, fixed orders
//   grad_peer_sum (world > 1) S summed over the ranks in fp32 through NVLink peer mappings: flag barrier, rank r sums
//                 slice r of every rank's S in rank order and stores it into every rank's reduced buffer, flag barrier
//                 (the scheme of peer_allreduce.cu, from one buffer into another)
//   grad_finish   S -> parameter gradients in the parameter dtype: dWv, dWk = scale q (x) R, d_qp = scale Wk . R,
//                 dWq = d_qp (x) q0, d_query = d_qp . Wq, the three in-projection bias thirds, dWo, d_out_proj_bias
//                 (reference: what autograd derives from torch/nn/functional.py:5854-5855, 6653 for a shared query)
//
// Everything downstream of S is linear in S, so summing S over the ranks BEFORE grad_finish gives every rank the gradients
// of the global batch with one rounding to the parameter dtype -- an N-rank run rounds like a 1-rank run (r1 reduced
// bf16-rounded gradients).  S is half the size of the parameter set (dWk, dWq, d_query are images of the H rows of R).
// All reductions are in fixed index order; the "last block" pattern only decides WHO does a final fold, never its order.
#include "common.cuh"
#include "grad_tail.cuh"

namespace aecf {

constexpr int TAIL_THREADS = 256;

// ---- grad_gather ---------------------------------------------------------------------------------------------------
struct GatherParams {
    const float* g_part; int g_splits; long long g_stride;      // [dWv ; R] partials: g_splits slabs of (D + HSP) * D
    const float* o_part; int o_splits; long long o_stride;      // dWo partials: o_splits slabs of D * D
    const void* d_out; long long rows, ld;                      // [rows, D] in the parameter dtype: column sums -> bo
    const float* pool_part; int pool_blocks;                    // [pool_blocks][3 D] per-CTA partials of the pool backward
    float* sums; TailLayout lay;
    float* colsum_part;                                         // [gridDim.x][D] scratch
    unsigned* ticket;                                           // zero on entry, zero again on exit
    int D, HSP;
};

template <typename T>
__global__ void __launch_bounds__(TAIL_THREADS)
grad_gather_kernel(const GatherParams p) {
    constexpr int V = Vec<T>::N;
    __shared__ float red[TAIL_THREADS * 8];                     // [row lane][column] of one pass over the chunk columns
    __shared__ int last;
    const int t = threadIdx.x;
    const int D = p.D;
    const int NC = D / V;                                       // 16-byte chunks per row of d_out
    const int ncp = NC < TAIL_THREADS ? NC : TAIL_THREADS;      // chunk columns handled per pass
    const int rl = TAIL_THREADS / ncp;                          // row lanes per pass
    pdl_wait();

    // (1) column sums of this block's rows of d_out
    if (p.d_out != nullptr) {
        const long long per = (p.rows + gridDim.x - 1) / gridDim.x;
        const long long r0 = per * blockIdx.x, r1 = min(p.rows, r0 + per);
        const T* x = static_cast<const T*>(p.d_out);
        for (int cb = 0; cb < NC; cb += ncp) {
            const int cc = t % ncp, lane_r = t / ncp;
            const bool active = lane_r < rl && cb + cc < NC;
            float acc[V];
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] = 0.f;
            if (active) {
                const T* col = x + static_cast<long long>(cb + cc) * V;
                constexpr int U = 8;
                long long r = r0 + lane_r;
                for (; r + static_cast<long long>(U - 1) * rl < r1; r += static_cast<long long>(U) * rl) {
                    uint4 raw[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) raw[u] = ldg_stream(col + (r + static_cast<long long>(u) * rl) * p.ld);
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        float f[V];
                        Vec<T>::unpack(raw[u], f);
#pragma unroll
                        for (int v = 0; v < V; ++v) acc[v] += f[v];
                    }
                }
                for (; r < r1; r += rl) {
                    float f[V];
                    Vec<T>::unpack(ldg_stream(col + r * p.ld), f);
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[v] += f[v];
                }
            }
            __syncthreads();                                    // the previous pass has been read
            if (active) {
#pragma unroll
                for (int v = 0; v < V; ++v) red[(lane_r * ncp + cc) * V + v] = acc[v];
            }
            __syncthreads();
            for (int c = t; c < ncp * V && cb * V + c < D; c += TAIL_THREADS) {
                float s = 0.f;
                for (int y = 0; y < rl; ++y) s += red[y * ncp * V + c];
                p.colsum_part[static_cast<long long>(blockIdx.x) * D + cb * V + c] = s;
            }
        }
    }

    // (2) the pool backward's per-CTA partials [d_q | d_bias_v | d_bias_k], one column per thread, blocks in order
    const long long gt = static_cast<long long>(blockIdx.x) * TAIL_THREADS + t;
    if (p.pool_part != nullptr && gt < 3LL * D) {
        float s = 0.f;
        constexpr int U = 8;
        int b = 0;
        for (; b + U <= p.pool_blocks; b += U) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = p.pool_part[static_cast<long long>(b + u) * 3 * D + gt];
#pragma unroll
            for (int u = 0; u < U; ++u) s += v[u];
        }
        for (; b < p.pool_blocks; ++b) s += p.pool_part[static_cast<long long>(b) * 3 * D + gt];
        p.sums[p.lay.pool + gt] = s;
    }

    // (3) split-K partials of the two weight-gradient products, splits in order, 16 bytes per thread
    const long long nthreads = static_cast<long long>(gridDim.x) * TAIL_THREADS;
    auto fold = [&](const float* part, int splits, long long stride, long long n, float* out) {
        if (part == nullptr) return;
        for (long long i = gt; i < n / 4; i += nthreads) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int z = 0; z < splits; ++z) {
                const float4 v = *reinterpret_cast<const float4*>(part + z * stride + 4 * i);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            *reinterpret_cast<float4*>(out + 4 * i) = s;
        }
    };
    fold(p.g_part, p.g_splits, p.g_stride, static_cast<long long>(D + p.HSP) * D, p.sums + p.lay.g);
    fold(p.o_part, p.o_splits, p.o_stride, static_cast<long long>(D) * D, p.sums + p.lay.o);

    // (4) the last block to get here folds the column-sum partials, blocks in order
    if (p.d_out == nullptr) return;
    __threadfence();
    __syncthreads();
    if (t == 0) last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int c = t; c < D; c += TAIL_THREADS) {
        float s = 0.f;
        constexpr int U = 8;
        int b = 0;
        for (; b + U <= static_cast<int>(gridDim.x); b += U) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = p.colsum_part[static_cast<long long>(b + u) * D + c];
#pragma unroll
            for (int u = 0; u < U; ++u) s += v[u];
        }
        for (; b < static_cast<int>(gridDim.x); ++b) s += p.colsum_part[static_cast<long long>(b) * D + c];
        p.sums[p.lay.bo + c] = s;
    }
    if (t == 0) *p.ticket = 0u;                                 // re-armed for the next call (stream order makes it visible)
}

// ---- grad_finish ---------------------------------------------------------------------------------------------------
struct FinishParams {
    const float* sums; TailLayout lay;
    const float* q_proj;            // [D] fp32: projected query (unscaled)
    const void* in_proj_weight;     // [3D, D]
    const void* query;              // [D]
    void* d_in_w;                   // [3D, D], nullable
    void* d_in_b;                   // [3D], nullable
    void* d_out_w;                  // [D, D], nullable
    void* d_out_b;                  // [D], nullable
    void* d_query;                  // [D], nullable
    float* dq_part;                 // [gridDim.x][D] scratch
    unsigned* ticket;
    int D, H, HSP;
    float scale;
};

// Block b owns the in-projection rows i = 8 b .. 8 b + 7, one warp per row (head h = i / head_dim):
//   dWv[i, :] = G[i, :]     dWk[i, :] = scale q[i] R[h, :]     d_qp[i] = scale Wk[i, :] . R[h, :]     dWo[i, :] = O[i, :]
//   dWq[i, :] = d_qp[i] q0[:]     d_query[:] += d_qp[i] Wq[i, :]  (8 rows folded in shared memory, blocks by the last block)
//   d_in_b = [ d_qp | d_bias_k | d_bias_v ]     d_out_b = colsum(d_out)
template <typename T>
__global__ void __launch_bounds__(TAIL_THREADS)
grad_finish_kernel(const FinishParams p) {
    __shared__ float red[8][128 * 4 + 4];                       // one 512-column tile of the 8 rows' d_query terms
    __shared__ int last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = p.D;
    const int i = blockIdx.x * 8 + warp;
    const bool row_ok = i < D;
    pdl_wait();
    const T* W = static_cast<const T*>(p.in_proj_weight);
    const T* q0 = static_cast<const T*>(p.query);
    T* d_in_w = static_cast<T*>(p.d_in_w);
    float d_qp = 0.f;
    if (row_ok) {
        const int h = i / (D / p.H);
        const float* R = p.sums + p.lay.g + static_cast<long long>(D + h) * D;
        const float* G = p.sums + p.lay.g + static_cast<long long>(i) * D;
        const float* O = p.sums + p.lay.o + static_cast<long long>(i) * D;
        const T* wk = W + (static_cast<long long>(D) + i) * D;
        const float sq = p.scale * __ldg(p.q_proj + i);
        float dot = 0.f;
        for (int d = lane * 4; d < D; d += 128) {
            const float4 r = *reinterpret_cast<const float4*>(R + d);
            dot = fmaf(to_float<T>(wk[d]), r.x, dot); dot = fmaf(to_float<T>(wk[d + 1]), r.y, dot);
            dot = fmaf(to_float<T>(wk[d + 2]), r.z, dot); dot = fmaf(to_float<T>(wk[d + 3]), r.w, dot);
            if (d_in_w != nullptr) {
                T* dwk = d_in_w + (static_cast<long long>(D) + i) * D + d;
                T* dwv = d_in_w + (2LL * D + i) * D + d;
                const float4 g = *reinterpret_cast<const float4*>(G + d);
                dwk[0] = from_float<T>(sq * r.x); dwk[1] = from_float<T>(sq * r.y);
                dwk[2] = from_float<T>(sq * r.z); dwk[3] = from_float<T>(sq * r.w);
                dwv[0] = from_float<T>(g.x); dwv[1] = from_float<T>(g.y); dwv[2] = from_float<T>(g.z); dwv[3] = from_float<T>(g.w);
            }
            if (p.d_out_w != nullptr) {
                T* dwo = static_cast<T*>(p.d_out_w) + static_cast<long long>(i) * D + d;
                const float4 o = *reinterpret_cast<const float4*>(O + d);
                dwo[0] = from_float<T>(o.x); dwo[1] = from_float<T>(o.y); dwo[2] = from_float<T>(o.z); dwo[3] = from_float<T>(o.w);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(FULL_MASK, dot, off);
        d_qp = p.scale * dot;
        if (lane == 0) {
            if (p.d_in_b != nullptr) {
                T* db = static_cast<T*>(p.d_in_b);
                db[i] = from_float<T>(d_qp);
                db[D + i] = from_float<T>(p.sums[p.lay.pool + 2LL * D + i]);       // d_bias_k
                db[2 * D + i] = from_float<T>(p.sums[p.lay.pool + D + i]);         // d_bias_v
            }
            if (p.d_out_b != nullptr) static_cast<T*>(p.d_out_b)[i] = from_float<T>(p.sums[p.lay.bo + i]);
        }
    }
    // query side, 512 columns at a time
    for (int c0 = 0; c0 < D; c0 += 512) {
        float4 term = make_float4(0.f, 0.f, 0.f, 0.f);
        const int d = c0 + lane * 4;
        // each lane covers columns c0 + 4 lane + {0..3} + 128 k, k = 0..3
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int dd = d + 128 * k;
            term = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_ok && dd < D) {
                if (d_in_w != nullptr) {
                    T* dwq = d_in_w + static_cast<long long>(i) * D + dd;
                    dwq[0] = from_float<T>(d_qp * to_float<T>(q0[dd])); dwq[1] = from_float<T>(d_qp * to_float<T>(q0[dd + 1]));
                    dwq[2] = from_float<T>(d_qp * to_float<T>(q0[dd + 2])); dwq[3] = from_float<T>(d_qp * to_float<T>(q0[dd + 3]));
                }
                const T* wq = W + static_cast<long long>(i) * D + dd;
                term = make_float4(d_qp * to_float<T>(wq[0]), d_qp * to_float<T>(wq[1]), d_qp * to_float<T>(wq[2]),
                                   d_qp * to_float<T>(wq[3]));
            }
            *reinterpret_cast<float4*>(&red[warp][(k * 32 + lane) * 4]) = term;
        }
        __syncthreads();
        // column c0 + 128 k + 4 lane + e  <->  red[.][(k * 32 + lane) * 4 + e]: the layout is column-linear
        for (int c = threadIdx.x; c < 512 && c0 + c < D; c += TAIL_THREADS) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w][c];
            p.dq_part[static_cast<long long>(blockIdx.x) * D + c0 + c] = s;
        }
        __syncthreads();
    }
    if (p.d_query == nullptr) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int c = threadIdx.x; c < D; c += TAIL_THREADS) {
        float s = 0.f;
        for (int b = 0; b < static_cast<int>(gridDim.x); ++b) s += p.dq_part[static_cast<long long>(b) * D + c];
        static_cast<T*>(p.d_query)[c] = from_float<T>(s);
    }
    if (threadIdx.x == 0) *p.ticket = 0u;
}

// ---- host side ----------------------------------------------------------------------------------------------------------
size_t grad_tail_scratch_bytes(int D, int sms) {
    const size_t gather = static_cast<size_t>(2 * sms) * D * sizeof(float);              // colsum partials
    const size_t finish = static_cast<size_t>((D + 7) / 8) * D * sizeof(float);           // d_query partials
    return ((gather + 255) & ~static_cast<size_t>(255)) + ((finish + 255) & ~static_cast<size_t>(255)) + 256;   // + tickets
}

int launch_grad_gather(const GradTailArgs& a, cudaStream_t s) {
    GatherParams p{};
    p.g_part = a.g.partial; p.g_splits = a.g.splits; p.g_stride = a.g.stride;
    p.o_part = a.o.partial; p.o_splits = a.o.splits; p.o_stride = a.o.stride;
    p.d_out = a.d_out; p.rows = a.rows; p.ld = a.D;
    p.pool_part = a.pool_part; p.pool_blocks = a.pool_blocks;
    p.sums = a.sums; p.lay = tail_layout(a.D, a.HSP);
    p.D = a.D; p.HSP = a.HSP;
    char* scratch = static_cast<char*>(a.scratch);
    const size_t gather = (static_cast<size_t>(2 * a.sms) * a.D * sizeof(float) + 255) & ~static_cast<size_t>(255);
    const size_t finish = (static_cast<size_t>((a.D + 7) / 8) * a.D * sizeof(float) + 255) & ~static_cast<size_t>(255);
    p.colsum_part = reinterpret_cast<float*>(scratch);
    p.ticket = reinterpret_cast<unsigned*>(scratch + gather + finish);
    const int V = a.dtype == AECF_BF16 ? 8 : 4;
    if (a.D % V != 0) return AECF_ERR_UNSUPPORTED;
    const dim3 grid(static_cast<unsigned>(2 * a.sms)), block(TAIL_THREADS);
    TimedLaunch timed(s, AECF_SITE_GRAD_GATHER);
    if (a.dtype == AECF_BF16) AECF_CUDA_OK(launch_pdl(grad_gather_kernel<__nv_bfloat16>, grid, block, 0, s, p));
    else AECF_CUDA_OK(launch_pdl(grad_gather_kernel<float>, grid, block, 0, s, p));
    count_launch();
    return AECF_OK;
}

int launch_grad_finish(const GradTailArgs& a, const float* sums, cudaStream_t s) {
    FinishParams p{};
    p.sums = sums; p.lay = tail_layout(a.D, a.HSP);
    p.q_proj = a.q_proj; p.in_proj_weight = a.in_proj_weight; p.query = a.query;
    p.d_in_w = a.d_in_w; p.d_in_b = a.d_in_b; p.d_out_w = a.d_out_w; p.d_out_b = a.d_out_b; p.d_query = a.d_query;
    p.D = a.D; p.H = a.H; p.HSP = a.HSP;
    p.scale = static_cast<float>(sqrt(1.0 / static_cast<double>(a.D / a.H)));
    char* scratch = static_cast<char*>(a.scratch);
    const size_t gather = (static_cast<size_t>(2 * a.sms) * a.D * sizeof(float) + 255) & ~static_cast<size_t>(255);
    const size_t finish = (static_cast<size_t>((a.D + 7) / 8) * a.D * sizeof(float) + 255) & ~static_cast<size_t>(255);
    p.dq_part = reinterpret_cast<float*>(scratch + gather);
    p.ticket = reinterpret_cast<unsigned*>(scratch + gather + finish) + 1;
    if (a.D % 4 != 0) return AECF_ERR_UNSUPPORTED;
    const dim3 grid(static_cast<unsigned>((a.D + 7) / 8)), block(TAIL_THREADS);
    TimedLaunch timed(s, AECF_SITE_GRAD_FINISH);
    if (a.dtype == AECF_BF16) AECF_CUDA_OK(launch_pdl(grad_finish_kernel<__nv_bfloat16>, grid, block, 0, s, p));
    else AECF_CUDA_OK(launch_pdl(grad_finish_kernel<float>, grid, block, 0, s, p));
    count_launch();
    return AECF_OK;
}

}  // namespace aecf
