// Row-wise combine of two partial attention results over disjoint key sets (HBM-bound):
//     (O_a, LSE_a) (+) (O_b, LSE_b):   m = max(LSE_a, LSE_b),  w_x = exp(LSE_x - m),
//                                      O = (w_a O_a + w_b O_b) / (w_a + w_b),   LSE = m + log(w_a + w_b)
// Used by the sequence ring (quantumattention_b200/parallel.py): each ring step attends the local queries to one block
// of keys and emits (O, LSE); the running result is kept in fp32.  The reference has no such operator - its kernel
// leaves the LSE output commented out (src/quantum_attn/tk/attention.py:333-346) and it never shards a sequence.
//
// One 16-byte vector of the 16-bit partial (8 elements) per thread, D / 8 neighbouring threads per row, grid-stride
// over rows.  Bytes per element: first step 2 in + 4 out, middle steps 4 + 2 in + 4 out, last step 4 + 2 in + 2 out.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "qattn_internal.h"

namespace qa {

constexpr int kMergeThreads = 256;

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]);
template <>
__device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& v, float (&f)[8]) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(p[i]);
        f[2 * i] = t.x, f[2 * i + 1] = t.y;
    }
}
template <>
__device__ __forceinline__ void unpack8<__half>(const uint4& v, float (&f)[8]) {
    const __half2* p = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(p[i]);
        f[2 * i] = t.x, f[2 * i + 1] = t.y;
    }
}
template <typename T>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]);
template <>
__device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float (&f)[8]) {
    uint4 v;
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
}
template <>
__device__ __forceinline__ uint4 pack8<__half>(const float (&f)[8]) {
    uint4 v;
    __half2* p = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    return v;
}

// FIRST: the accumulator is uninitialised (O_acc = O_new, LSE_acc = LSE_new).  `out` (nullable) additionally receives
// the merged rows in the 16-bit type - the last ring step - in which case the fp32 accumulator is not written back.
template <typename T, bool FIRST>
__global__ void __launch_bounds__(kMergeThreads)
merge_partials_kernel(float* __restrict__ o_acc, float* __restrict__ lse_acc, const T* __restrict__ o_new,
                      const float* __restrict__ lse_new, T* __restrict__ out, long long rows, int D) {
    const int vec_per_row = D >> 3;
    const int rows_per_cta = kMergeThreads / vec_per_row;
    const int v = threadIdx.x % vec_per_row;
    const int r_in = threadIdx.x / vec_per_row;
    for (long long r = (long long)blockIdx.x * rows_per_cta + r_in; r < rows; r += (long long)gridDim.x * rows_per_cta) {
        const long long e = r * D + v * 8;
        float fn[8];
        unpack8<T>(__ldcs(reinterpret_cast<const uint4*>(o_new + e)), fn);
        const float ln = lse_new[r];
        float fo[8], lse;
        if constexpr (FIRST) {
#pragma unroll
            for (int i = 0; i < 8; ++i) fo[i] = fn[i];
            lse = ln;
        } else {
            const float4 a0 = __ldcs(reinterpret_cast<const float4*>(o_acc + e));
            const float4 a1 = __ldcs(reinterpret_cast<const float4*>(o_acc + e) + 1);
            const float fa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float la = lse_acc[r];
            const float m = fmaxf(la, ln);
            // a side without any visible key has LSE = -inf and weight 0 (also when both sides are empty)
            const float wa = (la == -INFINITY) ? 0.f : __expf(la - m);
            const float wn = (ln == -INFINITY) ? 0.f : __expf(ln - m);
            const float ws = wa + wn;
            const float inv = ws > 0.f ? 1.f / ws : 0.f;
            const float ca = wa * inv, cn = wn * inv;
#pragma unroll
            for (int i = 0; i < 8; ++i) fo[i] = ca * fa[i] + cn * fn[i];
            lse = ws > 0.f ? m + __logf(ws) : -INFINITY;
        }
        if (out != nullptr) {
            *reinterpret_cast<uint4*>(out + e) = pack8<T>(fo);
        } else {
            float4* dst = reinterpret_cast<float4*>(o_acc + e);
            dst[0] = make_float4(fo[0], fo[1], fo[2], fo[3]);
            dst[1] = make_float4(fo[4], fo[5], fo[6], fo[7]);
        }
        if (v == 0) lse_acc[r] = lse;
    }
}

template <typename T>
static void launch_merge(const MergeArgs& a, int grid, cudaStream_t stream) {
    const T* o_new = static_cast<const T*>(a.o_new);
    T* out = static_cast<T*>(a.out);
    if (a.first)
        merge_partials_kernel<T, true><<<grid, kMergeThreads, 0, stream>>>(a.o_acc, a.lse_acc, o_new, a.lse_new, out, a.rows, a.D);
    else
        merge_partials_kernel<T, false><<<grid, kMergeThreads, 0, stream>>>(a.o_acc, a.lse_acc, o_new, a.lse_new, out, a.rows, a.D);
}

int merge_dispatch(const MergeArgs& a, cudaStream_t stream, int* launches) {
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const int rows_per_cta = kMergeThreads / (a.D >> 3);
    const long long want = (a.rows + rows_per_cta - 1) / rows_per_cta;
    const long long cap = (long long)sms * 8 * 4;  // 8 resident CTAs per SM, a few trips each
    const int grid = int(want < cap ? want : cap);
    if (a.dtype == QA_DT_BF16) launch_merge<__nv_bfloat16>(a, grid, stream);
    else launch_merge<__half>(a, grid, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("merge_partials_kernel launch", e);
    *launches += 1;
    return QA_OK;
}

}  // namespace qa
