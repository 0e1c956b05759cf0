// Host-side construction of CUtensorMap descriptors without linking libcuda: the driver entry point is fetched
// through the runtime (cudaGetDriverEntryPoint), so the shared library loads on machines with no driver
// (the CPU build box) and only fails, with a message, when a kernel is actually requested.
//
// Reference counterpart: the `gl<>` constructor re-encodes four tensor maps on every call
// (src/quantum_attn/tk_repo/include/types/global/tma.cuh:30-159, used from src/quantum_attn/tk/attention.py:487-491).
// Here descriptors describe [B*H, S, D] views and are built per launch from plain pointers (≈1 µs each); the
// kernel receives them as __grid_constant__ parameters, so nothing is copied to device memory.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <mutex>

namespace qa {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

// 3-D map over a dense [n_outer, n_rows, row_elems] tensor (row_elems contiguous).
// box = [1, box_rows, box_elems]; swizzle chosen by the caller to match the UMMA descriptor the kernel builds.
inline bool make_tmap_3d(CUtensorMap* out, CUtensorMapDataType dt, uint32_t elem_bytes, const void* base,
                         uint64_t row_elems, uint64_t n_rows, uint64_t n_outer, uint64_t row_stride_bytes,
                         uint64_t outer_stride_bytes, uint32_t box_elems, uint32_t box_rows,
                         CUtensorMapSwizzle swz) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {row_elems, n_rows, n_outer};
    cuuint64_t strides[2] = {row_stride_bytes, outer_stride_bytes};
    cuuint32_t box[3] = {box_elems, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    (void)elem_bytes;
    CUresult r = enc(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// 4-D map over a [B, H, S, D] tensor with arbitrary (16-byte aligned) batch / head / row strides, D contiguous:
// dims = {D, S, H, B}, box = [1, 1, box_rows, box_elems].  This is what lets the kernels consume a [B, S, H, D]-held
// tensor (the usual layout of DiT / Llama projections) in place - the reference copies it dense first
// (src/quantum_attn/tk/attention.py:419-421).
inline bool make_tmap_4d(CUtensorMap* out, CUtensorMapDataType dt, const void* base, uint64_t row_elems, uint64_t n_rows,
                         uint64_t n_heads, uint64_t n_batch, uint64_t row_stride_bytes, uint64_t head_stride_bytes,
                         uint64_t batch_stride_bytes, uint32_t box_elems, uint32_t box_rows, CUtensorMapSwizzle swz) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[4] = {row_elems, n_rows, n_heads, n_batch};
    cuuint64_t strides[3] = {row_stride_bytes, head_stride_bytes, batch_stride_bytes};
    cuuint32_t box[4] = {box_elems, box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, dt, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace qa
